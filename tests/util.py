"""Shared helpers for the tests: the same declaration run on the oracle and on the product."""
import numpy as np

from oracle.ihmpc_oracle import OracleIHMPC, closed_loop
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController

RTOL = 1e-6          # north_star: control moves and cost within 1e-6 relative
VIOL = 1e-8          # north_star: constraint violation <= 1e-8


def rel_err(a, b, floor=1e-9):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-3 * np.max(np.abs(b)))))


def assert_loop_close(du, J, ref, rtol=RTOL, tag=None, move_floor=1e-3, cost_floor=1e-6):
    """Closed-loop parity at the north-star tolerance, PER STEP: the moves of every step relative to the largest
    move of THAT step (floored at move_floor x the largest move of the trajectory, so steps at rest are not compared
    at rounding level), the cost of every step relative to itself (floored at cost_floor x the largest cost)."""
    rd, rj = np.asarray(ref["du"]), np.asarray(ref["J"])
    step_scale = np.maximum(np.max(np.abs(rd), axis=1), move_floor * np.max(np.abs(rd)))
    err = np.max(np.abs(np.asarray(du) - rd), axis=1) / step_scale
    assert np.max(err) <= rtol, (tag, "moves", int(np.argmax(err)), float(np.max(err)))
    if J is not None:
        ej = np.abs(np.asarray(J) - rj) / np.maximum(np.abs(rj), cost_floor * np.max(np.abs(rj)))
        assert np.max(ej) <= rtol, (tag, "cost", int(np.argmax(ej)), float(np.max(ej)))


def schedule_fn(seq, start):
    def f(k):
        seg = 0
        for g in range(1, len(start)):
            if start[g] <= k:
                seg = g
        return seq[seg]
    return f


def oracle_closed_loop(declare, sys, scen, idx, n_steps, bounds=None, plant=None):
    c = declare(OracleIHMPC, sys, **(bounds or {}))
    dist = None
    if "dist_step" in scen:
        dist = (int(scen["dist_step"][idx]), scen["dist"][idx])
    return closed_loop(c, scen["x0"][idx], scen["u0"][idx],
                       schedule_fn(scen["ysp_seq"][idx], scen["seg_start"][idx]), n_steps,
                       plant=plant, disturbance=dist)


def integrating_2x2():
    """2x2 plant whose four channels all integrate (K / (s (tau s + 1)) with dead times): siN != 0 on both outputs,
    so a siN sub-objective over BOTH outputs is a genuine ellipsoid row (ihmpccontroller.py:243-271)."""
    from sihmpc_b200.opom import OPOM, TransferFunction
    h = [[TransferFunction([1.0], [5.0, 1.0, 0.0], delay=1), TransferFunction([0.3], [4.0, 1.0, 0.0], delay=0)],
         [TransferFunction([0.2], [6.0, 1.0, 0.0], delay=0), TransferFunction([1.5], [3.0, 1.0, 0.0], delay=2)]]
    return OPOM(h, 1.0)


def declare_ellipsoid(cls, s, N=6, **bounds):
    if not bounds:
        bounds = dict(dulb=[-0.4, -0.4], duub=[0.4, 0.4])
    c = cls(s, N, **bounds)
    c.subObjComposed(y=[0], Q=1, sat=N * 1.0 ** 2)
    c.subObjComposed(y=[1], Q=1, sat=N * 1.0 ** 2)
    c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2)
    c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
    c.subObj(siN=[0, 1], Q=np.diag([1.0, 2.0]), addJ=False)      # sum_j q_j siN_j^2 <= ViN_ant
    c.set_terminal_objective()
    return c
