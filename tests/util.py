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
