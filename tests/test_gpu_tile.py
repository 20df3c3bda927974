"""GPU parity tests of the tile fast path v4 (csrc/tile_kernel.cu: tile-CSR DMMA products over 8
scenarios per CTA + per-scenario block LDL^T on 8x8 register tiles) against the general kernel
and the NumPy oracle.  Tolerances: BASELINE.json (moves and cost 1e-6 relative, violation 1e-8).
"""
import numpy as np
import pytest

from tests.util import RTOL, assert_loop_close, oracle_closed_loop
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController

pytestmark = pytest.mark.gpu


def _plant(ny, nu, seed, max_delay):
    rng = np.random.default_rng(seed)
    K = rng.normal(0.0, 0.15, size=(ny, nu))      # diagonally dominant: a well-conditioned gain matrix
    for i in range(min(ny, nu)):
        K[i, i] = rng.uniform(1.0, 2.0)
    tau = rng.uniform(4.0, 20.0, size=(ny, nu))
    theta = rng.integers(0, max_delay + 1, size=(ny, nu)).astype(float)
    return sc.fopdt_system(K, tau, theta, 1.0)


def _scenarios(B, sys_, seed):
    rng = np.random.default_rng(seed)
    ny, nu, nx = sys_.ny, sys_.nu, sys_.nx
    a = rng.uniform(-0.5, 0.5, size=(B, ny)); b = rng.uniform(-0.5, 0.5, size=(B, ny))
    seq = np.stack([np.zeros((B, ny)), a, a + b], axis=1)
    start = np.tile(np.array([0, 3, 17], dtype=np.int32), (B, 1))
    return dict(x0=np.zeros((B, nx)), u0=np.zeros((B, nu)), ysp_seq=seq, seg_start=start,
                dist_step=rng.integers(5, 25, size=B).astype(np.int32), dist=rng.normal(0, 0.1, size=(B, ny)))


def _run(monkeypatch, mode, declare, sys_, scen, T, bounds=None, spl=None, weights="sat2", alfa=0.0, inkernel=True):
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    monkeypatch.delenv("SIHMPC_FAST", raising=False)
    monkeypatch.delenv("SIHMPC_NO_INKERNEL_AS", raising=False)
    if not inkernel:
        monkeypatch.setenv("SIHMPC_NO_INKERNEL_AS", "1")
    if mode == "general":
        monkeypatch.setenv("SIHMPC_NO_FAST", "1")
    elif mode == "tile":
        monkeypatch.setenv("SIHMPC_FAST", "tile")
    c = declare(IHMPCController, sys_, **(bounds or {}))
    B = scen["x0"].shape[0]
    eng = c.engine(B, weights=weights, alfa=alfa) if (weights != "sat2" or alfa) else c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    if "dist_step" in scen:
        eng.set_disturbance(scen["dist_step"], scen["dist"])
    if "plant_val" in scen:
        eng.set_plant_values(scen["plant_val"])
    tr = eng.run(T, steps_per_launch=spl, record=True)
    out = dict(du=tr["du"].cpu().numpy(), J=tr["J"].cpu().numpy(), y=tr["y"].cpu().numpy(),
               x=eng.x.cpu().numpy(), u=eng.u.cpu().numpy(), pesos=eng.pesos.cpu().numpy(),
               stats=eng.stats.cpu().numpy(), w=eng.w_opt.cpu().numpy(), cnt=eng.dp.counters())
    monkeypatch.delenv("SIHMPC_NO_INKERNEL_AS", raising=False)
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    monkeypatch.delenv("SIHMPC_FAST", raising=False)
    return out


def _close(a, b, tol=1e-8):
    for k in ("du", "J", "y", "x", "u", "w"):
        scale = max(1.0, float(np.max(np.abs(b[k]))))
        assert np.max(np.abs(a[k] - b[k])) <= tol * scale, (k, np.max(np.abs(a[k] - b[k])))
    assert np.max(np.abs(a["pesos"] - b["pesos"]) / np.abs(b["pesos"])) <= 1e-6
    ny = a["y"].shape[2]
    assert np.allclose(a["stats"][:, :ny + 4], b["stats"][:, :ny + 4], rtol=1e-7, atol=1e-9)


# (ny, nu, N, max_delay): nv = N*nu covers every remainder mod 4 and mod 8, one to eight tile rows
SHAPES = [(2, 2, 4, 2),      # nv = 8 : right-hand side is the first row of its own tile row
          (3, 3, 3, 2),      # nv = 9 : one real column in the last pivot block
          (3, 5, 2, 1),      # nv = 10
          (3, 3, 5, 3),      # nv = 15
          (2, 2, 10, 7),     # nv = 20
          (5, 5, 6, 4),      # nv = 30
          (4, 4, 8, 4),      # nv = 32
          (7, 7, 9, 3)]      # nv = 63: the largest the path accepts (8 tile rows)


@pytest.mark.parametrize("ny,nu,N,dmax", SHAPES)
def test_tile_equals_general_on_random_shapes(ny, nu, N, dmax, monkeypatch):
    sys_ = _plant(ny, nu, 100 + 7 * ny + nu, dmax)
    declare = lambda cls, s, **b: sc.declare_mimo10(cls, s, N=N, **(b or dict(ulb=[-50.0] * nu, uub=[50.0] * nu)))
    B, T = 21, 30
    scen = _scenarios(B, sys_, 5)
    t = _run(monkeypatch, "tile", declare, sys_, scen, T, spl=13)
    g = _run(monkeypatch, "general", declare, sys_, scen, T, spl=13)
    assert (t["stats"][:, ny + 3] == 0).all() and (g["stats"][:, ny + 3] == 0).all()
    assert (t["stats"][:, ny + 4] == 0).all()            # nothing was handed over: the tile path did the work
    _close(t, g)
    for idx in (3, 11, 20):
        ref = oracle_closed_loop(declare, sys_, scen, idx, T)
        assert_loop_close(t["du"][:, idx], t["J"][:, idx], ref, tag=idx)


def test_tile_underdamped_plant_matches_oracle(monkeypatch):
    """Complex pole pairs in the model (block-diagonal F, nd = 12): tile path vs the oracle and the general kernel."""
    from tests.test_plan_cpu import _underdamped_system
    sys_, _ = _underdamped_system()
    declare = lambda cls, s, **b: sc.declare_mimo10(cls, s, N=6, **(b or dict(ulb=[-3.0, -3.0], uub=[3.0, 3.0])))
    scen = _scenarios(24, sys_, 21)
    t = _run(monkeypatch, "tile", declare, sys_, scen, 40)
    g = _run(monkeypatch, "general", declare, sys_, scen, 40)
    assert (t["stats"][:, 2 + 3] == 0).all()
    _close(t, g)
    for idx in (1, 17):
        ref = oracle_closed_loop(declare, sys_, scen, idx, 40)
        assert np.max(np.abs(t["du"][:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"]))
        assert np.max(np.abs(t["y"][:, idx] - ref["y"])) <= RTOL * max(1.0, np.max(np.abs(ref["y"])))


@pytest.mark.parametrize("inkernel", [False, True])
def test_tile_path_on_distillation_with_handover(inkernel, monkeypatch):
    """Forced tile path on the distillation controller with bounds that bind for a good share of the scenarios.
    inkernel=False: they are handed over to the exact active-set kernel mid-loop; inkernel=True: the resume pass of
    the tile kernel resolves them itself (tile_active_set.inc).  Either way: equals the general kernel alone and
    the oracle."""
    sys_ = sc.distillation_system()
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])
    B, T = 500, 120
    scen = sc.distillation_scenarios(B)
    t = _run(monkeypatch, "tile", sc.declare_distillation, sys_, scen, T, bounds=bounds, spl=50, inkernel=inkernel)
    g = _run(monkeypatch, "general", sc.declare_distillation, sys_, scen, T, bounds=bounds, spl=50)
    assert (t["stats"][:, 5] == 0).all()
    n_constrained = int((g["stats"][:, 6] > 0).sum())
    assert 0.05 * B < n_constrained < 0.95 * B
    handed, solved = t["cnt"]                              # of the last launch (steps 100..119)
    if inkernel:
        assert solved > 0 and handed <= 0.05 * B, (handed, solved)
    else:
        assert solved == 0 and handed > 0, (handed, solved)
    _close(t, g, tol=1e-8 if not inkernel else 2e-7)
    idx = int(np.argmax(g["stats"][:, 6]))
    ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, bounds=bounds)
    assert ref["n_active"].max() > 0
    assert np.max(np.abs(t["du"][:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"]))


def test_tile_path_mismatch_plants(monkeypatch):
    """Per-scenario plant values (CSR path inside the tile kernel) == general kernel."""
    sys_ = sc.distillation_system()
    scen = sc.mismatch_scenarios(77)
    t = _run(monkeypatch, "tile", sc.declare_distillation, sys_, scen, 60)
    g = _run(monkeypatch, "general", sc.declare_distillation, sys_, scen, 60)
    _close(t, g)


def test_tile_mimo10_with_active_bounds(monkeypatch):
    """Config 4 shape with an input box tight enough to bind: tile + hand-over == general kernel."""
    sys_ = sc.mimo10_system()
    bounds = dict(ulb=[-0.15] * 10, uub=[0.15] * 10)
    B, T = 40, 35
    scen = sc.mimo10_scenarios(B, sys_)
    scen["dist_step"][:] = np.minimum(scen["dist_step"], 20)
    t = _run(monkeypatch, "tile", sc.declare_mimo10, sys_, scen, T, bounds=bounds)
    g = _run(monkeypatch, "general", sc.declare_mimo10, sys_, scen, T, bounds=bounds)
    assert int((g["stats"][:, -1] > 0).sum()) > 0
    handed, solved = t["cnt"]
    assert solved > 0 and handed <= 2, (handed, solved)      # resolved inside the tile kernel, not handed over
    _close(t, g, tol=2e-7)
    idx = int(np.argmax(g["stats"][:, -1]))
    ref = oracle_closed_loop(sc.declare_mimo10, sys_, scen, idx, T, bounds=bounds)
    assert ref["n_active"].max() > 0
    assert np.max(np.abs(t["du"][:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"]))


def test_tile_bounce_rounds(monkeypatch):
    """Opt-in bounce rounds (SIHMPC_BOUNCE=1): scenarios the resume pass gives up on are carried through one step by
    the general kernel and sent back with the active set it found as the warm start of the in-kernel active set."""
    sys_ = sc.mimo10_system()
    bounds = dict(ulb=[-0.6] * 10, uub=[0.6] * 10)
    B, T = 64, 40
    scen = sc.mimo10_scenarios(B, sys_)
    scen["seg_start"][:, 1] = 3; scen["seg_start"][:, 2] = 20
    scen["dist_step"][:] = np.minimum(scen["dist_step"], 25)
    monkeypatch.setenv("SIHMPC_BOUNCE", "1")
    t = _run(monkeypatch, "tile", sc.declare_mimo10, sys_, scen, T, bounds=bounds)
    monkeypatch.delenv("SIHMPC_BOUNCE", raising=False)
    n = _run(monkeypatch, "tile", sc.declare_mimo10, sys_, scen, T, bounds=bounds)
    g = _run(monkeypatch, "general", sc.declare_mimo10, sys_, scen, T, bounds=bounds)
    assert int((g["stats"][:, -1] > 0).sum()) > 0
    assert t["cnt"][1] >= n["cnt"][1]                      # at least as many constrained steps resolved in the kernel
    _close(t, g, tol=2e-7)
    _close(n, g, tol=2e-7)


def test_tile_move_box_inkernel(monkeypatch):
    """Move box only (bound space = the moves themselves) on a 3x3 plant: resume pass == general kernel."""
    sys_ = _plant(3, 3, 9, 2)
    declare = lambda cls, s, **b: sc.declare_mimo10(cls, s, N=4, **(b or dict(dulb=[-0.03] * 3, duub=[0.04] * 3)))
    scen = _scenarios(50, sys_, 11)
    t = _run(monkeypatch, "tile", declare, sys_, scen, 30)
    g = _run(monkeypatch, "general", declare, sys_, scen, 30)
    assert int((g["stats"][:, -1] > 0).sum()) > 0
    handed, solved = t["cnt"]
    assert solved > 0 and handed <= 2, (handed, solved)
    _close(t, g, tol=2e-7)


@pytest.mark.parametrize("B", [1, 7, 9, 1000])
def test_tile_odd_batch_sizes(B, monkeypatch):
    sys_ = sc.mimo10_system()
    scen = sc.mimo10_scenarios(B, sys_)
    t = _run(monkeypatch, "tile", sc.declare_mimo10, sys_, scen, 12, spl=5)
    g = _run(monkeypatch, "general", sc.declare_mimo10, sys_, scen, 12, spl=5)
    _close(t, g)


@pytest.mark.parametrize("weights,alfa", [("sat1", 0.0), ("fixed", 0.0), ("sat2", 0.5)])
def test_tile_weight_rules(weights, alfa, monkeypatch):
    sys_ = _plant(3, 3, 9, 2)
    declare = lambda cls, s, **b: sc.declare_mimo10(cls, s, N=4, **(b or dict(ulb=[-50.0] * 3, uub=[50.0] * 3)))
    scen = _scenarios(19, sys_, 11)
    t = _run(monkeypatch, "tile", declare, sys_, scen, 25, weights=weights, alfa=alfa)
    g = _run(monkeypatch, "general", declare, sys_, scen, 25, weights=weights, alfa=alfa)
    _close(t, g)


def test_tile_single_step_solve(monkeypatch):
    """mpc() only (no advance) through the tile path: du, J, w_opt, ViNant."""
    sys_ = sc.mimo10_system()
    rng = np.random.default_rng(3)
    B = 37
    x = rng.normal(size=(B, sys_.nx)) * 0.05
    x[:, sys_.ny + sys_.nd:2 * sys_.ny + sys_.nd] = 0.0           # xi = 0 (non-integrating)
    u = rng.normal(size=(B, sys_.nu)) * 0.1
    ysp = rng.normal(size=(B, sys_.ny)) * 0.3
    outs = []
    for mode in ("tile", "general"):
        monkeypatch.delenv("SIHMPC_NO_FAST", raising=False); monkeypatch.delenv("SIHMPC_FAST", raising=False)
        monkeypatch.setenv("SIHMPC_FAST" if mode == "tile" else "SIHMPC_NO_FAST", "tile" if mode == "tile" else "1")
        c = sc.declare_mimo10(IHMPCController, sys_)
        eng = c.engine(B)
        eng.set_state(x, u)
        eng.solve(ysp)
        outs.append((eng.du.cpu().numpy(), eng.J.cpu().numpy(), eng.w_opt.cpu().numpy(), eng.vinant.cpu().numpy(),
                     eng.x.cpu().numpy()))
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False); monkeypatch.delenv("SIHMPC_FAST", raising=False)
    for a, b in zip(outs[0], outs[1]):
        fin = np.isfinite(b)
        assert (np.isfinite(a) == fin).all()
        assert np.max(np.abs(a[fin] - b[fin])) <= 1e-9 * max(1.0, np.max(np.abs(b[fin])))
    assert np.array_equal(outs[0][4], x)                           # no advance: the state is untouched


def test_tile_mimo10_box5_setpoint_transients_resolved_in_kernel(monkeypatch):
    """configs[3] as benchmarked: |u| <= 5, set-point changes at k = 11 and k = 101.  The scenarios below are the ones
    of the 65536 draw whose inputs saturate after the second change (large changes of the active set from one step to
    the next: up to 34 of 50 bounds active, the unconstrained optimum violates up to 48).  The in-kernel active set
    (primal-dual rounds + primal fall-back, tile_active_set.inc) must resolve EVERY such step -- nothing handed over.
    Parity is checked PER STEP FROM THE SAME STATE: once inputs saturate the closed loop amplifies a 1e-13 perturbation
    of the set-point to 3e-7 in the moves (the oracle against itself), so trajectories can only be compared to
    ~1e-5; the solve of each step from the state the device was in is compared at the north-star tolerance."""
    from oracle.ihmpc_oracle import OracleIHMPC
    from tests.util import schedule_fn
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    monkeypatch.setenv("SIHMPC_FAST", "tile")
    sys_ = sc.mimo10_system()
    hard = [14, 39, 135, 195, 252, 259, 316, 351]
    full = sc.mimo10_scenarios(max(hard) + 1, sys_)
    scen = {k: (v[hard] if isinstance(v, np.ndarray) else v) for k, v in full.items()}
    B, T, ny, nu = len(hard), 130, sys_.ny, sys_.nu
    c = sc.declare_mimo10(IHMPCController, sys_)                                  # default bounds: |u| <= 5
    eng = c.engine(B)
    assert eng.dp.kernel_name() == "tile_step_kernel"
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(scen["dist_step"], scen["dist"])
    eng.run(100)                                                                  # one launch up to the second change
    states, moves, costs, nact, handed, solved = [], [], [], [], 0, 0
    for k in range(100, T, 2):                      # then two steps per launch (the resume pass needs n_steps > 1),
        states.append((eng.x.cpu().numpy().copy(), eng.u.cpu().numpy().copy(), eng.pesos.cpu().numpy().copy()))
        tr = eng.run(2, record=True)                # keeping the state in front of every launch
        moves.append(tr["du"][0].cpu().numpy().copy()); costs.append(tr["J"][0].cpu().numpy().copy())
        umax = np.abs(states[-1][1] + moves[-1]).max(axis=1)
        nact.append((umax > 5.0 - 1e-9).astype(int))                               # first step of the pair sits on the box
        h, s_ = eng.dp.counters(); handed += h; solved += s_
    st = eng.stats.cpu().numpy()
    assert handed == 0 and solved >= 20, (handed, solved)
    assert (st[:, ny + 3] == 0).all() and st[:, ny + 1].max() <= 1e-8            # status clean, largest violation seen
    nact = np.array(nact)
    assert nact.sum() >= 8
    orc = sc.declare_mimo10(OracleIHMPC, sys_)
    checked = 0
    for col in range(4):
        f = schedule_fn(scen["ysp_seq"][col], scen["seg_start"][col])
        ks = [int(k) for k in np.nonzero(nact[:, col] > 0)[0]]
        ks = ks[:6] + ks[-2:]                                                     # the transient and the saturated tail
        for k in ks:
            x, u, pes = (a[col] for a in states[k])
            sol = orc.mpc(x0=x, ySP=f(100 + 2 * k), w0=[], u0=u, pesos=pes)
            dref = np.asarray(sol["du_opt"])                                       # nu x N
            assert sol["info"]["n_active"] > 0
            assert np.max(np.abs(moves[k][col] - dref[:, 0])) <= RTOL * np.max(np.abs(dref)), (col, k)
            assert abs(costs[k][col] - float(sol["J"])) <= RTOL * abs(float(sol["J"])), (col, k)
            checked += 1
    assert checked >= 12
    monkeypatch.delenv("SIHMPC_FAST", raising=False)
