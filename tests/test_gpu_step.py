"""GPU parity tests (run on the B200 box): CUDA step through the C ABI vs the NumPy oracle.

Tolerances are the ones BASELINE.json states: control moves and cost within 1e-6 relative,
constraint violation <= 1e-8.
"""
import numpy as np
import pytest

from tests.util import RTOL, VIOL, assert_loop_close, oracle_closed_loop, schedule_fn
from oracle.ihmpc_oracle import OracleIHMPC
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController

pytestmark = pytest.mark.gpu

BOUND_SETS = {
    "example": dict(ulb=[0, 0]),
    "free": dict(ulb=[-np.inf, -np.inf]),
    "tight": dict(ulb=[1.9, 1.6], uub=[2.0, 1.8], dulb=[-0.02, -0.02], duub=[0.02, 0.02]),
    "du_only": dict(dulb=[-0.01, -0.03], duub=[0.015, 0.02]),
}


def _random_states(rng, B):
    x = np.zeros((B, 22)); x[:, 0] = 96.0; x[:, 1] = 0.5
    x += rng.normal(size=(B, 22)) * 0.05
    x[:, 6:8] = 0.0                                   # non-integrating plant: xi stays 0
    u = np.tile([1.95, 1.71], (B, 1)) + rng.normal(size=(B, 2)) * 0.02
    ysp = np.tile([96.0, 1.0], (B, 1)) + rng.normal(size=(B, 2)) * 0.4
    pesos = np.array([0.1, 0.4, 4.4, 4.4]) * (1 + rng.random((B, 4)))
    return x, u, ysp, pesos


@pytest.mark.parametrize("bname", list(BOUND_SETS))
def test_single_step_parity_distillation(bname):
    bounds = BOUND_SETS[bname]
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_, **bounds)
    orc = sc.declare_distillation(OracleIHMPC, sys_, **bounds)
    B = 48
    rng = np.random.default_rng(123)
    x, u, ysp, pesos = _random_states(rng, B)
    eng = c.engine(B)
    eng.set_state(x, u, pesos)
    eng.solve(ysp)
    du = eng.du.cpu().numpy(); J = eng.J.cpu().numpy(); w = eng.w_opt.cpu().numpy()
    st = eng.status.cpu().numpy(); na = eng.n_active.cpu().numpy()
    assert (st == 0).all()
    n_act_or = []
    for b in range(B):
        sol = orc.mpc(x[b], ysp[b], [], u[b], pesos[b], ViN_ant=[np.inf, np.inf])
        n_act_or.append(sol["info"]["n_active"])
        scale = max(np.max(np.abs(sol["w_opt"][:20])), 1e-6)
        assert np.max(np.abs(w[b, :20] - sol["w_opt"][:20])) <= RTOL * scale, (b, bname)
        assert np.max(np.abs(w[b, 20:] - sol["w_opt"][20:])) <= RTOL * max(1.0, np.max(np.abs(sol["w_opt"][20:])))
        assert abs(J[b] - sol["J"]) <= RTOL * max(abs(sol["J"]), 1e-9), (b, bname, J[b], sol["J"])
        assert np.allclose(du[b], sol["du_opt"][:, 0], rtol=0, atol=RTOL * scale)
        # feasibility of the GPU answer on the reference's own rows
        kk = orc.kkt_report({"w_opt": w[b], "lam_rows": sol["lam_rows"]}, x[b], ysp[b], u[b], pesos[b],
                            [np.inf, np.inf])
        assert kk["violation"] <= VIOL
    if bname in ("tight", "du_only"):
        # duplicated normals (move box vs cumulative input box on the first move) make the active
        # SET non-unique; the optimum is unique, so only require that rows are indeed active
        assert max(n_act_or) > 0 and (na > 0).mean() > 0.9


def test_survey_self_check_value():
    """J* = 0.30935531, dU_0 = (-0.0143766510, -0.0668630750)."""
    c = sc.declare_distillation(IHMPCController)
    x = np.append([96, 0.5], np.zeros(20))
    sol = c.mpc(x0=x, ySP=[96, 1], w0=[], u0=[1.95, 1.71], pesos=[])
    assert abs(float(sol["J"]) - 0.30935531) < 1e-8
    assert np.allclose(c.du.reshape(-1), [-0.0143766510, -0.0668630750], atol=1e-9)
    w = sol["w_opt"].full().reshape(-1)
    assert np.allclose(w[20:22], [0.0213581761, -0.154677657], atol=1e-9)
    assert np.all(w[[15, 17, 18, 19]] == 0.0) or np.max(np.abs(w[[15, 17, 18, 19]])) < 1e-15


def test_closed_loop_parity_config3():
    """Config 3 scenarios (set-point steps + disturbance): GPU closed loop vs oracle loop."""
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    B, T = 4096, 200
    scen = sc.distillation_scenarios(B)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(scen["dist_step"], scen["dist"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy(); y = tr["y"].cpu().numpy()
    assert (eng.stats[:, 2 + 3].cpu().numpy() == 0).all()          # status OR
    picks = [0, 1, 777, 4095] + [int(i) for i in np.random.default_rng(2026).choice(B, 60, replace=False)]
    for idx in picks:                                              # 64 scenarios: moves and cost per step at 1e-6
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T)
        assert_loop_close(du[:, idx], J[:, idx], ref, tag=idx)
        assert np.max(np.abs(y[:, idx] - ref["y"])) <= RTOL * 100.0
    # size-independent properties over the whole batch: outputs end on the last set-point
    yend = y[-1]
    assert np.max(np.abs(yend - scen["ysp_seq"][:, 2] )) < 0.05
    x = eng.x.cpu().numpy()
    assert np.isfinite(x).all()


def test_steps_per_launch_invariance():
    """One 50-step launch and fifty 1-step launches must give bit-identical trajectories."""
    c = sc.declare_distillation(IHMPCController)
    B, T = 256, 50
    scen = sc.distillation_scenarios(B)
    outs = []
    for spl in (T, 1, 7):
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"])
        eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        eng.set_disturbance(scen["dist_step"], scen["dist"])
        tr = eng.run(T, steps_per_launch=spl, record=True)
        outs.append((tr["du"].cpu().numpy(), eng.x.cpu().numpy(), eng.stats.cpu().numpy()))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0])
        assert np.array_equal(o[1], outs[0][1])
        assert np.allclose(o[2], outs[0][2], rtol=1e-13, atol=0)


def test_closed_loop_constrained_parity():
    """Tight input/move boxes keep several rows active for many steps."""
    sys_ = sc.distillation_system()
    bounds = BOUND_SETS["tight"]
    c = sc.declare_distillation(IHMPCController, sys_, **bounds)
    B, T = 64, 120
    scen = sc.distillation_scenarios(B)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(scen["dist_step"], scen["dist"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy()
    st = eng.stats.cpu().numpy()
    assert (st[:, 5] == 0).all()
    assert st[:, 3].max() <= VIOL                                   # max violation seen
    assert st[:, 6].max() > 0                                       # active-set iterations happened
    J = tr["J"].cpu().numpy()
    for idx in range(0, 64, 4):
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, bounds=bounds)
        assert ref["n_active"].max() > 0
        assert_loop_close(du[:, idx], J[:, idx], ref, tag=idx)


def test_mismatch_closed_loop_parity():
    """Config 5: per-scenario plant (gain/time-constant mismatch), nominal controller."""
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    B, T = 128, 200
    scen = sc.mismatch_scenarios(B)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_plant_values(scen["plant_val"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy(); y = tr["y"].cpu().numpy(); J = tr["J"].cpu().numpy()
    for idx in [0, 17, 127] + [int(i) for i in np.random.default_rng(5).choice(B, 61, replace=False)]:
        plant = sc.plant_matrices_from_values(sys_, scen["plant_val"][idx])
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, plant=plant)
        assert_loop_close(du[:, idx], J[:, idx], ref, tag=idx)
        assert np.max(np.abs(y[:, idx] - ref["y"])) <= RTOL * 100.0


def test_mimo10_parity():
    """Config 4 shape (10x10, N=5, nv=50, nx=160): single steps and a short closed loop."""
    sys_ = sc.mimo10_system()
    c = sc.declare_mimo10(IHMPCController, sys_)
    B, T = 64, 40
    scen = sc.mimo10_scenarios(B, sys_)
    scen["dist_step"][:] = np.minimum(scen["dist_step"], 25)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(scen["dist_step"], scen["dist"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy()
    assert (eng.stats[:, sys_.ny + 3].cpu().numpy() == 0).all()
    for idx in [0, 21, 42, 63]:
        ref = oracle_closed_loop(sc.declare_mimo10, sys_, scen, idx, T)
        assert_loop_close(du[:, idx], J[:, idx], ref, tag=idx)


def test_dropin_example_loop_matches_oracle():
    """The reference's own driver loop (example_distillation.py:65-113) on the drop-in class."""
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    orc = sc.declare_distillation(OracleIHMPC, sys_)
    u = [1.95, 1.71]; x = np.append([96, 0.5], np.zeros(sys_.nx - 2)); ysp = [96, 0.5]
    uo, xo = np.array(u), x.copy()
    w0 = []; pesos = []; w0o = []; pesoso = []
    for k in range(30):
        if k > 10:
            ysp = [96, 1]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos)
        du = c.du
        res = c.dynF(x0=x, u0=u, du0=du)
        x = res['xkp1'].full(); u = res['ukp1'].full()
        w0 = c.warmStart(sol, ysp)
        pesos, _ = c.satWeights2(x, u, w0, ysp)
        solo = orc.mpc(x0=xo, ySP=ysp, w0=w0o, u0=uo, pesos=pesoso)
        r = orc.dynF(xo, uo, orc.du)
        xo, uo = r['xkp1'], r['ukp1']
        w0o = orc.warmStart(solo, ysp)
        pesoso, _ = orc.satWeights2(xo, uo, w0o, ysp)
        assert np.allclose(np.asarray(du).reshape(-1), orc.du.reshape(-1), rtol=0, atol=1e-9), k
        assert abs(float(sol['J']) - solo['J']) <= 1e-6 * max(abs(solo['J']), 1e-6)
        assert np.allclose(pesos, pesoso, rtol=1e-7)
        assert np.allclose(sol['x_pred'].full().reshape(-1), solo['x_pred'], atol=1e-8)
        assert np.allclose(sol['g'].full().reshape(-1), solo['g'], atol=1e-7)
    assert len(c.j_hist) == 30 and len(c.V[0].v_hist) == 30
    assert np.allclose([V.v_hist[-1] for V in c.V], [V.v_hist[-1] for V in orc.V], atol=1e-9)


def test_duals_match_oracle_on_the_example_loop():
    """sol['lam_w'] / sol['lam_g'] (ihmpccontroller.py:525-529; consumed as IPOPT warm start by
    example_distillation_log.py:101-103) on the reference's driver loop with an input box AND a move box that bind:
    multipliers of the move box (lam_w), of the U_pred rows and of the terminal residual equalities res1 / res2
    (lam_g) against the oracle's KKT multipliers of the un-reduced (w, g) problem, casadi's sign convention.
    Compared where the active set is not degenerate (every active multiplier well away from zero)."""
    sys_ = sc.distillation_system()
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78], dulb=[-0.03, -0.03], duub=[0.03, 0.03])
    c = sc.declare_distillation(IHMPCController, sys_, **bounds)
    orc = sc.declare_distillation(OracleIHMPC, sys_, **bounds)
    N, nu, ny, nx = c.N, c.nu, c.ny, c.nx
    nw = N * nu + 2 * ny
    ng = N * nx + N * nu + 2 * ny
    u = np.array([1.95, 1.71]); x = np.append([96, 0.5], np.zeros(nx - 2)); ysp = [96, 0.5]
    w0 = []; pesos = []
    compared = active_w = active_g = 0
    for k in range(40):
        if k > 5:
            ysp = [96.6, 0.2]
        if k > 22:
            ysp = [95.2, 1.1]
        # ViN_ant = inf: with the contraction bound at 0 (siN = 0 on this non-integrating plant) the rows res1 and
        # V_siN <= 0 share a normal and their multipliers are not unique (LICQ fails; the oracle's least squares
        # splits them evenly, IPOPT returns whatever it ends on)
        vin = [np.inf, np.inf]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos, ViN_ant=vin)
        solo = orc.mpc(x0=x, ySP=ysp, w0=[], u0=u, pesos=pesos, ViN_ant=vin)   # same state in: per-step parity
        lam_o = solo['lam_rows']
        lam_w, lam_g = sol['lam_w'].full().reshape(-1), sol['lam_g'].full().reshape(-1)
        assert lam_w.size == nw and lam_g.size == ng + len(c.V)           # dim_g of ctl.py:414-456
        assert np.allclose(sol['w_opt'].full().reshape(-1), solo['w_opt'], atol=1e-8)
        act = np.abs(lam_o[:nw + N * nx + N * nu])
        if not np.any((act > 0) & (act < 1e-7 * max(1.0, act.max()))):    # non-degenerate active set
            scale = max(1.0, np.max(np.abs(lam_o)))
            assert np.max(np.abs(lam_w - lam_o[:nw])) <= 1e-6 * scale, k
            assert np.max(np.abs(lam_g[:ng] - lam_o[nw:nw + ng])) <= 1e-6 * scale, k
            compared += 1
            active_w += int(np.count_nonzero(lam_w)); active_g += int(np.count_nonzero(lam_g[N * nx:N * nx + N * nu]))
        assert np.all(lam_g[ng:] == 0.0)                                   # vacuous rows 0 <= V_i
        du = c.du
        res = c.dynF(x0=x, u0=u, du0=du)
        x = res['xkp1'].full().reshape(-1); u = res['ukp1'].full().reshape(-1)
        w0 = c.warmStart(sol, ysp)
        pesos, _ = c.satWeights2(x, u, w0, ysp)
    assert compared >= 30 and active_w > 10 and active_g > 10, (compared, active_w, active_g)


def test_multi_index_siN_ellipsoid_matches_oracle():
    """subObj(siN=[0, 1], Q=diag(1, 2), addJ=False): ONE quadratic row q0 siN0^2 + q1 siN1^2 <= ViN_ant
    (ihmpccontroller.py:243-271) on a 2x2 plant whose channels integrate.  mpc() finds the row's multiplier by an outer
    search whose every trial is a device QP solve; moves, cost, the contracted limit and the multiplier against the
    oracle (same state in), the row itself to 1e-8."""
    from tests.util import integrating_2x2, declare_ellipsoid
    s = integrating_2x2()
    c = declare_ellipsoid(IHMPCController, s)
    orc = declare_ellipsoid(OracleIHMPC, s)
    N, nu, ny = c.N, c.nu, c.ny
    x = np.zeros(s.nx); u = np.zeros(2); w0 = []; pesos = []
    active = 0
    for k in range(30):
        ysp = [1.0, -0.5] if k < 15 else [0.2, 0.4]
        vin = list(c.ViNant)
        orc.ViNant = list(vin)
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos)
        solo = orc.mpc(x0=x, ySP=ysp, w0=[], u0=u, pesos=pesos)
        w = sol['w_opt'].full().reshape(-1)
        scale = max(np.max(np.abs(solo['w_opt'][:N * nu])), 1e-9)
        assert np.max(np.abs(w[:N * nu] - solo['w_opt'][:N * nu])) <= RTOL * scale, k
        assert abs(float(sol['J']) - solo['J']) <= RTOL * max(abs(solo['J']), 1e-9), k
        si = w[N * nu + ny:]
        Ve = 1.0 * si[0] ** 2 + 2.0 * si[1] ** 2
        assert Ve <= vin[0] * (1 + 1e-8) + 1e-300
        lam_e = solo['info'].get('lam_ellipsoid', 0.0)
        if lam_e > 0:
            active += 1
            mine = float(sol['lam_g'].full().reshape(-1)[N * s.nx + N * nu + 2 * ny + c.V.index(c._ell[0]['V'])])
            assert abs(mine - lam_e) <= 1e-5 * lam_e, (k, mine, lam_e)
        assert abs(c.ViNant[0] - orc.ViNant[0]) <= 1e-6 * max(orc.ViNant[0], 1e-12), k
        du = c.du
        res = c.dynF(x0=x, u0=u, du0=du)
        x = res['xkp1'].full().reshape(-1); u = res['ukp1'].full().reshape(-1)
        w0 = c.warmStart(sol, ysp)
        pesos, _ = c.satWeights2(x, u, w0, ysp)
    assert active >= 3
    with pytest.raises(NotImplementedError):
        c.engine(8)


@pytest.mark.parametrize("theta,N", [([[2, 1], [3, 2]], 10),      # nx = 14: other dead times, same horizon
                                     ([[1, 3], [7, 3]], 8),       # Wood & Berry with N = 8 (N*nu = 16)
                                     ([[1, 0], [0, 1]], 6),       # nx = 10, N*nu = 12
                                     ([[5, 9], [10, 6]], 12)])    # nx = 28, N*nu = 24
def test_tensor_fast_path_shape_family(theta, N, monkeypatch):
    """The tensor fast path (mma_step_kernel) is compiled for a family of 2x2 shapes -- any dead times (nx <= 32),
    horizons 6, 8, 10, 12 -- not for the benchmark's one: each member must be SERVED by it, resolve bounds that bind
    in its own resume pass, and match the general kernel and the oracle."""
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False); monkeypatch.delenv("SIHMPC_FAST", raising=False)
    wb = sc.WOOD_BERRY
    sys_ = sc.fopdt_system(wb["K"], wb["tau"], theta, 1.0)
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])
    declare = lambda cls, s, **b: sc.declare_distillation(cls, s, N=N, **(b or bounds))
    B, T = 96, 60
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, size=(B, 2)) * np.array([0.8, 0.4])
    x0 = np.zeros((B, sys_.nx)); x0[:, 0] = 96.0; x0[:, 1] = 0.5
    scen = dict(x0=x0, u0=np.tile([1.95, 1.71], (B, 1)),
                ysp_seq=np.stack([np.tile([96.0, 0.5], (B, 1)), np.array([96.0, 0.5]) + a], axis=1),
                seg_start=np.tile(np.array([0, 4], dtype=np.int32), (B, 1)))

    def run(no_fast):
        if no_fast:
            monkeypatch.setenv("SIHMPC_NO_FAST", "1")
        else:
            monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
        c = declare(IHMPCController, sys_)
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        tr = eng.run(T, record=True)
        return eng, tr["du"].cpu().numpy(), tr["J"].cpu().numpy()

    eng, du, J = run(False)
    assert eng.dp.kernel_name() == "mma_step_kernel"
    handed, solved = eng.dp.counters()
    st = eng.stats.cpu().numpy()
    assert (st[:, 2 + 3] == 0).all() and handed == 0
    engg, dug, Jg = run(True)
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    assert engg.dp.kernel_name() == "ihmpc_step_kernel"
    n_constrained = int((engg.stats.cpu().numpy()[:, 2 + 4] > 0).sum())
    assert n_constrained > 0 and solved > 0, (n_constrained, solved)
    assert np.max(np.abs(du - dug)) <= 1e-7 * max(1.0, np.max(np.abs(dug)))
    worst = int(np.argmax(engg.stats.cpu().numpy()[:, 2 + 4]))
    for idx in (0, worst, B - 1):
        ref = oracle_closed_loop(declare, sys_, scen, idx, T, bounds=bounds)
        assert np.max(np.abs(du[:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"])), idx
        assert np.max(np.abs(J[:, idx] - ref["J"]) / np.maximum(np.abs(ref["J"]), 1e-9)) <= RTOL, idx


def test_resident_controller_stepping_matches_closed_loop():
    """sihmpc_resident_step: the plant with the caller (here NumPy, one mismatched plant per scenario), measured state
    and set-point up, move and cost down, u / weights / warm start resident on the device.  Same numbers as the
    batched engine that simulates the same plants on the device, and as the oracle's driver loop."""
    from sihmpc_b200.engine import ResidentStepper
    sys_ = sc.distillation_system()
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])
    c = sc.declare_distillation(IHMPCController, sys_, **bounds)
    B, T = 128, 45
    scen = sc.mismatch_scenarios(B)
    scen["seg_start"][:, 1] = 5; scen["seg_start"][:, 2] = 25
    plants = [sc.plant_matrices_from_values(sys_, scen["plant_val"][b]) for b in range(B)]
    Ap = np.stack([p[0] for p in plants]); Bp = np.stack([p[1] for p in plants])
    rs = ResidentStepper(c.compile(), B, scen["u0"])
    x = scen["x0"].copy()
    du_r, J_r = [], []
    for k in range(T):
        seg = (scen["seg_start"] <= k).sum(axis=1) - 1
        ysp = scen["ysp_seq"][np.arange(B), seg]
        du, J = rs.step(x, ysp)
        du = du.numpy().copy(); du_r.append(du); J_r.append(J.numpy().copy())
        x = np.einsum('bij,bj->bi', Ap, x) + np.einsum('bij,bj->bi', Bp, du)
    du_r, J_r = np.array(du_r), np.array(J_r)
    assert (rs.status.numpy() == 0).all()
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_plant_values(scen["plant_val"])
    tr = eng.run(T, record=True)
    du_e, J_e = tr["du"].cpu().numpy(), tr["J"].cpu().numpy()
    assert int((eng.stats.cpu().numpy()[:, 2 + 1] >= 0).sum()) == B
    assert np.max(np.abs(du_r - du_e)) <= 1e-9 * max(1.0, np.max(np.abs(du_e)))
    assert np.max(np.abs(J_r - J_e) / np.maximum(np.abs(J_e), 1e-9)) <= 1e-7
    on_bound = (np.abs(scen["u0"][None] + np.cumsum(du_r, axis=0) - np.array(bounds["ulb"])) < 1e-9) | \
               (np.abs(scen["u0"][None] + np.cumsum(du_r, axis=0) - np.array(bounds["uub"])) < 1e-9)
    worst = int(np.argmax(on_bound.sum(axis=(0, 2))))
    assert on_bound[:, worst].sum() > 0
    for idx in (0, worst):
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, bounds=bounds, plant=plants[idx])
        assert np.max(np.abs(du_r[:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"])), idx


def _guarded(t, guard=512):
    """The same tensor between two guard regions of NaN-pattern words; returns (view, whole storage)."""
    import torch
    flat = t.contiguous().reshape(-1)
    if flat.dtype == torch.float64:
        store = torch.full((flat.numel() + 2 * guard,), float("nan"), dtype=torch.float64, device=t.device)
    else:
        store = torch.full((flat.numel() + 2 * guard,), -1431655766, dtype=flat.dtype, device=t.device)
    store[guard:guard + flat.numel()] = flat
    return store[guard:guard + flat.numel()].view(t.shape), store


def _guards_intact(store, n, guard=512):
    import torch
    g = torch.cat([store[:guard], store[guard + n:]])
    if store.dtype == torch.float64:
        return bool(torch.isnan(g).all().item())
    return bool((g == -1431655766).all().item())


@pytest.mark.parametrize("which", ["mma", "tile", "general", "mimo10", "log"])
def test_no_kernel_writes_outside_its_buffers(which, monkeypatch):
    """Stand-in for compute-sanitizer memcheck (closed on this pool, profiles/r02_sanitizer_unavailable.md): every
    array a step call receives sits between guard regions; closed loops with active bounds, hand-over, resume pass and
    an odd batch size must leave every guard word untouched."""
    import torch
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False); monkeypatch.delenv("SIHMPC_FAST", raising=False)
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])
    if which == "mimo10":
        sys_ = sc.mimo10_system()
        c = sc.declare_mimo10(IHMPCController, sys_, ulb=[-0.4] * 10, uub=[0.4] * 10)
        B = 37
        scen = sc.mimo10_scenarios(B, sys_)
    elif which == "log":
        from sihmpc_b200.ihmpccontroller_log import IHMPCController as LogC
        sys_ = sc.distillation_system()
        c = sc.declare_distillation_log(LogC, sys_)
        B = 37
        scen = sc.distillation_scenarios(B)
    else:
        if which == "tile":
            monkeypatch.setenv("SIHMPC_FAST", "tile")
        if which == "general":
            monkeypatch.setenv("SIHMPC_NO_FAST", "1")
        sys_ = sc.distillation_system()
        c = sc.declare_distillation(IHMPCController, sys_, **bounds)
        B = 4099
        scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 3; scen["seg_start"][:, 2] = 12
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(np.minimum(scen["dist_step"], 15), scen["dist"])
    stores = {}
    names = ["x", "u", "du", "J", "w_opt", "status", "stats", "ysp_seq", "seg_start", "dist", "dist_step"]
    names += ["w_warm", "iters", "ss_last"] if which == "log" else ["pesos", "vinant", "n_active"]
    for name in names:
        view, store = _guarded(getattr(eng, name))
        setattr(eng, name, view); stores[name] = (store, view.numel())
    if which != "log":
        ws, nb = eng.dp.workspace(B)
        eng.dp._ws = torch.full((nb + 8192,), 0xAB, dtype=torch.uint8, device=ws.device)[4096:4096 + nb]
        wstore = eng.dp._ws._base if eng.dp._ws._base is not None else None
    tr = eng.run(20, steps_per_launch=7, record=True)
    torch.cuda.synchronize()
    assert np.isfinite(tr["du"].cpu().numpy()).all()
    for name, (store, n) in stores.items():
        assert _guards_intact(store, n), name
    if which != "log" and wstore is not None:
        w = wstore.cpu().numpy()
        assert (w[:4096] == 0xAB).all() and (w[4096 + nb:] == 0xAB).all()
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False); monkeypatch.delenv("SIHMPC_FAST", raising=False)


def test_histories_of_a_scenario_subset_and_result_export(tmp_path):
    """Per-step histories for a SUBSET of a big batch (sihmpc_b200.history.trace_subset) equal what the drop-in class
    records scenario by scenario (j_hist, v_hist, peso_hist: ihmpccontroller.py:592-601) and what the batched engine
    computes for those scenarios inside the full batch; the pickle has the reference's layout and file name
    (example_distillation_log.py:263-269)."""
    import pickle
    from sihmpc_b200.history import trace_subset, export_results
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    B, T = 512, 30
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 4; scen["seg_start"][:, 2] = 15
    picks = [3, 200, 511]
    h = trace_subset(c, scen, picks, T)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_disturbance(scen["dist_step"], scen["dist"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy(); y = tr["y"].cpu().numpy()
    for a, idx in enumerate(picks):
        assert np.max(np.abs(h["du"][a].T - du[:, idx])) <= 1e-12
        assert np.max(np.abs(h["J"][a] - J[:, idx])) <= 1e-12 * max(1.0, np.max(np.abs(J[:, idx])))
        assert np.max(np.abs(h["y"][a][:, 1:].T - y[:, idx])) <= 1e-10
    # the drop-in class on scenario 200, step by step
    idx = 200
    cc = sc.declare_distillation(IHMPCController, sys_)
    x = scen["x0"][idx].copy(); u = scen["u0"][idx].copy(); w0 = []; pesos = []
    f = schedule_fn(scen["ysp_seq"][idx], scen["seg_start"][idx])
    for k in range(T):
        sol = cc.mpc(x0=x, ySP=f(k), w0=w0, u0=u, pesos=pesos)
        res = cc.dynF(x0=x, u0=u, du0=cc.du)
        x = res['xkp1'].full().reshape(-1); u = res['ukp1'].full().reshape(-1)
        if k == int(scen["dist_step"][idx]):
            x[:2] += scen["dist"][idx]
        w0 = cc.warmStart(sol, f(k))
        pesos, _ = cc.satWeights2(x, u, w0, f(k))
    a = picks.index(idx)
    assert np.allclose(h["J"][a], cc.j_hist, rtol=1e-9, atol=1e-14)
    assert np.allclose(h["v"][a].T, np.array([V.v_hist for V in cc.V]).T, rtol=1e-8, atol=1e-12)
    assert np.allclose(h["pesos"][a].T, np.array([V.peso_hist for V in cc.VJ]).T, rtol=1e-8)
    path = export_results(c, h, a, path=str(tmp_path / "out.dat"))
    tup = pickle.load(open(path, "rb"))
    assert len(tup) == 8 and tup[0].shape == (2, T + 1) and tup[1].shape == (2, T) and tup[3].shape == (sys_.nx, T)
    assert tup[4].shape == (len(c.V), T) and tup[5].shape == (len(c.VJ), T) and tup[6].shape == (T,)
    import os
    assert os.path.basename(export_results(c, h, a, path=None)).startswith("dist_10_[")
    os.remove(export_results(c, h, a, path=None))


def test_unstable_mismatch_scenario_is_flagged_not_nan():
    """One scenario in 10^6 of the mismatch Monte-Carlo (global index 843890) is closed-loop UNSTABLE (gain and
    time-constant mismatch; the oracle diverges on it too).  Its weights end up spanning 1e-9 ... 1e6 inside one
    Hessian and a pivot comes out <= 0 in FP64: the step must be FLAGGED (SIHMPC_ST_NOTPD in the statistics) and
    carried on with the shifted factorisation -- finite numbers, never NaN -- while its neighbours stay clean."""
    from sihmpc_b200 import _capi
    c = sc.declare_distillation(IHMPCController)
    first, B, T = 843890 - 3, 8, 200
    scen = sc.mismatch_scenarios(B, first=first)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    eng.set_plant_values(scen["plant_val"])
    eng.run(T)
    st = eng.stats.cpu().numpy(); x = eng.x.cpu().numpy()
    assert np.isfinite(x).all() and np.isfinite(st).all()
    flags = st[:, 2 + 3].astype(int)
    assert np.abs(x[3]).max() > 1e2                     # the unstable one has run away ...
    assert (np.abs(np.delete(x, 3, axis=0)) < 1e2 + 100).all()
    assert (np.delete(flags, 3) == 0).all()             # ... its neighbours are clean
    if flags[3]:
        assert flags[3] & _capi.ST_NOTPD


def test_fast_path_handover_many_rounds(monkeypatch):
    """Quad fast path with a tiny grid (many scenarios per quad, one after the other) on a bound
    set that makes a good share of the scenarios hand over to the exact active-set kernel mid-loop:
    must equal the general kernel run alone, and the oracle.  Regression test for a barrier placed
    inside a divergent branch (a handed-over quad left its warp one barrier out of phase)."""
    sys_ = sc.distillation_system()
    bounds = dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])
    B, T = 1024, 150
    scen = sc.distillation_scenarios(B)

    def run(no_fast, grid):
        monkeypatch.setenv("SIHMPC_NO_FAST", "1" if no_fast else "0")
        monkeypatch.setenv("SIHMPC_NO_INKERNEL_AS", "1")     # this test is about the hand-over itself
        if grid:
            monkeypatch.setenv("SIHMPC_QUAD_GRID", str(grid))
        else:
            monkeypatch.delenv("SIHMPC_QUAD_GRID", raising=False)
        c = sc.declare_distillation(IHMPCController, sys_, **bounds)
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"])
        eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        eng.set_disturbance(scen["dist_step"], scen["dist"])
        tr = eng.run(T, steps_per_launch=50, record=True)
        return tr["du"].cpu().numpy(), eng.x.cpu().numpy(), eng.stats.cpu().numpy()

    du_f, x_f, st_f = run(False, 3)
    du_g, x_g, st_g = run(True, 0)
    monkeypatch.delenv("SIHMPC_QUAD_GRID", raising=False)
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    monkeypatch.delenv("SIHMPC_NO_INKERNEL_AS", raising=False)
    assert (st_f[:, 5] == 0).all() and (st_g[:, 5] == 0).all()
    n_constrained = int((st_g[:, 6] > 0).sum())
    assert 0.05 * B < n_constrained < 0.95 * B          # both paths are exercised
    assert np.max(np.abs(du_f - du_g)) <= 1e-9 * np.max(np.abs(du_g))
    assert np.max(np.abs(x_f - x_g)) <= 1e-9 * 100.0
    assert np.allclose(st_f[:, :5], st_g[:, :5], rtol=1e-9, atol=1e-12)
    idx = int(np.argmax(st_g[:, 6]))
    ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, bounds=bounds)
    assert ref["n_active"].max() > 0
    assert np.max(np.abs(du_f[:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"]))


@pytest.mark.parametrize("bname,bounds", [
    ("input_box", dict(ulb=[1.80, 1.62], uub=[2.02, 1.78])),          # cumulative moves are the simple variables
    ("example_ulb", dict(ulb=[1.9, 1.55])),
    ("move_box", dict(dulb=[-0.01, -0.03], duub=[0.015, 0.02])),      # the moves themselves
])
def test_inkernel_active_set_matches_general_and_oracle(bname, bounds, monkeypatch):
    """Tensor fast path with bounds active: the scenarios are solved INSIDE mma_step_kernel (mma_active_set: the
    bound family restated as simple bounds, fixed variables masked out of the register-resident factorisation, KKT
    certificate), not handed over.  Must equal the general dual-active-set kernel and the oracle."""
    sys_ = sc.distillation_system()
    B, T = 512, 150
    scen = sc.distillation_scenarios(B)

    def run(no_fast):
        monkeypatch.setenv("SIHMPC_NO_FAST", "1" if no_fast else "0")
        c = sc.declare_distillation(IHMPCController, sys_, **bounds)
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"])
        eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        eng.set_disturbance(scen["dist_step"], scen["dist"])
        tr = eng.run(T, record=True)
        return tr["du"].cpu().numpy(), eng.x.cpu().numpy(), eng.stats.cpu().numpy(), eng.dp.counters(), eng.dp.kernel_name()

    du_f, x_f, st_f, cnt_f, name_f = run(False)
    du_g, x_g, st_g, _, name_g = run(True)
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    assert name_f == "mma_step_kernel" and name_g == "ihmpc_step_kernel"
    handed, inkernel = cnt_f
    n_constrained = int((st_g[:, 6] > 0).sum())
    assert n_constrained > 0.05 * B
    assert inkernel > 0 and handed <= 0.02 * B, (handed, inkernel, n_constrained)
    assert (st_f[:, 5] == 0).all() and (st_g[:, 5] == 0).all()
    assert st_f[:, 3].max() <= VIOL
    # two exact methods with different (tiny) feasibility tolerances, 150 closed-loop steps apart
    assert np.max(np.abs(du_f - du_g)) <= 2e-7 * np.max(np.abs(du_g))
    assert np.max(np.abs(x_f - x_g)) <= 2e-7 * 100.0
    assert np.allclose(st_f[:, :3], st_g[:, :3], rtol=1e-6, atol=1e-12)
    worst = int(np.argmax(np.abs(du_f - du_g).max(axis=(0, 2))))
    for idx in sorted({int(np.argmax(st_g[:, 6])), worst, 0}):
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, T, bounds=bounds)
        assert np.max(np.abs(du_f[:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"])), idx


def _siso_integrating():
    from sihmpc_b200.opom import OPOM, TransferFunction
    return OPOM([[TransferFunction([1.0], [5.0, 1.0, 0.0], delay=1)]], 1.0)


def _declare_siso(cls, s, **bounds):
    c = cls(s, 6, **(bounds or dict(dulb=[-0.5], duub=[0.5])))
    c.subObjComposed(y=[0], Q=1, sat=6.0)
    c.subObj(du=[0], Q=1, sat=1.0)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.set_terminal_objective()
    return c


def test_integrating_plant_contraction_rows():
    """Integrating channel: siN != 0, the (k+1-N) Ts siN term and the contraction row
    q siN^2 <= ViN_ant (ihmpccontroller.py:174,265-269,581-584) -- dense rows of the general kernel."""
    s = _siso_integrating()
    c = _declare_siso(IHMPCController, s)
    plan = c.compile()
    assert plan.ncd == 1 and plan.n_vin == 1
    B, T = 40, 60
    rng = np.random.default_rng(3)
    ysp = 1.0 + 0.5 * rng.uniform(-1, 1, size=(B, 1))
    eng = c.engine(B)
    eng.set_state(np.zeros((B, s.nx)), np.zeros((B, 1)))
    eng.set_schedule(ysp[:, None, :], np.zeros((B, 1), dtype=np.int32))
    tr = eng.run(T, steps_per_launch=7, record=True)
    du = tr["du"].cpu().numpy(); y = tr["y"].cpu().numpy()
    assert (eng.stats[:, 1 + 3].cpu().numpy() == 0).all()
    for idx in (0, 17, 39):
        orc = _declare_siso(OracleIHMPC, s)
        from oracle.ihmpc_oracle import closed_loop
        ref = closed_loop(orc, np.zeros(s.nx), [0.0], lambda k: ysp[idx], T)
        assert np.max(np.abs(du[:, idx, 0] - ref["du"][:, 0])) <= RTOL * np.max(np.abs(ref["du"])), idx
        assert np.max(np.abs(y[:, idx, 0] - ref["y"][:, 0])) <= RTOL * max(1.0, np.max(np.abs(ref["y"])))
        assert abs(ref["y"][-1, 0] - ysp[idx, 0]) < 5e-3
    vin = eng.vinant.cpu().numpy()
    assert np.isfinite(vin).all() and (vin >= 0).all()


def test_dense_rows_state_and_slack_boxes():
    """Finite xlb/xub (N*nx rows, ihmpccontroller.py:414-418) and slack boxes (:405-412)."""
    sys_ = sc.distillation_system()
    bounds = dict(ulb=[0, 0], xlb=np.r_[95.7, -np.inf * np.ones(21)], xub=np.r_[96.25, np.inf * np.ones(21)],
                  sylb=[-0.3, -0.3], syub=[0.3, 0.3])
    c = sc.declare_distillation(IHMPCController, sys_, **bounds)
    orc = sc.declare_distillation(OracleIHMPC, sys_, **bounds)
    assert c.compile().ncd == 9 + 2      # the k=0 row of x[0] has a zero normal (dead time >= 1): dropped
    B = 32
    rng = np.random.default_rng(11)
    x, u, ysp, pesos = _random_states(rng, B)
    ysp[:, 0] = 96.0 + rng.uniform(-0.45, 0.45, size=B)
    ysp[B - 1, 0] = 96.7          # |syN| <= 0.3 and xs(N) <= 96.25 cannot both hold: infeasible rows
    eng = c.engine(B)
    eng.set_state(x, u, pesos)
    eng.solve(ysp)
    w = eng.w_opt.cpu().numpy(); J = eng.J.cpu().numpy(); na = eng.n_active.cpu().numpy()
    st = eng.status.cpu().numpy()
    assert (st[:B - 1] == 0).all() and (st[B - 1] & 2)            # SIHMPC_ST_INFEASIBLE, like the oracle
    with pytest.raises(RuntimeError):
        orc.mpc(x[B - 1], ysp[B - 1], [], u[B - 1], pesos[B - 1], ViN_ant=[np.inf, np.inf])
    nact = 0
    for b in range(B - 1):
        sol = orc.mpc(x[b], ysp[b], [], u[b], pesos[b], ViN_ant=[np.inf, np.inf])
        nact += sol["info"]["n_active"] > 0
        scale = max(np.max(np.abs(sol["w_opt"][:20])), 1e-6)
        assert np.max(np.abs(w[b, :20] - sol["w_opt"][:20])) <= RTOL * scale, b
        assert abs(J[b] - sol["J"]) <= RTOL * max(abs(sol["J"]), 1e-9)
    assert nact > B // 4 and (na > 0).sum() > B // 4


@pytest.mark.parametrize("weights,alfa", [("sat1", 0.0), ("fixed", 0.0), ("sat2", 0.5)])
def test_weight_rules(weights, alfa):
    """satWeights (ihmpccontroller.py:607-617), fixed weights and the alfa blend of satWeights2 (:634)."""
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    B, T = 24, 40
    scen = sc.distillation_scenarios(B)
    eng = c.engine(B, weights=weights, alfa=alfa)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy()
    idx = 5
    orc = sc.declare_distillation(OracleIHMPC, sys_)
    f = schedule_fn(scen["ysp_seq"][idx], scen["seg_start"][idx])
    x = scen["x0"][idx].copy(); u = scen["u0"][idx].copy(); pesos = orc.init_pesos(); w0 = []
    for k in range(T):
        ysp = np.asarray(f(k), float)
        sol = orc.mpc(x, ysp, w0, u, pesos)
        d = orc.du.reshape(-1)
        assert np.max(np.abs(du[k, idx] - d)) <= RTOL * max(np.max(np.abs(d)), 1e-4), (k, weights)
        r = orc.dynF(x, u, d); x, u = r["xkp1"], r["ukp1"]
        w0 = orc.warmStart(sol, ysp)
        if weights == "sat1":
            pesos = orc.satWeights(x, u, w0, ysp)
        elif weights == "sat2":
            orc.pesos = np.asarray(pesos, float)
            pesos, _ = orc.satWeights2(x, u, w0, ysp, alfa=alfa)


@pytest.mark.parametrize("B", [1, 7, 33, 1000])
def test_odd_batch_sizes_fast_path_equals_general(B, monkeypatch):
    sys_ = sc.distillation_system()
    scen = sc.distillation_scenarios(B)
    outs = []
    for no_fast in ("0", "1"):
        monkeypatch.setenv("SIHMPC_NO_FAST", no_fast)
        c = sc.declare_distillation(IHMPCController, sys_)
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"])
        eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        eng.set_disturbance(scen["dist_step"], scen["dist"])
        tr = eng.run(30, steps_per_launch=11, record=True)
        outs.append((tr["du"].cpu().numpy(), tr["J"].cpu().numpy(), eng.x.cpu().numpy(), eng.w_opt.cpu().numpy()))
    monkeypatch.delenv("SIHMPC_NO_FAST", raising=False)
    for a, b in zip(outs[0], outs[1]):
        assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b)))


def test_host_buffer_entry_point_matches_device_engine():
    """sihmpc_step_host (small packed path and chunked pipeline) == sihmpc_step on device buffers."""
    import torch
    from sihmpc_b200 import _capi
    from sihmpc_b200.engine import HostStepper
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_)
    plan = c.compile()
    fl = _capi.F_ADVANCE | _capi.F_UPDATE_VIN | _capi.F_WEIGHTS_SAT2
    for B in (3, 5000):
        scen = sc.distillation_scenarios(B)
        ysp = np.ascontiguousarray(scen["ysp_seq"][:, 1, :])
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"])
        eng.ysp.copy_(torch.as_tensor(ysp))
        hs = HostStepper(plan, B)
        hs.x.copy_(torch.as_tensor(scen["x0"])); hs.u.copy_(torch.as_tensor(scen["u0"])); hs.ysp.copy_(torch.as_tensor(ysp))
        for k in range(4):
            hs.step(flags=fl)
            io = eng._io(fl, use_schedule=False)
            import ctypes as C
            ws, nb = eng.dp.workspace(B)
            st = torch.cuda.current_stream(eng.device).cuda_stream
            _capi.check(eng.dp.lib.sihmpc_step(eng.dp.handle, C.byref(io), B, 1, C.c_void_p(ws.data_ptr()), nb,
                                               C.c_void_p(st)), "sihmpc_step")
            torch.cuda.synchronize()
            assert np.array_equal(hs.du.numpy(), eng.du.cpu().numpy())
            assert np.array_equal(hs.x.numpy(), eng.x.cpu().numpy())
            assert np.array_equal(hs.pesos.numpy(), eng.pesos.cpu().numpy())
            assert np.array_equal(hs.J.numpy(), eng.J.cpu().numpy())
            assert np.array_equal(hs.w_opt.numpy(), eng.w_opt.cpu().numpy())


def test_log_variant_example_loop_matches_oracle():
    """BASELINE configs[1]: the loop of example_distillation_log.py:63-135 on the drop-in log class
    (sihmpc.ihmpccontroller_log) vs the log oracle; J within 1e-6 relative, du within 1e-6 of scale."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = declare_distillation_log(LogController, sys_)
    c.init_Qt                                                    # sic: example_distillation_log.py:61
    orc = declare_distillation_log(OracleIHMPCLog, sys_)
    assert [V.name for V in c.V] == [V.name for V in orc.V]
    u = [1.95, 1.71]; x = np.append([96, 0.5], np.zeros(sys_.nx - 2)); ysp = [96, 0.5]
    xo, uo = x.copy(), np.array(u)
    w0 = []; lam_w0 = []; lam_g0 = []; w0o = []
    scale = 0.0
    for k in range(40):
        if k > 5:
            ysp = [96, 1]
        if k > 25:
            ysp = [95.5, 1]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, lam_w0=lam_w0, lam_g0=lam_g0, ViN_ant=[])
        solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
        du = sol['du_opt'][:, 0].full()
        scale = max(scale, np.max(np.abs(solo['du_opt'][:, 0])))
        assert np.max(np.abs(du.reshape(-1) - solo['du_opt'][:, 0])) <= RTOL * max(scale, 1e-3), k
        assert abs(float(sol['J']) - solo['J']) <= RTOL * abs(solo['J']), k
        assert abs(float(sol['w_opt'][-1]) - solo['w_opt'][-1]) <= 1e-7
        v = [float(V.F(x, u, sol['w_opt'][:], ysp)) for V in c.V]
        vo = [orc.F(V, xo, uo, solo['w_opt'], ysp) for V in orc.V]
        assert np.allclose(v, vo, rtol=1e-6, atol=1e-9)
        res = c.dynF(x0=x, du0=du, u0=u)
        x = res['xkp1'].full(); u = res['ukp1'].full()
        r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
        w0 = c.warmStart(sol, ysp); w0o = orc.warmStart(solo, ysp)
        lam_w0 = sol['lam_w']; lam_g0 = sol['lam_g']
        assert np.allclose(c.satWeights(x, u, w0, ysp), orc.satWeights(xo, uo, w0o, ysp), rtol=1e-6)
    assert abs(float(sol['w_opt'][-1]) - np.sqrt(2.0)) < 0.05        # ss* ~ sqrt(m/2) near steady state
    assert len(sol['g'].full()) == 260                                 # SURVEY.md section 8 size table


@pytest.mark.parametrize("N,extra", [(16, False), (10, True)])
def test_log_variant_other_instantiations_match_oracle(N, extra):
    """The log kernel is compiled as <rows per lane, m>: <1,4> serves the reference's example; this covers <2,0>
    (N = 16: 33 Newton variables, two rows per lane, in-shared-memory Crout factorisation) and <1,0> (five
    sub-objectives: the generic loops over m) in closed loop against the log oracle, batched launch."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()

    def declare(cls):
        c = cls(sys_, N, ulb=[0, 0])
        c.subObjComposed(y=[0], Q=1, sat=N * 0.5 ** 2)
        c.subObjComposed(y=[1], Q=1, sat=N * 2.0 ** 2)
        c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2)
        c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
        if extra:
            c.subObj(y=[0], Q=1, sat=N * 0.7 ** 2)
        c.subObj(siN=[0], Q=1, addJ=False)
        c.subObj(siN=[1], Q=1, addJ=False)
        return c
    c = declare(LogController)
    p = c.compile()
    assert (p.nv + 1 > 32) == (N == 16) and p.m == (5 if extra else 4)
    B, T = 8, 14
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 3; scen["seg_start"][:, 2] = 9
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    tr = eng.run(T, steps_per_launch=5, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy()
    assert (eng.stats[:, sys_.ny + 3].cpu().numpy() == 0).all()
    for idx in (0, 7):
        orc = declare(OracleIHMPCLog)
        xo, uo, w0o = scen["x0"][idx].copy(), scen["u0"][idx].copy(), []
        ysp_of = schedule_fn(scen["ysp_seq"][idx], scen["seg_start"][idx])
        scale = 1e-3
        for k in range(T):
            ysp = np.asarray(ysp_of(k), float)
            solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
            scale = max(scale, np.max(np.abs(solo['du_opt'][:, 0])))
            assert np.max(np.abs(du[k, idx] - solo['du_opt'][:, 0])) <= RTOL * scale, (idx, k)
            assert abs(J[k, idx] - solo['J']) <= RTOL * abs(solo['J']), (idx, k)
            r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
            w0o = orc.warmStart(solo, ysp)


def test_log_variant_carried_ss_start_saves_iterations_not_accuracy():
    """The Newton iteration of a step starts at ss of the scenario's last optimum (carried in io.pesos between launches)
    instead of the ss = 1 the reference's warm start writes (log.py:548): same minimiser (moves within 1e-9 of a run
    without the carried value), fewer iterations; one step per launch and many per launch agree bit for bit."""
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = sc.declare_distillation_log(LogController, sys_)
    B, T = 16, 24
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 4; scen["seg_start"][:, 2] = 15
    runs = {}
    for tag, spl, carry in (("one", 1, True), ("many", 8, True), ("reset", 1, False)):
        eng = c.engine(B)
        eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
        du = []
        for k in range(0, T, spl):
            if not carry:
                eng.ss_last.zero_()
            tr = eng.run(spl, record=True)
            du.append(tr["du"].cpu().numpy())
        runs[tag] = (np.concatenate(du), eng.stats[:, sys_.ny + 4].cpu().numpy().copy())
        assert (eng.stats[:, sys_.ny + 3].cpu().numpy() == 0).all()
    assert np.array_equal(runs["one"][0], runs["many"][0])
    scale = np.max(np.abs(runs["reset"][0]))
    assert np.max(np.abs(runs["one"][0] - runs["reset"][0])) <= 1e-8 * scale          # both stop at rounding level of phi
    assert runs["one"][1].sum() <= 0.9 * runs["reset"][1].sum()          # Newton iterations of the whole loop


def test_log_variant_batched_closed_loop_matches_oracle():
    """sihmpc_log_loop: 24 scenarios of the log-variant driver loop (example_distillation_log.py:85-135: mpc,
    first move to the plant, warmStart with ss = 1) in one launch vs the log oracle run scenario by scenario."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = declare_distillation_log(LogController, sys_)
    B, T = 24, 30
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 4; scen["seg_start"][:, 2] = 18
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    tr = eng.run(T, steps_per_launch=11, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy(); y = tr["y"].cpu().numpy()
    assert (eng.stats[:, sys_.ny + 3].cpu().numpy() == 0).all()
    assert (eng.stats[:, sys_.ny + 4].cpu().numpy() > 0).all()          # Newton iterations were counted
    f = schedule_fn
    for idx in (0, 9, 23):
        orc = declare_distillation_log(OracleIHMPCLog, sys_)
        xo, uo, w0o = scen["x0"][idx].copy(), scen["u0"][idx].copy(), []
        ysp_of = f(scen["ysp_seq"][idx], scen["seg_start"][idx])
        scale = 1e-3
        for k in range(T):
            ysp = np.asarray(ysp_of(k), float)
            solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
            scale = max(scale, np.max(np.abs(solo['du_opt'][:, 0])))
            assert np.max(np.abs(du[k, idx] - solo['du_opt'][:, 0])) <= RTOL * scale, (idx, k)
            assert abs(J[k, idx] - solo['J']) <= RTOL * abs(solo['J']), (idx, k)
            r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
            assert np.max(np.abs(y[k, idx] - r['ykp1'])) <= RTOL * 100.0
            w0o = orc.warmStart(solo, ysp)
        assert np.max(np.abs(eng.x[idx].cpu().numpy() - xo)) <= RTOL * 100.0


def _declare_log(cls, s, **b):
    N = 10
    c = cls(s, N, **b)
    c.subObjComposed(y=[0], Q=1, sat=N * 0.5 ** 2)
    c.subObjComposed(y=[1], Q=1, sat=N * 2.0 ** 2)
    c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2)
    c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.subObj(siN=[1], Q=1, addJ=False)
    return c


@pytest.mark.parametrize("bname,bounds", [
    ("input_box", dict(ulb=[1.93, 1.60], uub=[2.2, 1.9])),
    ("move_box", dict(ulb=[0, 0], dulb=[-0.01, -0.01], duub=[0.01, 0.01])),
])
def test_log_variant_active_inequalities_match_oracle(bname, bounds):
    """Log-barrier variant with explicit inequalities of the reference NLP ACTIVE at the optimum (input box
    log.py:420-424, move box :375-389): the kernel's interior-point rounds vs the oracle's KKT-certified constrained
    optimum (SciPy trust-constr + active-set Newton polish), closed loop of example_distillation_log.py:85-135."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = _declare_log(LogController, sys_, **bounds)
    orc = _declare_log(OracleIHMPCLog, sys_, **bounds)
    u = [1.95, 1.71]; x = np.append([96, 0.5], np.zeros(sys_.nx - 2)); ysp = [96, 0.5]
    xo, uo = x.copy(), np.array(u)
    w0 = []; w0o = []
    scale, n_constrained = 0.0, 0
    for k in range(12):
        if k > 1:
            ysp = [96, 1]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, lam_w0=[], lam_g0=[], ViN_ant=[])
        solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
        if min(solo['info']['margins']) <= 1e-8:
            n_constrained += 1
            assert orc.last_kkt['stationarity'] <= 1e-8 and orc.last_kkt['lam_min'] >= -1e-9
            # duals (log.py:489-513; fed back as IPOPT's warm start by example_distillation_log.py:101-103) against
            # the oracle's KKT multipliers, casadi's sign convention
            N_, nu_, nx_ = 10, 2, sys_.nx
            lw = sol['lam_w'].full().reshape(-1); lg = sol['lam_g'].full().reshape(-1)
            assert lw.size == N_ * nu_ + 4 + 1 and lg.size == 260
            big = max(1.0, max(abs(v) for v in orc.last_kkt['duals'].values()))
            for key, val in orc.last_kkt['duals'].items():
                if abs(val) < 1e-6 * big:
                    continue                                   # weakly active rows: the set is not unique
                mine = lw[key[1]] if key[0] == 'dU' else (lg[N_ * nx_ + key[1]] if key[0] == 'U' else
                       (lw[-1] if key[0] == 'ss' else lg[N_ * nx_ + N_ * nu_ + 4 + len(c.V) + key[1]]))
                assert abs(mine - val) <= 1e-4 * big, (k, key, mine, val)
        assert c.status == 0
        du = sol['du_opt'][:, 0].full().reshape(-1)
        scale = max(scale, np.max(np.abs(solo['du_opt'])))
        assert np.max(np.abs(sol['du_opt'].full() - solo['du_opt'])) <= RTOL * max(scale, 1e-3), k
        assert abs(float(sol['J']) - solo['J']) <= RTOL * abs(solo['J']), k
        up = sol['u_pred'].full().reshape(-1, 2)
        assert (up >= np.asarray(bounds.get('ulb', [-np.inf] * 2)) - VIOL).all()
        res = c.dynF(x0=x, du0=du, u0=u)
        x = res['xkp1'].full(); u = res['ukp1'].full()
        r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
        w0 = c.warmStart(sol, ysp); w0o = orc.warmStart(solo, ysp)
    assert n_constrained >= 3


def _declare_siso_log(cls, s, **bounds):
    c = cls(s, 6, **(bounds or dict(dulb=[-0.5], duub=[0.5])))
    c.subObjComposed(y=[0], Q=1, sat=6.0)
    c.subObj(du=[0], Q=1, sat=1.0)
    c.subObj(siN=[0], Q=1, addJ=False)
    return c


def test_log_variant_integrating_plant_contraction_row():
    """Log-barrier variant on an INTEGRATING channel: siN != 0 and the contraction row q siN^2 <= ViN_ant
    (ihmpccontroller_log.py:245-270, row bounds :429-434) with the limit contracting along the loop (ViN_ant=None ->
    self.ViNant, log.py:554-561).  The row is a dense row |siN| <= sqrt(ViN_ant/q) of the kernel's interior-point
    phase; moves, cost and the limit itself against the log oracle, same state in."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    s = _siso_integrating()
    c = _declare_siso_log(LogController, s)
    orc = _declare_siso_log(OracleIHMPCLog, s)
    assert c.compile().ncd == 1 and c.compile().n_vin == 1
    x = np.zeros(s.nx); u = np.zeros(1); w0 = []
    tight = 0
    for k in range(40):
        ysp = [1.0] if k < 20 else [0.3]
        orc.ViNant = list(c.ViNant)
        vin = c.ViNant[0]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, lam_w0=[], lam_g0=[])
        solo = orc.mpc(x, ysp, w0, u, [], [])
        assert c.status == 0, (k, c.status)
        scale = max(np.max(np.abs(solo['du_opt'])), 1e-3)
        assert np.max(np.abs(sol['du_opt'].full() - solo['du_opt'])) <= RTOL * scale, k
        assert abs(float(sol['J']) - solo['J']) <= RTOL * max(abs(solo['J']), 1e-9), k
        siN = float(sol['w_opt'].full().reshape(-1)[-2])
        if np.isfinite(vin):
            assert abs(siN) <= np.sqrt(vin) + VIOL, k              # the row itself, north-star violation bound
            if abs(siN) >= np.sqrt(vin) * (1 - 1e-6):
                tight += 1
        assert abs(np.sqrt(c.ViNant[0]) - np.sqrt(orc.ViNant[0])) <= max(VIOL, 1e-6 * np.sqrt(orc.ViNant[0])), k
        res = c.dynF(x0=x, du0=sol['du_opt'][:, 0].full(), u0=u)
        x = res['xkp1'].full().reshape(-1); u = res['ukp1'].full().reshape(-1)
        w0 = c.warmStart(sol, ysp)
    assert abs(float((s.C @ x)[0]) - 0.3) < 0.1          # on its way to the last set-point (the pinned siN slows it)
    assert tight >= 10         # the row binds on most steps once the limit has contracted


def test_log_variant_ss_bound_active():
    """One weighted objective only: the stationary ss = sqrt(m/2) < 1, so ss >= 1 (log.py:390-392) is active."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    s = sc.fopdt_system([[1.5]], [[6.0]], [[0]], 1.0)

    def declare(cls):
        c = cls(s, 6)
        c.subObjComposed(y=[0], Q=1, sat=6.0)
        return c
    c = declare(LogController); orc = declare(OracleIHMPCLog)
    x = np.zeros(s.nx); u = [0.0]
    for ysp in ([0.3], [0.0]):
        sol = c.mpc(x0=x, ySP=ysp, w0=[], u0=u, lam_w0=[], lam_g0=[], ViN_ant=[])
        solo = orc.mpc(x, ysp, [], np.array(u), [], [], ViN_ant=[])
        assert abs(solo['w_opt'][-1] - 1.0) <= 1e-9 and orc.last_kkt['lam_min'] >= -1e-9
        assert abs(float(sol['w_opt'][-1]) - 1.0) <= 1e-7
        assert np.max(np.abs(sol['du_opt'].full() - solo['du_opt'])) <= RTOL * max(np.max(np.abs(solo['du_opt'])), 1e-3)
        assert abs(float(sol['J']) - solo['J']) <= RTOL * max(abs(solo['J']), 1e-3)


@pytest.mark.parametrize("ranges", [dict(rslb=[-0.1, -0.05], rsub=[0.1, 0.05]),
                                    dict(rslb=[-0.02, 0.0], rsub=[0.02, 0.0]),
                                    dict(rslb=[0.01, -0.05], rsub=[0.1, 0.05], ulb=[1.93, 1.60], uub=[2.2, 1.9])])
def test_log_variant_terminal_residual_ranges_match_oracle(ranges):
    """Log variant with terminal residual RANGES (rslb != rsub, ihmpccontroller_log.py:413-427): the ranged residual is
    a decision variable of the Newton / interior-point kernel after the moves, its range a dense row.  Closed loop
    of example_distillation_log.py:85-135 through the drop-in mpc() against the log oracle (range inactive, range
    binding, range that excludes 0 together with the input box), then the batched device loop."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
    from sihmpc.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = declare_distillation_log(LogController, sys_, **ranges)
    orc = declare_distillation_log(OracleIHMPCLog, sys_, **ranges)
    p = c.compile()
    assert p.n_extra == sum(1 for a, b in zip(ranges['rslb'], ranges['rsub']) if a != b)
    u = [1.95, 1.71]; x = np.append([96, 0.5], np.zeros(sys_.nx - 2)); ysp = [96, 0.5]
    xo, uo = x.copy(), np.array(u)
    w0 = []; w0o = []
    scale, n_bound = 0.0, 0
    lo, hi = np.asarray(ranges['rslb'], float), np.asarray(ranges['rsub'], float)
    for k in range(14):
        if k > 1:
            ysp = [96, 1]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, lam_w0=[], lam_g0=[], ViN_ant=[])
        solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
        assert c.status == 0, (k, c.status)
        scale = max(scale, np.max(np.abs(solo['du_opt'])))
        assert np.max(np.abs(sol['du_opt'].full() - solo['du_opt'])) <= RTOL * max(scale, 1e-3), k
        assert abs(float(sol['J']) - solo['J']) <= RTOL * abs(solo['J']), k
        w = sol['w_opt'].full().reshape(-1)
        assert w.size == 25 and np.max(np.abs(w - solo['w_opt'])) <= RTOL * max(np.max(np.abs(solo['w_opt'])), 1e-3), k
        res2 = sol['g'].full().reshape(-1)[10 * sys_.nx + 20 + 2:10 * sys_.nx + 20 + 4]       # g layout, log.py:394-427
        assert (res2 >= lo - VIOL).all() and (res2 <= hi + VIOL).all(), (k, res2)
        if min(solo['info']['margins']) <= 1e-8:
            n_bound += 1
            assert orc.last_kkt['stationarity'] <= 1e-8 and orc.last_kkt['lam_min'] >= -1e-9
        du = sol['du_opt'][:, 0].full().reshape(-1)
        res = c.dynF(x0=x, du0=du, u0=u)
        x = res['xkp1'].full(); u = res['ukp1'].full()
        r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
        w0 = c.warmStart(sol, ysp); w0o = orc.warmStart(solo, ysp)
    if lo[1] != hi[1]:
        assert n_bound >= 3                                    # |rho_s[1]| = 0.05 binds after the set-point change
    # batched device loop (sihmpc_log_loop) on the same declaration: first moves and costs vs the oracle
    B, T = 8, 12
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 2; scen["seg_start"][:, 2] = 9
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    tr = eng.run(T, steps_per_launch=5, record=True)
    du = tr["du"].cpu().numpy(); J = tr["J"].cpu().numpy()
    assert (eng.stats[:, sys_.ny + 3].cpu().numpy() == 0).all()
    for idx in (0, 5):
        orc = declare_distillation_log(OracleIHMPCLog, sys_, **ranges)
        xo, uo, w0o = scen["x0"][idx].copy(), scen["u0"][idx].copy(), []
        ysp_of = schedule_fn(scen["ysp_seq"][idx], scen["seg_start"][idx])
        sc_ = 1e-3
        for k in range(T):
            ysp = np.asarray(ysp_of(k), float)
            solo = orc.mpc(xo, ysp, w0o, uo, [], [], ViN_ant=[])
            sc_ = max(sc_, np.max(np.abs(solo['du_opt'][:, 0])))
            assert np.max(np.abs(du[k, idx] - solo['du_opt'][:, 0])) <= RTOL * sc_, (idx, k)
            assert abs(J[k, idx] - solo['J']) <= RTOL * abs(solo['J']), (idx, k)
            r = orc.dynF(xo, uo, solo['du_opt'][:, 0]); xo, uo = r['xkp1'], r['ukp1']
            w0o = orc.warmStart(solo, ysp)


@pytest.mark.parametrize("ranges", [dict(rslb=[-0.1, -0.05], rsub=[0.1, 0.05]),
                                    dict(rslb=[-0.02, 0.0], rsub=[0.02, 0.0], ulb=[1.9, 1.6])])
def test_terminal_residual_ranges_match_oracle(ranges):
    """Terminal residual RANGES (rslb != rsub, ihmpccontroller.py:441-449): single steps through the drop-in mpc()
    and a batched closed loop, general kernel with the ranged residuals as extra decision variables, vs the oracle's
    un-condensed problem."""
    sys_ = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, sys_, **ranges)
    orc = sc.declare_distillation(OracleIHMPC, sys_, **ranges)
    rng = np.random.default_rng(2)
    for t in range(4):
        x = np.append([96, 0.5], np.zeros(20)) + rng.normal(size=22) * 0.05
        x[6:8] = 0
        u = [1.95, 1.71]; ysp = np.array([96, 1.0]) + rng.normal(size=2) * 0.3
        pesos = np.array([0.1, 0.4, 4.4, 4.4]) * (1 + rng.random(4))
        sol = c.mpc(x0=x, ySP=ysp, w0=[], u0=u, pesos=pesos, ViN_ant=[np.inf, np.inf])
        solo = orc.mpc(x, ysp, [], np.array(u), pesos, ViN_ant=[np.inf, np.inf])
        w = sol['w_opt'].full().reshape(-1)
        assert w.size == 24
        assert np.max(np.abs(w - solo['w_opt'])) <= RTOL * np.max(np.abs(solo['w_opt'])), t
        assert abs(float(sol['J']) - solo['J']) <= RTOL * max(abs(solo['J']), 1e-6), t
    B, T = 16, 40
    scen = sc.distillation_scenarios(B)
    scen["seg_start"][:, 1] = 3; scen["seg_start"][:, 2] = 20
    eng = c.engine(B)
    assert eng.dp.kernel_name() == "ihmpc_step_kernel"
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    tr = eng.run(T, record=True)
    du = tr["du"].cpu().numpy()
    assert (eng.stats[:, 5].cpu().numpy() == 0).all()
    scen2 = {k: v for k, v in scen.items() if k not in ("dist_step", "dist")}
    for idx in (0, 7):
        ref = oracle_closed_loop(sc.declare_distillation, sys_, scen2, idx, T, bounds=ranges)
        assert np.max(np.abs(du[:, idx] - ref["du"])) <= RTOL * np.max(np.abs(ref["du"])), idx
