"""CPU tests of the host logic: OPOM builder, plan compiler, scenario generators, drop-in surface."""
import numpy as np
import pytest

from oracle.ihmpc_oracle import OracleIHMPC, solve_qp_exact
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController, DM
from sihmpc_b200.opom import OPOM, TransferFunction
from sihmpc_b200.plan import assemble_qp


def test_opom_matches_analytic_step_responses():
    wb = sc.WOOD_BERRY
    s = sc.distillation_system()
    t = np.arange(61)
    for j in range(2):
        y = s.step_response(60, j)
        for i in range(2):
            K, tau, th = wb["K"][i][j], wb["tau"][i][j], wb["theta"][i][j]
            ana = np.where(t >= th, K * (1 - np.exp(-(t - th) / tau)), 0.0)
            assert np.max(np.abs(y[:, i] - ana)) < 1e-12
    g = TransferFunction([2.0], [3.0, 1, 0], delay=2)              # integrating channel
    s2 = OPOM([[g]], 0.5)
    y = s2.step_response(30, 0)[:, 0]
    tt = np.maximum(np.arange(31) * 0.5 - 2, 0)
    assert np.max(np.abs(y - 2 * (tt - 3 * (1 - np.exp(-tt / 3))))) < 1e-12
    g2 = TransferFunction([1.0], np.polymul([2.0, 1], [5.0, 1]), delay=0)   # two real poles
    s3 = OPOM([[g2]], 1.0)
    y = s3.step_response(40, 0)[:, 0]
    tt = np.arange(41.0)
    ana = 1 - (5 * np.exp(-tt / 5) - 2 * np.exp(-tt / 2)) / 3
    assert np.max(np.abs(y - ana)) < 1e-12


def _underdamped_system():
    """2x2 plant with complex pole pairs (under-damped second order), a third-order channel and dead times."""
    from sihmpc_b200.opom import OPOM, TransferFunction
    wn, z = 0.5, 0.3
    h = [[TransferFunction([2.0 * wn * wn], np.polymul([1, 2 * z * wn, wn * wn], [4.0, 1.0]), delay=2),
          TransferFunction([1.5], [3.0, 1.0], delay=0)],
         [TransferFunction([0.7], [1.0, 0.4, 0.29], delay=1),
          TransferFunction([-1.0], [5.0, 1.0], delay=3)]]
    return OPOM(h, 1.0), h


def test_opom_complex_poles_match_continuous_step_responses():
    """Under-damped channels: two real states per conjugate pair, F block diagonal."""
    from scipy import signal
    m, h = _underdamped_system()
    assert m.na == 3 and m.nd == 12 and m.nx == 2 * 2 + 12 + 2 * 3
    assert not np.allclose(m.F, np.diag(np.diag(m.F)))           # genuinely block diagonal
    n = 80
    t = np.arange(0, n + 1) * 1.0
    for j in range(2):
        y = m.step_response(n, j)
        for i in range(2):
            _, yy = signal.step(signal.TransferFunction(h[i][j].num, h[i][j].den), T=t)
            d = int(h[i][j].delay)
            ref = np.concatenate([np.zeros(d), yy[:n + 1 - d]])
            assert np.max(np.abs(y[:, i] - ref)) <= 1e-12


def test_opom_repeated_poles_and_fractional_dead_times():
    """OPOM generality: repeated real poles (Jordan block in F) and dead times that
    are not a multiple of Ts (step-response terms evaluated at the fraction of the sample left after the dead time)
    against the analytic step responses at the sample instants."""
    from sihmpc_b200.opom import OPOM, TransferFunction
    # double pole: K/(tau s + 1)^2 e^{-2s}
    K, tau, th = 2.0, 5.0, 2.0
    o = OPOM([[TransferFunction([K], [tau * tau, 2 * tau, 1.0], delay=th)]], 1.0)
    assert o.na == 2 and o.nx == 2 + 2 + 2
    t = np.maximum(np.arange(41) - th, 0)
    assert np.max(np.abs(o.step_response(40)[:, 0] - K * (1 - np.exp(-t / tau) * (1 + t / tau)))) < 1e-13
    # triple pole
    den = np.polymul(np.polymul([3, 1], [3, 1]), [3, 1])
    o = OPOM([[TransferFunction([1.0], den, delay=1.0)]], 1.0)
    a = np.maximum(np.arange(41) - 1.0, 0) / 3
    assert o.na == 3 and np.max(np.abs(o.step_response(40)[:, 0] - (1 - np.exp(-a) * (1 + a + a * a / 2)))) < 1e-13
    # fractional dead time, first order: first visible sample is ceil(theta/Ts)
    K, tau, th = 1.5, 4.0, 2.5
    o = OPOM([[TransferFunction([K], [tau, 1.0], delay=th)]], 1.0)
    t = np.arange(31) - th
    ya = np.where(t > 0, K * (1 - np.exp(-np.maximum(t, 0) / tau)), 0.0)
    assert o.theta_max == 2 and abs(o.tau_e[0, 0] - 0.5) < 1e-15
    assert np.max(np.abs(o.step_response(30)[:, 0] - ya)) < 1e-13
    # fractional dead time on an integrating channel, Ts = 0.5
    o = OPOM([[TransferFunction([0.3], [2.0, 1.0, 0.0], delay=0.4)]], 0.5)
    t = np.arange(31) * 0.5 - 0.4
    tt = np.maximum(t, 0)
    assert np.max(np.abs(o.step_response(30)[:, 0] - np.where(t > 0, 0.3 * (tt - 2.0 * (1 - np.exp(-tt / 2.0))), 0.0))) < 1e-12
    # repeated complex-conjugate pair (each complex Jordan state as Re/Im), integer and fractional dead times
    import scipy.signal as ss
    den = np.polymul([1, 0.4, 1.0], [1, 0.4, 1.0])
    T = np.linspace(0, 40, 400001)
    _, yy = ss.step(ss.TransferFunction([1.0], den), T=T)
    for delay in (0.0, 2.0, 1.5):
        o = OPOM([[TransferFunction([1.0], den, delay=delay)]], 0.5)
        t = np.arange(61) * 0.5 - delay
        assert o.na == 4
        assert np.max(np.abs(o.step_response(60)[:, 0] - np.where(t > 0, np.interp(np.maximum(t, 0), T, yy), 0.0))) < 1e-9
    # a 2x2 plant mixing both with plain channels compiles into a controller plan whose prediction is the oracle's
    h = [[TransferFunction([12.8], [16.7, 1], delay=1), TransferFunction([-18.9], [21.0, 1], delay=3.3)],
         [TransferFunction([6.6], np.polymul([10.9, 1], [10.9, 1]), delay=7), TransferFunction([-19.4], [14.4, 1], delay=3)]]
    s = OPOM(h, 1.0)
    assert s.nx == 2 * 2 + 2 * 2 * 2 + 7 * 2
    c = sc.declare_distillation(IHMPCController, s)
    orc = sc.declare_distillation(OracleIHMPC, s)
    p = c.compile()
    rng = np.random.default_rng(5)
    x = rng.normal(size=s.nx) * 0.1; x[s.ny + s.nd:2 * s.ny + s.nd] = 0.0
    dU = rng.normal(size=p.nv) * 0.05
    xi = orc._xi(np.concatenate([dU, np.zeros(2 * s.ny)]), x, np.zeros(s.nu), np.zeros(s.ny))
    ypred = np.concatenate([Yk @ xi for Yk in orc.Y_pred])
    assert np.max(np.abs(p.Myx @ x + p.Myu @ dU - ypred)) < 1e-12


def test_plan_with_complex_poles_reproduces_oracle_problem():
    """The compiled plan of a controller on the under-damped plant == the oracle's un-condensed problem."""
    from oracle.ihmpc_oracle import OracleIHMPC
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller import IHMPCController
    m, _ = _underdamped_system()
    declare = lambda cls: sc.declare_mimo10(cls, m, N=6, ulb=[-3.0, -3.0], uub=[3.0, 3.0])
    c, orc = declare(IHMPCController), declare(OracleIHMPC)
    plan = c.compile()
    rng = np.random.default_rng(4)
    x = rng.normal(size=m.nx) * 0.1
    x[2 + m.nd:4 + m.nd] = 0.0
    u = rng.normal(size=2) * 0.1
    ysp = np.array([0.4, -0.3])
    pesos = np.asarray(c.init_pesos(), float) * (1 + rng.random(plan.m))
    sol = orc.mpc(x, ysp, [], u, pesos)
    dU = sol["w_opt"][:plan.nv]
    # condensed cost of the plan at the oracle's optimum == the oracle's cost; gradient vanishes (no bound active)
    e = plan.Px @ x + plan.Py @ ysp + plan.e0
    om = plan.Wrow @ pesos + plan.w0row
    r = plan.E @ dU + e
    J = float(om @ (r * r) + (plan.Wdu @ pesos) @ (dU * dU))
    assert abs(J - sol["J"]) <= 1e-9 * max(1.0, abs(sol["J"]))
    if sol["info"]["n_active"] == 0:
        g = 2 * plan.E.T @ (om * r) + 2 * (plan.Wdu @ pesos) * dU
        assert np.max(np.abs(g)) <= 1e-8 * max(1.0, np.max(np.abs(2 * plan.E.T @ (om * e))))


def test_constructor_asserts_like_reference():
    s = sc.distillation_system()
    with pytest.raises(AssertionError):
        IHMPCController(s, 7)                                      # N must exceed theta_max (ctl.py:22)
    with pytest.raises(AssertionError):
        IHMPCController(object(), 10)                              # ctl.py:15
    c = IHMPCController(s, 10, ulb=[0, 0])
    assert (c.nx, c.nu, c.ny, c.nxs, c.nxd, c.nxi, c.nxz) == (22, 2, 2, 2, 4, 2, 14)
    assert c.du.shape == (2, 1)


@pytest.mark.parametrize("bounds", [dict(ulb=[0, 0]),
                                    dict(ulb=[1.9, 1.6], uub=[2.0, 1.8], dulb=[-0.02, -0.02], duub=[0.02, 0.02]),
                                    dict(xlb=np.r_[95.7, -np.inf * np.ones(21)], xub=np.r_[96.3, np.inf * np.ones(21)],
                                         sylb=[-0.3, -0.3], syub=[0.3, 0.3])])
def test_plan_reproduces_oracle_problem(bounds):
    s = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, s, **bounds)
    orc = sc.declare_distillation(OracleIHMPC, s, **bounds)
    plan = c.compile()
    assert [V.name for V in c.V] == [V.name for V in orc.V]
    assert np.allclose(c.Q_bar, orc.Q_bar) and np.allclose(c.init_pesos(), orc.init_pesos())
    rng = np.random.default_rng(0)
    for t in range(6):
        x = np.append([96, 0.5], np.zeros(20)) + rng.normal(size=22) * 0.05
        x[6:8] = 0
        u = np.array([1.95, 1.71]); ysp = np.array([96, 1.0]) + rng.normal(size=2) * 0.3
        pesos = np.array([0.1, 0.4, 4.4, 4.4]) * (1 + rng.random(4))
        sol = orc.mpc(x, ysp, [], u, pesos, ViN_ant=[np.inf, np.inf])
        H, f, c0, G, lo, hi, *_ = assemble_qp(plan, x, u, ysp, pesos, [np.inf, np.inf])
        v, lam, info = solve_qp_exact(H, f, G, lo, hi)
        J = c0 + f @ v + 0.5 * v @ H @ v
        assert np.max(np.abs(v - sol["w_opt"][:20])) < 1e-10
        assert abs(J - sol["J"]) < 1e-10 * max(1.0, abs(sol["J"]))
        assert info["n_active"] == sol["info"]["n_active"]


@pytest.mark.parametrize("ranges", [dict(rslb=[-0.1, -0.05], rsub=[0.1, 0.05]),
                                    dict(rslb=[-0.02, 0.0], rsub=[0.02, 0.0], ulb=[1.9, 1.6]),
                                    dict(rslb=[0.05, -0.2], rsub=[0.05, 0.3])])
def test_plan_with_terminal_residual_ranges_reproduces_oracle_problem(ranges):
    """Terminal residual RANGES (rslb != rsub, ctl.py:441-449): the ranged residual becomes a decision variable
    after the moves (Plan.n_extra); the condensed QP must have the optimum of the oracle's un-condensed problem,
    with syN = xs(N) - ysp - rho."""
    s = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, s, **ranges)
    orc = sc.declare_distillation(OracleIHMPC, s, **ranges)
    plan = c.compile()
    n_ranged = int(np.sum(np.asarray(ranges["rslb"]) != np.asarray(ranges["rsub"])))
    assert plan.n_extra == n_ranged and plan.E.shape[1] == plan.nv + n_ranged and plan.nv == 20
    rng = np.random.default_rng(1)
    for t in range(5):
        x = np.append([96, 0.5], np.zeros(20)) + rng.normal(size=22) * 0.05
        x[6:8] = 0
        u = np.array([1.95, 1.71]); ysp = np.array([96, 1.0]) + rng.normal(size=2) * 0.3
        pesos = np.array([0.1, 0.4, 4.4, 4.4]) * (1 + rng.random(4))
        sol = orc.mpc(x, ysp, [], u, pesos, ViN_ant=[np.inf, np.inf])
        H, f, c0, G, lo, hi, *_ = assemble_qp(plan, x, u, ysp, pesos, [np.inf, np.inf])
        v, lam, info = solve_qp_exact(H, f, G, lo, hi)
        J = c0 + f @ v + 0.5 * v @ H @ v
        assert np.max(np.abs(v[:20] - sol["w_opt"][:20])) < 1e-9
        assert abs(J - sol["J"]) < 1e-9 * max(1.0, abs(sol["J"]))
        res2 = sol["g"][-len(orc.V) - 2:-len(orc.V)] if "g" in sol else None
        rho = v[20:]
        assert (rho >= np.asarray(ranges["rslb"])[np.asarray(ranges["rslb"]) != np.asarray(ranges["rsub"])] - 1e-9).all()


def test_plan_mimo10_shapes_and_parity():
    s = sc.mimo10_system()
    assert (s.nx, s.nd, s.theta_max) == (160, 100, 4)
    c = sc.declare_mimo10(IHMPCController, s)
    orc = sc.declare_mimo10(OracleIHMPC, s)
    plan = c.compile()
    assert (plan.nv, plan.m, plan.mh, plan.R) == (50, 20, 10, 170)
    rng = np.random.default_rng(1)
    x = rng.normal(size=160) * 0.1; x[110:120] = 0
    u = rng.normal(size=10) * 0.1; ysp = rng.normal(size=10)
    pesos = plan.peso0 * (1 + rng.random(20))
    sol = orc.mpc(x, ysp, [], u, pesos, ViN_ant=[np.inf] * 10)
    H, f, c0, G, lo, hi, *_ = assemble_qp(plan, x, u, ysp, pesos, [np.inf] * 10)
    v, lam, info = solve_qp_exact(H, f, G, lo, hi)
    assert np.max(np.abs(v - sol["w_opt"][:50])) < 1e-9


def test_scenario_generators_are_subset_stable():
    a = sc.distillation_scenarios(10, first=4090)
    b = sc.distillation_scenarios(5000)
    assert np.array_equal(a["ysp_seq"], b["ysp_seq"][4090:4100])
    assert np.array_equal(a["dist_step"], b["dist_step"][4090:4100])
    ms = sc.mismatch_scenarios(4, first=10000)
    nom = sc.distillation_system()
    for k in range(4):
        sysp = sc.fopdt_system(ms["K"][k], ms["tau"][k], sc.WOOD_BERRY["theta"])
        A, B = sc.plant_matrices_from_values(nom, ms["plant_val"][k])
        assert np.max(np.abs(A - sysp.A)) < 1e-12 and np.max(np.abs(B - sysp.B)) < 1e-12


def test_dm_behaves_like_casadi_dm():
    a = DM(np.arange(6.0))
    assert a.shape == (6, 1) and a[2:4].shape == (2, 1) and a[-2:].full().ravel().tolist() == [4.0, 5.0]
    m = DM(np.arange(6.0).reshape(2, 3))
    assert m[:, 0].shape == (2, 1) and m[:, 0].full().ravel().tolist() == [0.0, 3.0]
    assert float(DM(3.5)) == 3.5
    with pytest.raises(TypeError):
        float(a)


def test_unsupported_declarations_fail_loudly():
    s = sc.distillation_system()
    c2 = sc.declare_distillation(IHMPCController, s)
    c2.subObj(siN=[0, 1], Q=np.eye(2), addJ=False)          # accepted: mpc() carries the quadratic row ...
    assert c2._ell and c2.compile().m == len(c2.VJ) + 1      # ... as one more weight of the plan (its multiplier)
    with pytest.raises(NotImplementedError):
        c2.engine(4)                                         # ... the batched closed-loop engine does not


def test_rows_the_moves_cannot_influence_are_recorded_and_checked():
    """silb/siub on a plant without an integrating mode: xi(N) = xi(0) whatever the moves, the rows have a zero
    normal.  They are not sent to the device, but the plan keeps them and reports their violation (the reference's
    IPOPT would call the problem infeasible; ihmpccontroller.py:405-412)."""
    from sihmpc_b200.plan import fixed_row_violation
    s = sc.distillation_system()
    c = sc.declare_distillation(IHMPCController, s, ulb=[0, 0], silb=[-1.0, -1.0], siub=[1.0, 1.0])
    with pytest.warns(UserWarning, match="do not depend on the moves"):
        p = c.compile()
    assert p.ncd == 0 and len(p.fixed_rows) >= 2
    x = np.zeros(s.nx); x[0], x[1] = 96.0, 0.5
    assert fixed_row_violation(p, x, [96.0, 0.5]) <= 0.0
    x[s.ny + s.nd] = 2.5                                   # xi_0 outside [-1, 1]
    assert abs(fixed_row_violation(p, x, [96.0, 0.5]) - 1.5) < 1e-12


def test_opom_markov_parameters_match_an_independent_discretisation():
    """OPOM builder against code it shares NOTHING with: every channel discretised on its own by
    scipy.signal.cont2discrete (zero-order hold) with an explicit delay line, stepped by scipy.signal.dlsim.  The
    OPOM model is driven by MOVES, so its Markov parameters C A^(k-1) B (+ D at k = 0) are the plant's unit STEP
    response; A, B, C, D of the product must reproduce it sample by sample -- first-order + dead time (Wood & Berry),
    an integrating channel, an under-damped pair and a third-order channel."""
    from scipy.signal import cont2discrete, dlsim
    wn, z = 0.5, 0.3
    chans = [[([12.8], [16.7, 1.0], 1), ([-18.9], [21.0, 1.0], 3)],
             [([2.0], [3.0, 1.0, 0.0], 2), ([2.0 * wn * wn], list(np.polymul([1, 2 * z * wn, wn * wn], [4.0, 1.0])), 2)]]
    for Ts in (1.0, 0.5):
        h = [[TransferFunction(num, den, delay=d * Ts) for (num, den, d) in row] for row in chans]
        s = OPOM(h, Ts)
        n = 80
        A, B, C, D = (np.asarray(getattr(s, k), float) for k in "ABCD")
        markov = [D.copy()]
        Ak = np.eye(s.nx)
        for k in range(1, n + 1):
            markov.append(C @ Ak @ B)
            Ak = A @ Ak
        markov = np.array(markov)                                   # [k][i][j]
        for i, row in enumerate(chans):
            for j, (num, den, d) in enumerate(row):
                numd, dend, _ = cont2discrete((num, den), Ts, method="zoh")
                _, ystep = dlsim((numd.reshape(-1), dend, Ts), np.ones(n + 1))
                ref = np.concatenate([np.zeros(d), ystep.reshape(-1)])[:n + 1]      # delay line of d samples
                scale = max(1.0, np.max(np.abs(ref)))
                assert np.max(np.abs(markov[:, i, j] - ref)) <= 1e-9 * scale, (Ts, i, j)
