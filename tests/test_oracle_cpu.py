"""CPU tests of the oracle itself: structural fixtures of the reference tree, KKT certificates,
an independent scipy cross-check on the UN-reduced problem, and the committed golden vectors."""
import os

import numpy as np
import pytest
from scipy.optimize import minimize

from oracle.ihmpc_oracle import OracleIHMPC, solve_qp_exact, closed_loop  # noqa: F401
from sihmpc_b200 import scenarios as sc
from tests.util import oracle_closed_loop

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TIGHT = dict(ulb=[1.9, 1.6], uub=[2.0, 1.8], dulb=[-0.02, -0.02], duub=[0.02, 0.02])


@pytest.fixture(scope="module")
def sys_():
    return sc.distillation_system()


def test_structural_facts_of_reference_scratch_vectors(sys_):
    """/root/reference/teste.py:66-105 pins nx = 22, nw = 24, state order [xs|xd|xi|z] with
    z_1 = du(k-1) first, and zero moves at dU indices 15,17,18,19."""
    c = sc.declare_distillation(OracleIHMPC, sys_)
    assert sys_.nx == 22 and c.nw == 24 and sys_.nd == 4 and sys_.nz == 14 and sys_.theta_max == 7
    x = np.append([96, 0.5], np.zeros(20))
    sol = c.mpc(x, [96, 1], [], [1.95, 1.71], [])
    w = sol["w_opt"]
    assert np.all(np.abs(w[[15, 17, 18, 19]]) < 1e-14)
    assert np.all(np.abs(np.delete(w[:20], [15, 17, 18, 19])) > 1e-5)
    # z_1 = previous move: after one plant step x[8:10] equals the applied du
    x1 = c.dynF(x, [1.95, 1.71], c.du)["xkp1"]
    assert np.allclose(x1[8:10], c.du.reshape(-1))
    assert len(sol["g"]) == 255                                     # SURVEY.md section 8 size table
    assert [V.name for V in c.V] == ['Vy_[0]', 'VsyN_[0]', 'VyC_[0]', 'Vy_[1]', 'VsyN_[1]', 'VyC_[1]',
                                     'Vdu_[0]', 'Vdu_[1]', 'VsiN_[0]', 'VsiN_[1]', 'Vt']
    assert np.allclose(np.diag(c.Qt), [0.1, 0.4])


def test_survey_self_check_value(sys_):
    c = sc.declare_distillation(OracleIHMPC, sys_)
    x = np.append([96, 0.5], np.zeros(20))
    sol = c.mpc(x, [96, 1], [], [1.95, 1.71], [])
    assert abs(sol["J"] - 0.30935531) < 1e-8
    assert np.allclose(sol["du_opt"][:, 0], [-0.0143766510, -0.0668630750], atol=1e-9)
    assert np.allclose(sol["w_opt"][20:22], [0.0213581761, -0.154677657], atol=1e-9)
    assert sol["info"]["n_active"] == 0


@pytest.mark.parametrize("bounds", [dict(ulb=[0, 0]), TIGHT, dict(dulb=[-0.01, -0.03], duub=[0.015, 0.02])])
def test_kkt_certificate_and_slsqp_crosscheck(sys_, bounds):
    c = sc.declare_distillation(OracleIHMPC, sys_, **bounds)
    rng = np.random.default_rng(5)
    for t in range(4):
        x = np.append([96, 0.5], np.zeros(20)) + rng.normal(size=22) * 0.05
        x[6:8] = 0
        u = np.array([1.95, 1.71]); ysp = np.array([96, 1.0]) + rng.normal(size=2) * 0.3
        pesos = np.array([0.1, 0.4, 4.4, 4.4]) * (1 + rng.random(4))
        vin = [np.inf, np.inf]
        sol = c.mpc(x, ysp, [], u, pesos, ViN_ant=vin)
        kk = c.kkt_report(sol, x, ysp, u, pesos, vin)
        assert kk["stationarity"] < 1e-10 and kk["violation"] < 1e-10 and kk["complementarity"] < 1e-10
        # independent solver on the un-reduced (w, g) problem of ctl.py:374-456
        H, f, c0, G, lo, hi = c._problem(x, u, ysp, pesos, vin)
        cons = []
        for i in range(G.shape[0]):
            if np.isfinite(lo[i]) and lo[i] == hi[i]:
                cons.append({"type": "eq", "fun": lambda w, i=i: G[i] @ w - lo[i], "jac": lambda w, i=i: G[i]})
            else:
                if np.isfinite(lo[i]):
                    cons.append({"type": "ineq", "fun": lambda w, i=i: G[i] @ w - lo[i], "jac": lambda w, i=i: G[i]})
                if np.isfinite(hi[i]):
                    cons.append({"type": "ineq", "fun": lambda w, i=i: hi[i] - G[i] @ w, "jac": lambda w, i=i: -G[i]})
        r = minimize(lambda w: 0.5 * w @ H @ w + f @ w + c0, np.zeros(24), jac=lambda w: H @ w + f,
                     constraints=cons, method="SLSQP", options=dict(ftol=1e-15, maxiter=500))
        gv = G @ r.x
        feas = max(0.0, np.max(lo - gv), np.max(gv - hi))
        if r.success:
            assert abs(r.fun - sol["J_quad"]) < 1e-8 * max(1.0, abs(sol["J_quad"]))
            assert np.max(np.abs(r.x - sol["w_opt"])) < 2e-5
        else:
            # SLSQP stalls on heavily degenerate active sets; its (feasible) point can then only be
            # worse than the certified optimum
            assert feas < 1e-8 and r.fun >= sol["J_quad"] - 1e-9
        assert abs(sol["J"] - sol["J_quad"]) < 1e-10 * max(1.0, abs(sol["J"]))


def test_solver_handles_dependent_and_infeasible_rows():
    H = np.eye(3); f = np.array([-1.0, -1.0, -1.0])
    G = np.array([[1.0, 0, 0], [1.0, 0, 0], [1.0, 1.0, 0]])     # duplicated normal
    w, lam, info = solve_qp_exact(H, f, G, np.array([-np.inf] * 3), np.array([0.5, 0.5, 0.8]))
    assert np.allclose(w, [0.4, 0.4, 1.0])
    with pytest.raises(RuntimeError):
        solve_qp_exact(H, f, np.array([[1.0, 0, 0], [1.0, 0, 0]]), np.array([1.0, -np.inf]), np.array([np.inf, 0.0]))


def test_contraction_row_integrating_plant():
    """An integrating channel exercises siN, ViN_ant and the (k+1-N) Ts siN term (ctl.py:174,265-269)."""
    from sihmpc_b200.opom import OPOM, TransferFunction
    h = [[TransferFunction([1.0], [5.0, 1.0, 0.0], delay=1)]]
    s = OPOM(h, 1.0)
    c = OracleIHMPC(s, 6, dulb=[-0.5], duub=[0.5])
    c.subObjComposed(y=[0], Q=1, sat=6.0)
    c.subObj(du=[0], Q=1, sat=1.0)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.set_terminal_objective()
    out = closed_loop(c, np.zeros(s.nx), [0.0], lambda k: [1.0], 60)
    assert abs(out["y"][-1, 0] - 1.0) < 1e-3
    assert np.all(np.diff(np.abs(out["w_opt"][:, -1])) <= 1e-9)   # |siN| is non-increasing


def test_golden_vectors_reproduce(sys_):
    g = np.load(os.path.join(GOLD, "example_distillation_oracle.npz"))
    from tests.golden.make_golden import example_loop
    r = example_loop(OracleIHMPC, sys_, n_steps=40)
    assert np.allclose(r["du"], g["du"][:40], rtol=0, atol=1e-12)
    assert np.allclose(r["J"], g["J"][:40], rtol=1e-10, atol=1e-14)
    assert np.allclose(g["y"][-1], [95.5, 1.0], atol=1e-3)          # example converges to its set-point
    g3 = np.load(os.path.join(GOLD, "config3_oracle.npz"))
    scen = sc.distillation_scenarios(4)
    a = oracle_closed_loop(sc.declare_distillation, sys_, scen, 2, 60)
    assert np.allclose(a["du"], g3["du_2"][:60], rtol=0, atol=1e-12)
    b = oracle_closed_loop(sc.declare_distillation, sys_, scen, 1, 40, bounds=TIGHT)
    assert np.allclose(b["du"], g3["tight_du_1"][:40], rtol=0, atol=1e-12)
    assert g3["tight_nact_1"].max() > 0


def test_log_oracle_stationarity_and_crosscheck(sys_):
    """Log-barrier variant (ihmpccontroller_log.py): the oracle's point is a stationary point of the
    literal objective Vt + ss^2 - sum log(ss*gamma_i - V_i) evaluated through fObj.F-style calls, an
    independent scipy minimisation agrees, and at steady state ss* = sqrt(m/2)."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
    c = declare_distillation_log(OracleIHMPCLog, sys_)
    assert [V.name for V in c.V][:2] == ['Vt', 'Vss'] and len(c.V) == 12 and len(c.VJ) == 4
    x = np.append([96, 0.5], np.zeros(20)); u = np.array([1.95, 1.71])
    sol = c.mpc(x, [96, 0.5], [], u, [], [], ViN_ant=[])
    assert abs(sol['w_opt'][-1] - np.sqrt(2.0)) < 1e-9 and np.max(np.abs(sol['du_opt'])) < 1e-12
    rng = np.random.default_rng(2)
    x2 = x + rng.normal(size=22) * 0.05; x2[6:8] = 0
    ysp = np.array([96.2, 0.8])
    sol = c.mpc(x2, ysp, [], u, [], [], ViN_ant=[])
    w = sol['w_opt']
    gam = np.array([V.gamma for V in c.VJ])

    def literal(v):                      # objective as written in log.py, slacks from the equalities
        dU, ss = v[:20], v[20]
        wq = np.concatenate([dU, np.zeros(4)])
        xi = OracleIHMPC._xi(c, wq, x2, u, ysp) if False else None
        XN = c.X_pred[-1]
        base = np.zeros(c.nxi_); base[:20] = dU
        base[c.o_x:c.o_x + 22] = x2; base[c.o_u:c.o_u + 2] = u; base[c.o_ysp:c.o_ysp + 2] = ysp; base[c.o_one] = 1
        syN = XN[0:2] @ base - ysp; siN = XN[6:8] @ base
        full = np.concatenate([dU, syN, siN, [ss]])
        Vi = np.array([c.F(V, x2, u, full, ysp) for V in c.VJ])
        d = ss * gam - Vi
        if np.any(d <= 0):
            return np.inf
        return c.F(c.Vt, x2, u, full, ysp) + ss ** 2 - np.sum(np.log(d))

    v = np.concatenate([w[:20], [w[-1]]])
    f0 = literal(v)
    assert abs(f0 - sol['J']) < 1e-10 * max(1.0, abs(f0))
    g = np.array([(literal(v + 1e-6 * e) - literal(v - 1e-6 * e)) / 2e-6 for e in np.eye(21)])
    assert np.max(np.abs(g)) < 1e-6
    r = minimize(literal, v + 0.01 * rng.normal(size=21) * np.r_[np.ones(20) * 0.1, 1.0], method="BFGS",
                 options=dict(gtol=1e-9, maxiter=2000))
    assert r.fun >= f0 - 1e-9 and np.max(np.abs(r.x - v)) < 1e-4


def test_log_oracle_constrained_optimum_is_kkt_certified():
    """Log-variant oracle with explicit inequalities active (input box / move box / ss >= 1): the constrained
    optimum carries its own certificate -- stationarity of the Lagrangian, multipliers of the right sign, primal
    feasibility -- and is not worse than any feasible point tried around it."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog
    s = sc.distillation_system()

    def declare(**b):
        N = 10
        c = OracleIHMPCLog(s, N, **b)
        c.subObjComposed(y=[0], Q=1, sat=N * 0.5 ** 2); c.subObjComposed(y=[1], Q=1, sat=N * 2.0 ** 2)
        c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2); c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
        c.subObj(siN=[0], Q=1, addJ=False); c.subObj(siN=[1], Q=1, addJ=False)
        return c
    x0 = np.zeros(22); x0[0] = 96; x0[1] = 0.5
    u0 = np.array([1.95, 1.71])
    free = declare(ulb=[0, 0]).mpc(x0, [96.0, 1.0], [], u0, [], [], ViN_ant=[])
    for b in (dict(ulb=[1.93, 1.60], uub=[2.2, 1.9]), dict(ulb=[0, 0], dulb=[-0.01, -0.01], duub=[0.01, 0.01])):
        c = declare(**b)
        sol = c.mpc(x0, [96.0, 1.0], [], u0, [], [], ViN_ant=[])
        assert min(sol['info']['margins']) >= -1e-9
        assert min(sol['info']['margins']) <= 1e-9                     # something is active
        assert c.last_kkt['stationarity'] <= 1e-9 and c.last_kkt['lam_min'] >= -1e-10 and c.last_kkt['n_active'] > 0
        assert sol['J'] >= free['J'] - 1e-12                           # constraining cannot improve the cost


def test_log_oracle_terminal_residual_ranges():
    """Log variant with terminal residual RANGES (rslb != rsub, ihmpccontroller_log.py:413-427): the ranged residual
    rho = xs(N) - ysp - syN is a decision variable in [lb, ub].  A range that is wide enough changes nothing against
    a free residual; a tight one binds, carries a KKT certificate, costs more than the wide one and less than the
    equality rho = 0."""
    from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
    s = sc.distillation_system()
    x = np.append([96, 0.5], np.zeros(20)); u = np.array([1.95, 1.71]); ysp = [96.0, 1.0]
    eq = declare_distillation_log(OracleIHMPCLog, s).mpc(x, ysp, [], u, [], [], ViN_ant=[])
    wide = declare_distillation_log(OracleIHMPCLog, s, rslb=[-5, -5], rsub=[5, 5])
    sw = wide.mpc(x, ysp, [], u, [], [], ViN_ant=[])
    assert min(sw['info']['margins']) > 1e-3                            # nothing active: plain Newton optimum
    tight = declare_distillation_log(OracleIHMPCLog, s, rslb=[-0.1, -0.05], rsub=[0.1, 0.05])
    st = tight.mpc(x, ysp, [], u, [], [], ViN_ant=[])
    assert abs(min(st['info']['margins'])) <= 1e-9
    assert tight.last_kkt['stationarity'] <= 1e-9 and tight.last_kkt['lam_min'] >= -1e-10
    assert sw['J'] <= st['J'] + 1e-12 and st['J'] <= eq['J'] + 1e-12
    # the reported residual is inside its range: res2 = xs(N) - ysp - syN
    XN = tight.X_pred[-1]
    base = np.zeros(tight.nxi_); base[:20] = st['w_opt'][:20]
    base[tight.o_x:tight.o_x + 22] = x; base[tight.o_u:tight.o_u + 2] = u; base[tight.o_ysp:tight.o_ysp + 2] = ysp
    base[tight.o_one] = 1
    res2 = XN[0:2] @ base - np.asarray(ysp) - st['w_opt'][20:22]
    assert (res2 >= np.array([-0.1, -0.05]) - 1e-9).all() and (res2 <= np.array([0.1, 0.05]) + 1e-9).all()
    assert np.max(np.abs(np.abs(res2) - np.array([0.1, 0.05]))[1:]) <= 1e-8      # the second one sits on its end


def test_multi_index_siN_ellipsoid_row():
    """siN=[0,1]: ONE quadratic row q0 siN0^2 + q1 siN1^2 <= ViN_ant (ihmpccontroller.py:243-271) on a plant whose
    channels integrate.  The outer multiplier search lands on the boundary when the row is active, the limit
    contracts along the loop, and an independent SLSQP solve of the un-reduced problem WITH the quadratic
    constraint agrees."""
    from tests.util import integrating_2x2, declare_ellipsoid
    s = integrating_2x2()
    c = declare_ellipsoid(OracleIHMPC, s)
    assert len(c.ViNant) == 1 and c.V[[V.name for V in c.V].index('VsiN_[0, 1]')].R.shape[0] == 2
    x = np.zeros(s.nx); u = np.zeros(2); w0 = []; pesos = []
    vin_hist, active, checked = [], 0, 0
    for k in range(30):
        ysp = [1.0, -0.5] if k < 15 else [0.2, 0.4]
        vin = list(c.ViNant)
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos)
        Ve = c.F(c.F_ViN[0], x, u, sol["w_opt"], ysp)
        assert Ve <= vin[0] * (1 + 1e-10) + 1e-300
        lam_e = sol["info"].get("lam_ellipsoid", 0.0)
        if lam_e > 0:
            active += 1
            assert abs(Ve - vin[0]) <= 1e-9 * vin[0]                     # on the boundary
            if checked < 4:
                H, f, c0, G, lo, hi = c._problem(x, u, np.asarray(ysp, float), np.asarray(pesos if len(pesos) else c.init_pesos(), float), vin)
                Rw, r0, W, vmax = c._ellipsoids[0]
                cons = [{"type": "ineq", "fun": lambda w: vmax - (Rw @ w + r0) @ W @ (Rw @ w + r0),
                         "jac": lambda w: -2.0 * Rw.T @ W @ (Rw @ w + r0)}]
                for i in range(G.shape[0]):
                    if np.isfinite(lo[i]) and lo[i] == hi[i]:
                        cons.append({"type": "eq", "fun": lambda w, i=i: G[i] @ w - lo[i], "jac": lambda w, i=i: G[i]})
                    else:
                        if np.isfinite(lo[i]):
                            cons.append({"type": "ineq", "fun": lambda w, i=i: G[i] @ w - lo[i], "jac": lambda w, i=i: G[i]})
                        if np.isfinite(hi[i]):
                            cons.append({"type": "ineq", "fun": lambda w, i=i: hi[i] - G[i] @ w, "jac": lambda w, i=i: -G[i]})
                r = minimize(lambda w: 0.5 * w @ H @ w + f @ w + c0, sol["w_opt"] * 0.9, jac=lambda w: H @ w + f,
                             constraints=cons, method="SLSQP", options=dict(ftol=1e-15, maxiter=800))
                if r.success:
                    assert abs(r.fun - sol["J_quad"]) < 1e-7 * max(1.0, abs(sol["J_quad"]))
                    assert np.max(np.abs(r.x - sol["w_opt"])) < 5e-5
                    checked += 1
        vin_hist.append(c.ViNant[0])
        r_ = c.dynF(x, u, c.du)
        x, u = r_['xkp1'], r_['ukp1']
        w0 = c.warmStart(sol, ysp)
        pesos, _ = c.satWeights2(x, u, w0, ysp)
    assert active >= 3 and checked >= 2, (active, checked)
    assert np.isfinite(vin_hist).all() and vin_hist[-1] < 1e-2 * max(vin_hist)     # the limit dies out with the loop
