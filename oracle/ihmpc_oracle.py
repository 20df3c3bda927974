# -*- coding: utf-8 -*-
"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy FP64 restatement of the
reference's weighted satisficing IHMPC step, /root/reference/sihmpc/ihmpccontroller.py.

PARITY UNPINNED: the reference's own solver path is casadi.nlpsol('ipopt') on a model
from the third-party `opom` package; neither casadi (pinned casadi==3.5.1,
requirements.txt:4) nor opom (unpinned git install, README.md:1-3) is present in
/root/reference or importable in this image, and the reference ships no tests or golden
outputs.  This file therefore restates the *problem* the
reference hands to IPOPT, line by line, and solves it exactly; the weighted problem is a
strictly convex QP, so any exact solver must return the point IPOPT
converges to.  Certificates: KKT residuals of the UN-reduced (w, g) problem
(`kkt_report`) and an independent scipy SLSQP cross-check in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (sihmpc_b200/) never does.

Every expression is kept as an affine map of  xi = [w ; x ; u ; ysp ; 1]  with
w = [dU_0..dU_{N-1} ; syN ; siN]  (decision vector layout of ctl.py:399-412), so the
N-step prediction is built by the same recursion the reference uses (ctl.py:346-371)
rather than from closed-form prediction matrices.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_discrete_lyapunov, cho_factor, cho_solve, solve_triangular

INF = np.inf


class QuadForm:
    """V(xi) = (R xi)^T W (R xi): one sub-objective (the reference's fObj, ctl.py:116-149)."""

    def __init__(self, R, W, name=""):
        self.R = np.atleast_2d(np.asarray(R, dtype=float))
        W = np.asarray(W, dtype=float)
        self.W = np.diag(W) if W.ndim == 1 else W
        self.name = name
        self.min = 0.0              # ctl.py:120
        self.max = INF              # ctl.py:121
        self.gamma = INF            # ctl.py:122
        self.peso = 0.0             # ctl.py:128
        self.Vi = []
        self.varType = None
        self.index = None
        self.Q = None
        self.type = "simple"
        self.v_hist = []
        self.peso_hist = []
        self.ViN_slot = None        # index into ViNant when max is the symbol ViN_ant

    def value(self, xi):
        r = self.R @ xi
        return float(r @ self.W @ r)


class OracleIHMPC:
    """Restatement of IHMPCController (weighted), same declaration calls as the reference."""

    # ------------------------------------------------------------------ construction
    def __init__(self, sys, N, **kwargs):                       # ctl.py:12-85
        self.sys = sys
        self.Ts = sys.Ts
        self.N = int(N)
        self.theta_max = sys.theta_max
        assert self.N > self.theta_max                           # ctl.py:22
        self.nx, self.nu, self.ny = sys.nx, sys.nu, sys.ny
        self.nxs, self.nxd, self.nxi, self.nxz = sys.ny, sys.nd, sys.ny, sys.nz
        assert self.nx == self.nxs + self.nxd + self.nxi + self.nxz
        nx, nu, ny = self.nx, self.nu, self.ny
        g = kwargs.get
        f = lambda v: np.asarray(v, dtype=float).reshape(-1)
        self.xlb = f(g('xlb', np.ones(nx) * -INF)); self.xub = f(g('xub', np.ones(nx) * INF))
        self.ulb = f(g('ulb', np.ones(nu) * -INF)); self.uub = f(g('uub', np.ones(nu) * INF))
        self.dulb = f(g('dulb', np.ones(nu) * -INF)); self.duub = f(g('duub', np.ones(nu) * INF))
        self.sylb = f(g('sylb', np.ones(ny) * -INF)); self.syub = f(g('syub', np.ones(ny) * INF))
        self.silb = f(g('silb', np.ones(ny) * -INF)); self.siub = f(g('siub', np.ones(ny) * INF))
        self.rslb = f(g('rslb', np.zeros(ny))); self.rsub = f(g('rsub', np.zeros(ny)))
        self.rilb = f(g('rilb', np.zeros(ny))); self.riub = f(g('riub', np.zeros(ny)))

        # layout of xi = [w ; x ; u ; ysp ; 1]
        self.nw = self.N * nu + 2 * ny
        self.o_w = 0
        self.o_sy = self.N * nu
        self.o_si = self.N * nu + ny
        self.o_x = self.nw
        self.o_u = self.o_x + nx
        self.o_ysp = self.o_u + nu
        self.o_one = self.o_ysp + ny
        self.nxi_ = self.o_one + 1

        self._prediction()
        self.V = []
        self.VJ = []
        self.F_ViN = []
        self.ViNant = []
        self.pesos = np.zeros(0)
        self.j_value = 0.0
        self.j_hist = []
        self.du = np.zeros((nu, 1))
        self._s = []
        self.Vt = None
        self.Q_bar = None

    def _sel(self, off, n):
        E = np.zeros((n, self.nxi_))
        E[np.arange(n), off + np.arange(n)] = 1.0
        return E

    def _prediction(self):                                       # ctl.py:346-371 with ctl.py:93-113
        A, B, C, D = self.sys.A, self.sys.B, self.sys.C, self.sys.D
        N, nu = self.N, self.nu
        Xk = self._sel(self.o_x, self.nx)
        Uk = self._sel(self.o_u, nu)
        self.X_pred, self.Y_pred, self.U_pred, self.dU_pred = [], [], [], []
        for k in range(N):
            dUk = self._sel(self.o_w + k * nu, nu)
            Xk = A @ Xk + B @ dUk                                # ctl.py:106
            Yk = C @ Xk + D @ dUk                                # ctl.py:107
            Uk = Uk + dUk                                        # ctl.py:108
            self.X_pred.append(Xk); self.Y_pred.append(Yk); self.U_pred.append(Uk)
            self.dU_pred.append(dUk)
        self.syN = self._sel(self.o_sy, self.ny)
        self.siN = self._sel(self.o_si, self.ny)
        self.Ysp = self._sel(self.o_ysp, self.ny)

    # ------------------------------------------------------------------ sub-objectives
    def _add(self, Vobj, kwargs, to_j_default=True):
        if 'sat' in kwargs:
            Vobj.gamma = kwargs['sat']                           # ctl.py:183
            Vobj.peso = 1.0 / kwargs['sat']                      # ctl.py:184
        self.V.append(Vobj)
        if 'addJ' not in kwargs or kwargs['addJ'] != False:      # ctl.py:187
            self.VJ.append(Vobj)

    def subObj(self, **kwargs):                                  # ctl.py:151-271
        N, Ts = self.N, self.Ts
        if 'Q' in kwargs:
            Q = kwargs['Q']
            if isinstance(Q, (int, np.integer)):
                Q = [Q]
        else:
            Q = np.eye(self.ny)
        # NB: the reference calls np.diag(Q)[j]; for a 1-D list Q=[q] np.diag builds the
        # 1x1 matrix [[q]] and [j] then returns a ROW; with a single index j=0 this is q.
        # We take the diagonal weights as the flat list of q's, which is the same number.
        qd = np.asarray(Q, dtype=float).reshape(-1) if np.ndim(Q) == 1 else np.diag(np.asarray(Q, dtype=float))

        if 'y' in kwargs:                                        # ctl.py:168-192
            inds = kwargs['y']
            rows, wts = [], []
            for j, ind in enumerate(inds):
                for k in range(N):
                    r = (self.Y_pred[k][ind] - self.Ysp[ind] - self.syN[ind]
                         - (k + 1 - N) * Ts * self.siN[ind])     # ctl.py:174
                    rows.append(r); wts.append(qd[j])
            V = QuadForm(np.array(rows), np.array(wts), 'Vy_' + str(inds))
            V.varType = 'y'; V.index = inds; V.Q = Q
            self._add(V, kwargs)
            return V
        if 'du' in kwargs:                                       # ctl.py:194-217
            inds = kwargs['du']
            rows, wts = [], []
            for j, ind in enumerate(inds):
                for k in range(N):
                    rows.append(self.dU_pred[k][ind]); wts.append(qd[j])   # ctl.py:200
            V = QuadForm(np.array(rows), np.array(wts), 'Vdu_' + str(inds))
            V.varType = 'du'; V.index = inds; V.Q = Q
            self._add(V, kwargs)
            return V
        if 'syN' in kwargs:                                      # ctl.py:220-241
            inds = kwargs['syN']
            rows = [self.syN[ind] for ind in inds]
            wts = [qd[j] for j, _ in enumerate(inds)]            # ctl.py:224
            V = QuadForm(np.array(rows), np.array(wts), 'VsyN_' + str(inds))
            V.varType = 'syN'; V.index = inds; V.Q = Q
            self._add(V, kwargs)
            return V
        if 'siN' in kwargs:                                      # ctl.py:243-271
            inds = kwargs['siN']
            rows = [self.siN[ind] for ind in inds]
            wts = [qd[j] for j, _ in enumerate(inds)]            # ctl.py:247
            V = QuadForm(np.array(rows), np.array(wts), 'VsiN_' + str(inds))
            V.varType = 'siN'; V.index = inds; V.Q = Q
            self._add(V, kwargs)
            self.F_ViN.append(V)                                 # ctl.py:259
            V.min = 0.0
            V.ViN_slot = len(self.ViNant)                        # max = symbol ViN_ant (ctl.py:266-268)
            self.ViNant.append(INF)                              # ctl.py:269
            return V
        raise ValueError("subObj needs one of y=, du=, syN=, siN=")

    def subObjComposed(self, **kwargs):                          # ctl.py:273-305
        inds = kwargs['y']
        Vy = self.subObj(y=inds, Q=kwargs['Q'], addJ=False)      # ctl.py:283
        VyN = self.subObj(syN=inds, Q=kwargs['Q'], addJ=False)   # ctl.py:284
        R = np.vstack([Vy.R, VyN.R])
        W = np.concatenate([np.diag(Vy.W), self.N * np.diag(VyN.W)])   # ctl.py:287
        Vc = QuadForm(R, W, 'VyC_' + str(inds))
        if 'addJ' not in kwargs or kwargs['addJ'] != False:      # ctl.py:289-292
            self.VJ.append(Vc)
        Vc.type = 'composed'; Vc.varType = 'y'; Vc.Vi = [Vy, VyN]; Vc.Q = Vy.Q; Vc.index = inds
        if 'sat' in kwargs:                                      # ctl.py:300-302
            Vc.gamma = kwargs['sat']
            Vc.peso = 1.0 / kwargs['sat']
        self.V.append(Vc)                                        # ctl.py:304
        return Vc

    def set_terminal_objective(self):                            # ctl.py:307-343
        Qt = []
        for V in self.VJ:
            if V.varType == 'y':
                qd = np.asarray(V.Q, dtype=float)
                qd = qd.reshape(-1) if qd.ndim == 1 else np.diag(qd)
                Qt = np.append(Qt, (1.0 / V.gamma) * qd)         # ctl.py:312
        self.Qt = np.diag(Qt)
        Psi, F = self.sys.Psi, self.sys.F
        Q_lyap = F.T @ Psi.T @ self.Qt @ Psi @ F                 # ctl.py:338
        Q_bar = solve_discrete_lyapunov(F.T, Q_lyap)             # ctl.py:339
        self.Q_bar = Q_bar
        XdN = self.X_pred[-1][self.nxs:self.nxs + self.nxd]      # ctl.py:328-329
        Vt = QuadForm(XdN, 0.5 * (Q_bar + Q_bar.T), 'Vt')        # ctl.py:340
        Vt.varType = 'terminal'
        self.Vt = Vt
        self.V.append(Vt)                                        # ctl.py:322

    def init_pesos(self):                                        # ctl.py:87-91
        self.pesos = np.array([V.peso for V in self.VJ], dtype=float)
        return self.pesos

    # ------------------------------------------------------------------ evaluation helpers
    def _xi(self, w, x, u, ysp):
        xi = np.zeros(self.nxi_)
        xi[:self.nw] = np.asarray(w, dtype=float).reshape(-1)
        xi[self.o_x:self.o_x + self.nx] = np.asarray(x, dtype=float).reshape(-1)
        xi[self.o_u:self.o_u + self.nu] = np.asarray(u, dtype=float).reshape(-1)
        xi[self.o_ysp:self.o_ysp + self.ny] = np.asarray(ysp, dtype=float).reshape(-1)
        xi[self.o_one] = 1.0
        return xi

    def F(self, V, x, u, w, ysp):
        """fObj.F(x0,u0,var,ysp) -> Value (ctl.py:124-126)."""
        return V.value(self._xi(w, x, u, ysp))

    def dynF(self, x0, u0, du0):                                 # ctl.py:93-113
        x0 = np.asarray(x0, dtype=float).reshape(-1)
        u0 = np.asarray(u0, dtype=float).reshape(-1)
        du0 = np.asarray(du0, dtype=float).reshape(-1)
        xk = self.sys.A @ x0 + self.sys.B @ du0
        yk = self.sys.C @ xk + self.sys.D @ du0
        return {'xkp1': xk, 'ykp1': yk, 'ukp1': u0 + du0}

    # ------------------------------------------------------------------ the NLP of ctl.py:374-477
    def _problem(self, x, u, ysp, pesos, ViN_ant):
        """Return the QP in w:  0.5 w'Hw + f'w + c,  rows  lo <= Gw <= hi  (+ the literal g map)."""
        nw = self.nw
        par = self._xi(np.zeros(nw), x, u, ysp)                  # xi with w = 0

        def split(R):
            return R[:, :nw], R @ par                            # R xi = Rw w + r0

        H = np.zeros((nw, nw)); f = np.zeros(nw); c = 0.0
        terms = [(p, V) for p, V in zip(pesos, self.VJ)]         # J = sum weight*V  (ctl.py:190,215,292)
        if self.Vt is not None:
            terms.append((1.0, self.Vt))                         # ctl.py:323
        for p, V in terms:
            Rw, r0 = split(V.R)
            H += 2.0 * p * Rw.T @ V.W @ Rw
            f += 2.0 * p * Rw.T @ V.W @ r0
            c += p * r0 @ V.W @ r0

        rows, lo, hi = [], [], []
        self._ellipsoids = []            # (Rw, r0, W, vmax) of every multi-index siN row with a finite limit
        # bounds on w (ctl.py:399-412)
        lbw = np.concatenate([np.tile(self.dulb, self.N), self.sylb, self.silb])
        ubw = np.concatenate([np.tile(self.duub, self.N), self.syub, self.siub])
        Iw = np.eye(nw)
        Gw_rows = [Iw]; r0_rows = [np.zeros(nw)]; lo_rows = [lbw]; hi_rows = [ubw]
        # g rows (ctl.py:414-449)
        for k in range(self.N):
            Rw, r0 = split(self.X_pred[k]); Gw_rows.append(Rw); r0_rows.append(r0)
            lo_rows.append(self.xlb); hi_rows.append(self.xub)
        for k in range(self.N):
            Rw, r0 = split(self.U_pred[k]); Gw_rows.append(Rw); r0_rows.append(r0)
            lo_rows.append(self.ulb); hi_rows.append(self.uub)
        XN = self.X_pred[-1]
        XiN = XN[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        Rw, r0 = split(XiN - self.siN); Gw_rows.append(Rw); r0_rows.append(r0)          # res1
        lo_rows.append(self.rilb); hi_rows.append(self.riub)
        XsN = XN[0:self.nxs]
        Rw, r0 = split(XsN - self.Ysp - self.syN); Gw_rows.append(Rw); r0_rows.append(r0)   # res2
        lo_rows.append(self.rslb); hi_rows.append(self.rsub)
        # sub-objective rows (ctl.py:451-456): 0 <= V_i <= max_i.  V_i >= 0 holds identically.
        for V in self.V:
            if V.ViN_slot is not None:
                vmax = ViN_ant[V.ViN_slot]
                if np.isfinite(vmax):
                    if V.R.shape[0] != 1:
                        # sum_j q_j siN_j^2 <= ViN_ant (ctl.py:246-247,266): ONE convex quadratic row.  With the limit
                        # at 0 it pins every siN_j (linear rows); otherwise it is kept aside for mpc()'s outer search
                        Rw, r0 = split(V.R)
                        if vmax <= 0.0:
                            Gw_rows.append(Rw); r0_rows.append(r0)
                            lo_rows.append(np.zeros(Rw.shape[0])); hi_rows.append(np.zeros(Rw.shape[0]))
                        else:
                            self._ellipsoids.append((Rw, r0, V.W, float(vmax)))
                        continue
                    q = float(V.W[0, 0])
                    if q > 0:
                        bound = np.sqrt(max(vmax, 0.0) / q)
                        Rw, r0 = split(V.R); Gw_rows.append(Rw); r0_rows.append(r0)
                        lo_rows.append(np.array([-bound])); hi_rows.append(np.array([bound]))
            elif np.isfinite(V.max):
                raise NotImplementedError("finite upper limit on a general sub-objective")
        G = np.vstack(Gw_rows)
        r0 = np.concatenate(r0_rows)
        lo = np.concatenate(lo_rows) - r0
        hi = np.concatenate(hi_rows) - r0
        return H, f, c, G, lo, hi

    def g_vector(self, w, x, u, ysp):
        """Literal g of ctl.py:414-456 evaluated at w."""
        xi = self._xi(w, x, u, ysp)
        XN = self.X_pred[-1]
        XiN = XN[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        XsN = XN[0:self.nxs]
        parts = [Xk @ xi for Xk in self.X_pred] + [Uk @ xi for Uk in self.U_pred]
        parts += [(XiN - self.siN) @ xi, (XsN - self.Ysp - self.syN) @ xi]
        parts += [np.array([V.value(xi)]) for V in self.V]
        return np.concatenate(parts)

    # ------------------------------------------------------------------ step (ctl.py:567-604)
    def mpc(self, x0, ySP, w0, u0, pesos, lam_w0=[], lam_g0=[], ViN_ant=None):
        if ViN_ant is None:                                      # ctl.py:571-572
            ViN_ant = self.ViNant
        elif len(ViN_ant) == 0:                                  # casadi: empty arg = zeros
            ViN_ant = [0.0] * len(self.ViNant)
        if len(pesos) != len(self.VJ):                           # ctl.py:575-577
            pesos = self.init_pesos()
        pesos = np.asarray(pesos, dtype=float).reshape(-1)
        x0 = np.asarray(x0, dtype=float).reshape(-1)
        u0 = np.asarray(u0, dtype=float).reshape(-1)
        ySP = np.asarray(ySP, dtype=float).reshape(-1)

        H, f, c, G, lo, hi = self._problem(x0, u0, ySP, pesos, ViN_ant)
        w, lam_rows, info = solve_qp_exact(H, f, G, lo, hi)
        if self._ellipsoids:
            w, lam_rows, info = self._solve_with_ellipsoid(H, f, G, lo, hi, w, lam_rows, info)
        nw, N, nu, ny = self.nw, self.N, self.nu, self.ny
        xi = self._xi(w, x0, u0, ySP)
        J = c + f @ w + 0.5 * w @ H @ w
        Jlit = sum(p * V.value(xi) for p, V in zip(pesos, self.VJ)) + (self.Vt.value(xi) if self.Vt else 0.0)
        sol = {
            'J': float(Jlit), 'J_quad': float(J),
            'du_opt': w[:N * nu].reshape(N, nu).T.copy(),       # nu x N  (ctl.py:510-515)
            'w_opt': w.copy(),
            'lam_rows': lam_rows,
            'g': self.g_vector(w, x0, u0, ySP),
            'x_pred': np.concatenate([Xk @ xi for Xk in self.X_pred]),
            'y_pred': np.concatenate([Yk @ xi for Yk in self.Y_pred]),
            'u_pred': np.concatenate([Uk @ xi for Uk in self.U_pred]),
            'info': info,
        }
        warm = self.warmStart(sol, ySP)                          # ctl.py:581-584
        for i, V in enumerate(self.F_ViN):
            self.ViNant[i] = self.F(V, [0] * self.nx, [0] * self.nu, warm, [0] * self.ny)
        self.du = sol['du_opt'][:, 0].reshape(nu, 1).copy()      # ctl.py:595
        self.j_value = sol['J']
        self.j_hist.append(self.j_value)
        for V in self.V:
            V.v_hist.append(self.F(V, x0, u0, w, ySP))           # ctl.py:598-599
        for i, V in enumerate(self.VJ):
            V.peso_hist.append(pesos[i])
        return sol

    def _solve_with_ellipsoid(self, H, f, G, lo, hi, w, lam_rows, info):
        """The QP plus ONE convex quadratic row  V_e(w) = (Rw w + r0)' W (Rw w + r0) <= vmax  (multi-index siN,
        ctl.py:243-271).  KKT: grad J + lam_e grad V_e + G' lam = 0, lam_e >= 0, lam_e (V_e - vmax) = 0.  For fixed
        lam_e the rest is the strictly convex QP of J + lam_e V_e, whose V_e decreases monotonically in lam_e:
        bracket, then Brent on phi(lam_e) = V_e - vmax down to rounding, keeping the feasible end."""
        from scipy.optimize import brentq
        if len(self._ellipsoids) > 1:
            raise NotImplementedError("more than one multi-index siN row with a finite limit")
        Rw, r0, W, vmax = self._ellipsoids[0]
        He, fe = 2.0 * Rw.T @ W @ Rw, 2.0 * Rw.T @ W @ r0

        def val(wv):
            r = Rw @ wv + r0
            return float(r @ W @ r)

        if val(w) <= vmax * (1.0 + 1e-12):
            info = dict(info, lam_ellipsoid=0.0)
            return w, lam_rows, info
        cache = {}

        def phi(lam):
            ww, ll, ii = solve_qp_exact(H + lam * He, f + lam * fe, G, lo, hi)
            cache[lam] = (ww, ll, ii)
            return val(ww) - vmax

        hi_l = 1.0
        while phi(hi_l) > 0.0:
            hi_l *= 8.0
            if hi_l > 1e16:
                raise RuntimeError("ellipsoid row cannot be met (infeasible)")
        lo_l = 0.0 if hi_l == 1.0 else hi_l / 8.0
        lam = brentq(phi, lo_l, hi_l, xtol=1e-300, rtol=8.9e-16, maxiter=200)
        feas = [l for l in cache if val(cache[l][0]) <= vmax * (1.0 + 1e-12)]
        lam = min(feas, key=lambda l: abs(l - lam))              # the feasible end of the last bracket
        ww, ll, ii = cache[lam]
        return ww, ll, dict(ii, lam_ellipsoid=float(lam))

    def warmStart(self, sol, ysp):                               # ctl.py:534-564
        w0 = np.asarray(sol['w_opt']).reshape(-1)
        ny, nu, nx = self.ny, self.nu, self.nx
        dustart = np.concatenate([w0[:-2 * ny][nu:], np.zeros(nu)])
        xN = np.asarray(sol['x_pred']).reshape(-1)[-nx:]
        Xknext = self.sys.A @ xN                                 # dynF with du = 0 (ctl.py:544)
        siNnext = Xknext[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        syNnext = w0[-2 * ny:-ny]                                # "Opcao 1" (ctl.py:558-560)
        return np.concatenate([dustart, syNnext, siNnext])

    def satWeights2(self, x, u, w_start, ysp, alfa=0):           # ctl.py:619-636
        l = len(self.VJ)
        s = np.zeros(l); gamma = np.zeros(l); Vnext = np.zeros(l)
        for i in range(l):
            Vnext[i] = self.F(self.VJ[i], x, u, w_start, ysp)
            gamma[i] = self.VJ[i].gamma
            s[i] = Vnext[i] / gamma[i]
        smax = max(s)
        gamma = np.maximum(gamma, gamma * smax + 1e-6)
        new_pesos = 1.0 / (gamma - Vnext)
        old = self.pesos if len(self.pesos) == l else np.zeros(l)
        self.pesos = alfa * old + (1 - alfa) * new_pesos
        self._s = s
        return self.pesos, gamma

    def satWeights(self, x, u, w_start, ysp):                    # ctl.py:607-617
        pesos = []
        for V in self.VJ:
            Vnext = self.F(V, x, u, w_start, ysp)
            pesos.append(1.0 / (V.gamma - np.clip(Vnext, 0, 0.99 * V.gamma)))
        return np.array(pesos)

    # ------------------------------------------------------------------ certificate
    def kkt_report(self, sol, x0, ySP, u0, pesos, ViN_ant):
        """KKT residuals of the un-reduced problem at sol['w_opt'] (stationarity via NNLS-free LSQ)."""
        H, f, c, G, lo, hi = self._problem(np.asarray(x0, float).reshape(-1), np.asarray(u0, float).reshape(-1),
                                           np.asarray(ySP, float).reshape(-1), np.asarray(pesos, float), ViN_ant)
        w = sol['w_opt']
        grad = H @ w + f
        gv = G @ w
        viol = max(0.0, float(np.max(lo - gv)), float(np.max(gv - hi)))
        lam = sol['lam_rows']
        stat = grad + G.T @ lam
        # sign & complementarity: lam>0 only at upper, lam<0 only at lower (casadi convention)
        eq = np.isfinite(lo) & (lo == hi)
        comp = 0.0
        for i in np.nonzero(~eq)[0]:
            if lam[i] > 0:
                comp = max(comp, abs(lam[i] * (hi[i] - gv[i])) if np.isfinite(hi[i]) else INF)
            elif lam[i] < 0:
                comp = max(comp, abs(lam[i] * (gv[i] - lo[i])) if np.isfinite(lo[i]) else INF)
        return {'stationarity': float(np.max(np.abs(stat))), 'violation': viol, 'complementarity': comp}


# ---------------------------------------------------------------------- exact QP solve
def solve_qp_exact(H, f, G, lo, hi, tol=1e-11):
    """min 0.5 w'Hw + f'w  s.t. lo <= G w <= hi, rows with lo == hi are equalities.

    Equalities are removed with an orthonormal null-space basis (SVD); the remaining
    strictly convex inequality QP is solved with a dual active-set iteration whose linear
    algebra is LAPACK (Cholesky of the reduced Hessian, least squares on the active
    normals).  Returns (w, multipliers per row in casadi sign convention, info).
    """
    n = H.shape[0]
    finite_lo, finite_hi = np.isfinite(lo), np.isfinite(hi)
    eq = finite_lo & finite_hi & (lo == hi)
    Aeq, beq = G[eq], lo[eq]
    if Aeq.shape[0]:
        U, s, Vt = np.linalg.svd(Aeq, full_matrices=True)
        rank = int(np.sum(s > 1e-12 * max(1.0, s[0])))
        wp = Vt[:rank].T @ ((U[:, :rank].T @ beq) / s[:rank])
        if np.max(np.abs(Aeq @ wp - beq)) > 1e-9 * (1 + np.max(np.abs(beq))):
            raise RuntimeError("inconsistent equality constraints")
        Z = Vt[rank:].T
    else:
        wp = np.zeros(n); Z = np.eye(n)
    Hr = Z.T @ H @ Z
    Hr = 0.5 * (Hr + Hr.T)
    fr = Z.T @ (H @ wp + f)
    ineq = np.nonzero(~eq & (finite_lo | finite_hi))[0]
    Gi = G[ineq] @ Z
    gi0 = G[ineq] @ wp
    # candidate constraints  n'v >= b
    Ns, bs, src, sgn = [], [], [], []
    for k, r in enumerate(ineq):
        nrm = np.linalg.norm(Gi[k])
        if finite_lo[r]:
            Ns.append(Gi[k]); bs.append(lo[r] - gi0[k]); src.append(r); sgn.append(-1.0)
        if finite_hi[r]:
            Ns.append(-Gi[k]); bs.append(-(hi[r] - gi0[k])); src.append(r); sgn.append(+1.0)
    Ns = np.array(Ns).reshape(-1, Z.shape[1]); bs = np.array(bs)
    norms = np.linalg.norm(Ns, axis=1) if len(bs) else np.zeros(0)
    cf = cho_factor(Hr, lower=True)
    L = np.tril(cf[0])
    v = -cho_solve(cf, fr)
    act, lam = [], []
    iters = 0
    while len(bs):
        viol = bs - Ns @ v
        scale = 1.0 + np.abs(bs)
        cand = viol / scale
        cand[act] = -INF
        cand[norms <= 1e-13] = -INF          # vacuous rows (zero normal)
        p = int(np.argmax(cand))
        if cand[p] <= tol:
            break
        lam_p = 0.0
        npv = Ns[p]
        while True:
            iters += 1
            if iters > 50 * (len(bs) + 1):
                raise RuntimeError("dual active set did not terminate")
            y = solve_triangular(L, npv, lower=True)
            if act:
                Y = solve_triangular(L, Ns[act].T, lower=True)
                r, *_ = np.linalg.lstsq(Y, y, rcond=None)
                d = y - Y @ r
            else:
                r = np.zeros(0); d = y
            zn = float(d @ d)
            t1, kdrop = INF, -1
            for a in range(len(act)):
                if r[a] > 1e-14:
                    ta = lam[a] / r[a]
                    if ta < t1:
                        t1, kdrop = ta, a
            if zn <= 1e-22 * max(1.0, float(y @ y)) or zn <= 1e-14 * float(y @ y):
                if not np.isfinite(t1):
                    raise RuntimeError("QP infeasible")
                lam = [la - t1 * ra for la, ra in zip(lam, r)]
                lam_p += t1
                del act[kdrop]; del lam[kdrop]
                continue
            z = solve_triangular(L.T, d, lower=False)
            t2 = float(bs[p] - npv @ v) / zn
            t = min(t1, t2)
            v = v + t * z
            lam = [la - t * ra for la, ra in zip(lam, r)]
            lam_p += t
            if t2 <= t1:
                act.append(p); lam.append(lam_p)
                break
            del act[kdrop]; del lam[kdrop]
    w = wp + Z @ v
    lam_rows = np.zeros(G.shape[0])
    for a, la in zip(act, lam):
        lam_rows[src[a]] += sgn[a] * la
    # equality multipliers from stationarity of the full problem
    if Aeq.shape[0]:
        resid = H @ w + f + G.T @ lam_rows
        mu, *_ = np.linalg.lstsq(Aeq.T, -resid, rcond=None)
        lam_rows[np.nonzero(eq)[0]] = mu
    return w, lam_rows, {'n_active': len(act), 'iters': iters,
                         'active_rows': [int(src[a]) for a in act]}


# ---------------------------------------------------------------------- closed loop (ex.py:65-113)
def closed_loop(ctrl, x0, u0, ysp_of_k, n_steps, plant=None, disturbance=None, record=True):
    """The driver loop of example_distillation.py:81-113 around an OracleIHMPC.

    plant: optional (A_p, B_p) for plant-model mismatch (state fed back in full, ex.py:92,102-103).
    disturbance: optional (k_d, d) -- d is added to xs after the plant update of step k_d.
    """
    x = np.asarray(x0, dtype=float).reshape(-1).copy()
    u = np.asarray(u0, dtype=float).reshape(-1).copy()
    A_p, B_p = (ctrl.sys.A, ctrl.sys.B) if plant is None else plant
    C, D = ctrl.sys.C, ctrl.sys.D
    w0 = []
    pesos = []
    out = {'du': [], 'x': [], 'u': [], 'y': [], 'J': [], 'pesos': [], 'w_opt': [], 'n_active': []}
    for k in range(n_steps):
        ysp = np.asarray(ysp_of_k(k), dtype=float)
        sol = ctrl.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos)
        du = ctrl.du.reshape(-1)
        x = A_p @ x + B_p @ du
        if disturbance is not None and k == disturbance[0]:
            x[:ctrl.ny] += np.asarray(disturbance[1], dtype=float)
        y = C @ x + D @ du
        u = u + du
        w0 = ctrl.warmStart(sol, ysp)
        pesos, _ = ctrl.satWeights2(x, u, w0, ysp)
        if record:
            out['du'].append(du.copy()); out['x'].append(x.copy()); out['u'].append(u.copy())
            out['y'].append(y.copy()); out['J'].append(sol['J']); out['pesos'].append(np.array(pesos))
            out['w_opt'].append(sol['w_opt'].copy()); out['n_active'].append(sol['info']['n_active'])
    return {k_: np.array(v) for k_, v in out.items()}
