# -*- coding: utf-8 -*-
"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy FP64 restatement of the
log-barrier satisficing IHMPC step, /root/reference/sihmpc/ihmpccontroller_log.py.

PARITY UNPINNED (same reasons as oracle/ihmpc_oracle.py: casadi/opom absent, no reference tests).

Problem handed to IPOPT by the reference (log.py:349-461):
    w = [dU ; syN ; siN ; ss]                                  (log.py:375-392, ss in [1, inf))
    J = Vt + ss^2 - sum_{i in VJ} log(ss*gamma_i - V_i)        (log.py:71-92,196,219,241,262,292)
    g = [X_pred ; U_pred ; xi(N)-siN ; xs(N)-ysp-syN ; V_l for all l ; ss*gamma_i - V_i >= 0.01]
                                                               (log.py:394-440)
This oracle eliminates the terminal equalities like the weighted one, runs a damped Newton method
on (dU, ss) for the smooth convex objective and then checks whether every explicit inequality
(ss >= 1, ss*gamma_i - V_i >= 0.01, input/move boxes) is strictly inactive at the minimiser --
which is the case for the reference's own example (stationary ss = sqrt(m/2)).
If one is active the constrained problem is solved with SciPy's trust-constr and polished by Newton steps on the
identified active set (_solve_constrained; KKT stationarity kept in last_kkt).
The sub-objective forms, prediction recursion and V ordering come from OracleIHMPC.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_discrete_lyapunov

from oracle.ihmpc_oracle import OracleIHMPC, QuadForm, INF


class OracleIHMPCLog(OracleIHMPC):
    def __init__(self, sys, N, **kwargs):                        # log.py:11-92
        super().__init__(sys, N, **kwargs)
        self.Qt = np.asarray(kwargs.get('Qt', np.eye(sys.ny)), dtype=float)      # log.py:55
        self.nw = self.N * self.nu + 2 * self.ny + 1              # + ss (log.py:390)
        self._terminal()
        self.Vss = QuadForm(np.zeros((1, self.nxi_)), np.zeros(1), 'Vss')
        self.V = [self.Vt, self.Vss]                              # log.py:83
        self.VJ = []

    def _terminal(self):                                          # log.py:304-317
        F, Psi = self.sys.F, self.sys.Psi
        Q_lyap = F.T @ Psi.T @ self.Qt @ Psi @ F
        Q_bar = solve_discrete_lyapunov(F, Q_lyap, method='bilinear')           # log.py:314 (F, not F.T)
        self.Q_bar = Q_bar
        XdN = self.X_pred[-1][self.nxs:self.nxs + self.nxd]
        self.Vt = QuadForm(XdN, 0.5 * (Q_bar + Q_bar.T), 'Vt')
        self.Vt.varType = 'terminal'

    def set_terminal_objective(self):                             # not part of the log surface
        raise AttributeError("the log controller builds its terminal cost in __init__ (log.py:71-75)")

    def init_Qt(self):                                            # log.py:94-101
        Qt = []
        for V in self.VJ:
            if V.varType == 'y':
                qd = np.asarray(V.Q, dtype=float)
                qd = qd.reshape(-1) if qd.ndim == 1 else np.diag(qd)
                Qt = np.append(Qt, (1.0 / V.gamma) * qd)
        self.Qt = np.diag(Qt)
        # the reference only overwrites V[0].V here, not J nor V[0].F (log.py:100-101):
        # the optimisation problem keeps the terminal cost built in __init__.

    # ---- evaluation: var = [dU ; syN ; siN ; ss]
    def F(self, V, x, u, w, ysp):
        w = np.asarray(w, dtype=float).reshape(-1)
        if V is self.Vss:
            return float(w[-1] ** 2)
        return V.value(self._xi(w[:-1], x, u, ysp))

    def _xi(self, w, x, u, ysp):
        w = np.asarray(w, dtype=float).reshape(-1)
        if w.size == self.nw:
            w = w[:-1]
        saved, self.nw = self.nw, self.nw - 1
        try:
            return super()._xi(w, x, u, ysp)
        finally:
            self.nw = saved

    def mpc(self, x0, ySP, w0, u0, lam_w0=None, lam_g0=None, ViN_ant=None):      # log.py:552-563
        if ViN_ant is None:
            ViN_ant = self.ViNant
        elif len(ViN_ant) == 0:                                   # casadi: empty argument = zeros
            ViN_ant = [0.0] * len(self.ViNant)
        x0 = np.asarray(x0, float).reshape(-1); u0 = np.asarray(u0, float).reshape(-1)
        ySP = np.asarray(ySP, float).reshape(-1)
        nv, ny, N, nu = self.N * self.nu, self.ny, self.N, self.nu
        nwq = nv + 2 * ny
        # affine maps in dU after eliminating syN = xs(N)-ysp-rs, siN = xi(N)-ri
        saved, self.nw = self.nw, nwq
        try:
            xi0 = OracleIHMPC._xi(self, np.zeros(nwq), x0, u0, ySP)
        finally:
            self.nw = saved
        XN = self.X_pred[-1]
        XsN = XN[0:self.nxs]; XiN = XN[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        # w(z) = T z + t0 with z = [dU ; rho]: a terminal residual held as an equality (rs/ri lb == ub, the default) is
        # eliminated, syN = xs(N) - ysp - rs, siN = xi(N) - ri; a RANGED one (lb != ub, log.py:413-427) stays a decision
        # variable rho in [lb, ub] after the moves: syN = xs(N) - ysp - rho_s, siN = xi(N) - rho_i
        ext = [('s', i) for i in range(ny) if self.rslb[i] != self.rsub[i]] + \
              [('i', i) for i in range(ny) if self.rilb[i] != self.riub[i]]
        ne = len(ext)
        nz = nv + ne
        T = np.zeros((nwq, nz)); t0 = np.zeros(nwq)
        T[:nv, :nv] = np.eye(nv)
        T[nv:nv + ny, :nv] = XsN[:, :nv]; t0[nv:nv + ny] = XsN @ xi0 - ySP - self.rslb
        T[nv + ny:, :nv] = XiN[:, :nv]; t0[nv + ny:] = XiN @ xi0 - self.rilb
        for c_, (kind, i) in enumerate(ext):
            row = nv + i if kind == 's' else nv + ny + i
            T[row, nv + c_] = -1.0
            t0[row] += self.rslb[i] if kind == 's' else self.rilb[i]     # no constant offset: rho is the variable
        # contraction rows q siN^2 <= ViN_ant (log.py:245-270, row bounds :429-434): |siN| <= sqrt(ViN_ant / q) is linear
        # in dU on an integrating plant; where siN cannot be moved (non-integrating: xi(N) = xi(0)) it must hold as is
        self._extra_lin = []                                   # (row over dU, lo, hi)
        for V in self.V:
            if V.ViN_slot is not None and np.isfinite(ViN_ant[V.ViN_slot]):
                if len(V.index) != 1:
                    raise NotImplementedError("multi-index siN row in the log variant")
                k = nv + ny + V.index[0]
                bound = np.sqrt(max(ViN_ant[V.ViN_slot], 0.0) / float(V.W[0, 0]))
                if np.linalg.norm(T[k]) <= 1e-13:
                    if abs(t0[k]) > bound + 1e-12:
                        raise RuntimeError("siN contraction row violated and out of reach of the moves")
                else:
                    self._extra_lin.append((T[k].copy(), -bound - t0[k], bound - t0[k]))
        for c_, (kind, i) in enumerate(ext):                    # lb <= rho <= ub
            row = np.zeros(nz); row[nv + c_] = 1.0
            lo_, hi_ = (self.rslb[i], self.rsub[i]) if kind == 's' else (self.rilb[i], self.riub[i])
            self._extra_lin.append((row, float(lo_), float(hi_)))

        def quad(V):
            Rw = V.R[:, :nwq] @ T
            r0 = V.R[:, :nwq] @ t0 + V.R @ xi0
            return Rw, r0, V.W
        terms = [quad(V) for V in self.VJ]
        gam = np.array([V.gamma for V in self.VJ], dtype=float)
        Rt, rt0, Wt = quad(self.Vt)
        m = len(terms)

        def evaluate(v):
            z, ss = v[:nz], v[nz]
            rt = Rt @ z + rt0
            Vt = rt @ Wt @ rt
            Vi = np.empty(m); gi = np.zeros((m, nz + 1)); Hi = []
            for i, (Rw, r0, W) in enumerate(terms):
                r = Rw @ z + r0
                Vi[i] = r @ W @ r
                gi[i, :nz] = 2 * Rw.T @ W @ r
                Hi.append(2 * Rw.T @ W @ Rw)
            d = ss * gam - Vi
            return Vt, Vi, d, gi, Hi, rt

        def phi(v):
            Vt, Vi, d, *_ = evaluate(v)
            if np.any(d <= 0):
                return INF
            return Vt + v[nz] ** 2 - np.sum(np.log(d))

        # start: warm start if given, else zeros with ss large enough
        w0 = np.asarray(w0, float).reshape(-1)
        v = np.zeros(nz + 1)
        if w0.size == nwq + 1:
            v[:nv] = w0[:nv]; v[nz] = w0[-1]
        Vt, Vi, d, *_ = evaluate(v)
        v[nz] = max(v[nz], 1.0, 1.1 * float(np.max(Vi / gam)) + 0.1)
        iters = 0
        for iters in range(1, 200):
            Vt, Vi, d, gi, Hi, rt = evaluate(v)
            g = np.zeros(nz + 1); H = np.zeros((nz + 1, nz + 1))
            g[:nz] = 2 * Rt.T @ Wt @ rt
            H[:nz, :nz] = 2 * Rt.T @ Wt @ Rt
            g[nz] = 2 * v[nz]; H[nz, nz] = 2.0
            for i in range(m):
                a = gi[i].copy(); a[nz] = -gam[i]               # grad of V_i - ss*gamma_i  (= -grad d_i)
                g += a / d[i]
                H[:nz, :nz] += Hi[i] / d[i]
                H += np.outer(a, a) / d[i] ** 2
            p = -np.linalg.solve(H, g)
            dec = -g @ p
            if dec <= 1e-18 * (1.0 + abs(phi(v))):
                break
            t = 1.0
            f0 = phi(v)
            while phi(v + t * p) > f0 - 1e-4 * t * dec and t > 1e-12:
                t *= 0.5
            if t <= 1e-12:
                break                                          # rounding level reached
            v = v + t * p
        Vt, Vi, d, gi, Hi, rt = evaluate(v)
        z, dU, ss = v[:nz], v[:nv], v[nz]
        # explicit inequalities must be strictly inactive
        u_pred = np.asarray(u0)[None, :] + np.cumsum(dU.reshape(N, nu), axis=0)
        def margins(z, ss, d, u_pred):
            dU = z[:nv]
            mg = [ss - 1.0, float(np.min(d - 0.01)),
                  float(np.min(u_pred - self.ulb)), float(np.min(self.uub - u_pred)),
                  float(np.min(dU.reshape(N, nu) - self.dulb)), float(np.min(self.duub - dU.reshape(N, nu)))]
            for row, lo_, hi_ in self._extra_lin:
                mg.append(float(min(row @ z - lo_, hi_ - row @ z)))
            return mg
        margin = margins(z, ss, d, u_pred)
        if min(margin) <= 1e-9:
            # an explicit inequality of the NLP is active: solve the constrained problem with SciPy's trust-constr
            # (its own interior point / SQP machinery -- independent of the product's barrier continuation), from a
            # strictly feasible start, then polish and certify through the KKT conditions below.
            v, kkt = self._solve_constrained(evaluate, phi, Rt, Wt, gam, u0, nv, N, nu, m, nz, v_free=v.copy())
            Vt, Vi, d, gi, Hi, rt = evaluate(v)
            z, dU, ss = v[:nz], v[:nv], v[nz]
            u_pred = np.asarray(u0)[None, :] + np.cumsum(dU.reshape(N, nu), axis=0)
            margin = margins(z, ss, d, u_pred)
            self.last_kkt = kkt
        w = np.concatenate([T @ z + t0, [ss]])
        xi = self._xi(w, x0, u0, ySP)
        J = float(Vt + ss ** 2 - np.sum(np.log(d)))
        sol = {
            'J': J, 'du_opt': dU.reshape(N, nu).T.copy(), 'w_opt': w,
            'x_pred': np.concatenate([Xk @ xi for Xk in self.X_pred]),
            'y_pred': np.concatenate([Yk @ xi for Yk in self.Y_pred]),
            'u_pred': np.concatenate([Uk @ xi for Uk in self.U_pred]),
            'lam_w': np.zeros(w.size), 'lam_g': np.zeros(0),
            'info': {'iters': iters, 'margins': margin, 'V': Vi, 'd': d, 'Vt': float(Vt)},
        }
        warm = self.warmStart(sol, ySP)
        for i, V in enumerate(self.F_ViN):
            self.ViNant[i] = self.F(V, [0] * self.nx, [0] * self.nu, warm, [0] * self.ny)
        return sol

    def _solve_constrained(self, evaluate, phi, Rt, Wt, gam, u0, nv, N, nu, m, nz=None, v_free=None):
        """min phi(v) s.t. ss >= 1, ss*gamma_i - V_i >= 0.01, ulb <= u0 + cumsum(dU) <= uub, dulb <= dU <= duub
        (log.py:390-392,420-424,435-440).  Returns (v, kkt) with kkt = dict(stationarity, min multiplier)."""
        from scipy.optimize import minimize, LinearConstraint, NonlinearConstraint, Bounds
        nz = nv if nz is None else nz                     # decision variables z = [dU ; rho] (ranged terminal residuals), then ss
        n = nz + 1

        def fgh(v):
            Vt, Vi, d, gi, Hi, rt = evaluate(v)
            g = np.zeros(n); H = np.zeros((n, n))
            g[:nz] = 2 * Rt.T @ Wt @ rt; H[:nz, :nz] = 2 * Rt.T @ Wt @ Rt
            g[nz] = 2 * v[nz]; H[nz, nz] = 2.0
            for i in range(m):
                a = gi[i].copy(); a[nz] = -gam[i]
                g += a / d[i]; H[:nz, :nz] += Hi[i] / d[i]; H += np.outer(a, a) / d[i] ** 2
            return g, H
        # linear rows: cumulative input box and move box
        Tc = np.kron(np.tril(np.ones((N, N))), np.eye(nu))
        ne = len(getattr(self, '_extra_lin', []))
        A = np.zeros((2 * nv + ne, n)); A[:nv, :nv] = Tc; A[nv:2 * nv, :nv] = np.eye(nv)
        lo = np.concatenate([np.tile(self.ulb - u0, N), np.tile(self.dulb, N), np.zeros(ne)])
        hi = np.concatenate([np.tile(self.uub - u0, N), np.tile(self.duub, N), np.zeros(ne)])
        for k, (row, lo_, hi_) in enumerate(getattr(self, '_extra_lin', [])):       # siN contraction rows
            A[2 * nv + k, :nz] = row; lo[2 * nv + k] = lo_; hi[2 * nv + k] = hi_
        nrow = A.shape[0]

        def dfun(v):
            return evaluate(v)[2] - 0.01

        def djac(v):
            _, _, _, gi, _, _ = evaluate(v)
            J = -gi.copy(); J[:, nz] = gam
            return J
        def start_active(v, from_violation):
            cs = A @ v
            active = []
            tol_ = 0.0 if from_violation else 1e-7
            for k in range(nrow):
                if np.isfinite(lo[k]) and cs[k] - lo[k] < tol_: active.append(('lin', k, 1.0))
                elif np.isfinite(hi[k]) and hi[k] - cs[k] < tol_: active.append(('lin', k, -1.0))
            if v[nz] - 1.0 < tol_: active.append(('ss', 0, 1.0))
            for i in range(m):
                if dfun(v)[i] < tol_: active.append(('d', i, 1.0))
            return active

        # first attempt: primal active-set Newton from the UNCONSTRAINED optimum with the rows it violates held active
        # (linear rows: one Newton step lands on them); accepted only with its KKT certificate
        if v_free is not None:
            try:
                v1, kkt1 = self._polish(v_free.copy(), start_active(v_free, True), fgh, dfun, djac, phi, A, lo, hi, n, nz, nv, m)
                c1 = A @ v1
                feas = min(float(np.min(c1 - lo)), float(np.min(hi - c1)), v1[nz] - 1.0, float(np.min(dfun(v1))))
                if kkt1['stationarity'] <= 1e-9 and kkt1['lam_min'] >= -1e-10 and feas >= -1e-9:
                    return v1, kkt1
            except np.linalg.LinAlgError:
                pass
        # strictly feasible start: no moves, ss large
        v0 = np.zeros(n)
        Vi0 = evaluate(v0)[1]
        v0[nz] = max(2.0, 1.1 * float(np.max((Vi0 + 0.02) / gam)) + 0.1)
        big = 1e30

        def fsafe(v):
            f = phi(v)
            return f if np.isfinite(f) else big
        res = minimize(fsafe, v0, jac=lambda v: fgh(v)[0], hess=lambda v: fgh(v)[1], method='trust-constr',
                       bounds=Bounds(np.r_[np.full(nz, -np.inf), 1.0], np.full(n, np.inf), keep_feasible=True),
                       constraints=[LinearConstraint(A, lo, hi),      # (a state on a bound up to rounding: no keep_feasible)
                                    NonlinearConstraint(dfun, 0.0, np.inf, jac=djac, keep_feasible=True)],
                       options=dict(gtol=1e-8, xtol=1e-12, barrier_tol=1e-9, maxiter=400, initial_barrier_parameter=0.1))
        v = res.x
        return self._polish(v, start_active(v, False), fgh, dfun, djac, phi, A, lo, hi, n, nz, nv, m)

    @staticmethod
    def _polish(v, active, fgh, dfun, djac, phi, A, lo, hi, n, nz, nv, m):
        """Primal active-set Newton.  Rows: ('lin', k, +1/-1) of A v >= lo / <= hi, ('ss',), ('d', i).  Returns the
        point and its KKT record (stationarity of the Lagrangian, multipliers in casadi's convention)."""
        nrow = A.shape[0]
        lam = np.zeros(0); stat = np.inf
        for _ in range(1000):
            g, H = fgh(v)
            dj = djac(v); dv_ = dfun(v); cs = A @ v
            rows, rhs = [], []
            for kind, k, sg in active:
                if kind == 'lin': rows.append(A[k]); rhs.append((lo[k] if sg > 0 else hi[k]) - cs[k])
                elif kind == 'ss': e_ = np.zeros(n); e_[nz] = 1.0; rows.append(e_); rhs.append(1.0 - v[nz])
                else: rows.append(dj[k]); rhs.append(-dv_[k])
            Aa = np.array(rows).reshape(-1, n); ba = np.array(rhs)
            K = np.block([[H, Aa.T], [Aa, np.zeros((len(ba), len(ba)))]])
            sol_ = np.linalg.lstsq(K, np.r_[-g, ba], rcond=None)[0]
            step = sol_[:n]; lam = -sol_[n:]
            # longest feasible fraction of the step; the blocking row joins the active set
            t, block = 1.0, None
            act_lin = {(k, sg) for kind, k, sg in active if kind == 'lin'}
            ds = A @ step
            for k in range(nrow):
                if np.isfinite(lo[k]) and (k, 1.0) not in act_lin and ds[k] < -1e-16:
                    tk = (lo[k] - cs[k]) / ds[k]
                    if tk < t: t, block = max(tk, 0.0), ('lin', k, 1.0)
                if np.isfinite(hi[k]) and (k, -1.0) not in act_lin and ds[k] > 1e-16:
                    tk = (hi[k] - cs[k]) / ds[k]
                    if tk < t: t, block = max(tk, 0.0), ('lin', k, -1.0)
            if ('ss', 0, 1.0) not in active and step[nz] < -1e-16:
                tk = (1.0 - v[nz]) / step[nz]
                if tk < t: t, block = max(tk, 0.0), ('ss', 0, 1.0)
            restoring = len(ba) > 0 and float(np.max(np.abs(ba))) > 1e-9      # step back ONTO violated rows: phi may rise
            f_here, slope = phi(v), min(float(g @ step), 0.0)
            while t > 1e-10:                                   # barrier domain and d_i >= 0.01 (nonlinear): backtrack;
                vn = v + t * step                              # Armijo on phi: a Newton step from far away (trust-constr
                dn = dfun(vn)                                  # stopped early) must not leave the region where it helps
                inact = [i for i in range(m) if ('d', i, 1.0) not in active]
                f_new = phi(vn)
                if np.isfinite(f_new) and all(dn[i] >= -1e-13 for i in inact) and \
                        (restoring or f_new <= f_here + 1e-4 * t * slope + 1e-13 * (1.0 + abs(f_here))):
                    break
                t *= 0.5; block = None
            v = v + t * step
            if block is not None:
                active.append(block)
                continue
            stat = float(np.max(np.abs(g - Aa.T @ lam))) if len(ba) else float(np.max(np.abs(g)))
            if np.linalg.norm(t * step) < 1e-13 * (1 + np.linalg.norm(v)):
                sg = np.array([a_[2] for a_ in active])
                if len(active) and np.min(sg * lam) < -1e-10:
                    active.pop(int(np.argmin(sg * lam)))     # wrong-sign multiplier: release that row
                    continue
                break
        sg = np.array([a_[2] for a_ in active])
        lam_min = float(np.min(sg * lam)) if len(active) else 0.0       # >= 0: multipliers have the KKT sign
        # multipliers in casadi's convention (grad f + sum lam grad g = 0, > 0 at an upper bound): row kinds
        # ('U', k) input box row k of U_pred, ('dU', k) move box, ('lin', k) extra linear row, ('ss',), ('d', i)
        duals = {}
        for (kind, k, sg_), la in zip(active, lam):
            key = ('U', k) if (kind == 'lin' and k < nv) else (('dU', k - nv) if (kind == 'lin' and k < 2 * nv) else
                  (('lin', k - 2 * nv) if kind == 'lin' else ((kind,) if kind == 'ss' else (kind, k))))
            duals[key] = -float(la)
        return v, dict(stationarity=stat, n_active=len(active), lam=lam, lam_min=lam_min, duals=duals)

    def warmStart(self, sol, ysp):                                # log.py:518-549
        w0 = np.asarray(sol['w_opt']).reshape(-1)
        ny, nu, nx = self.ny, self.nu, self.nx
        dustart = np.concatenate([w0[:-2 * ny - 1][nu:], np.zeros(nu)])
        xN = np.asarray(sol['x_pred']).reshape(-1)[-nx:]
        Xknext = self.sys.A @ xN
        siNnext = Xknext[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        syNnext = w0[-2 * ny:-ny]          # log.py:543 verbatim: on the ss-extended vector this is [syN[1:], siN[0]]
        return np.concatenate([dustart, syNnext, siNnext, [1.0]])       # log.py:548

    def satWeights(self, x, u, w0, ysp):                          # log.py:566-574
        return np.array([1.0 / (V.gamma - self.F(V, x, u, w0, ysp)) for V in self.VJ])


def declare_distillation_log(cls, sys, **bounds):
    """example_distillation_log.py:43-61 (init_Qt is referenced without being called there)."""
    N = 10
    c = cls(sys, N, **{**dict(ulb=[0, 0]), **bounds})
    c.subObjComposed(y=[0], Q=1, sat=N * 0.5 ** 2)
    c.subObjComposed(y=[1], Q=1, sat=N * 2.0 ** 2)
    c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2)
    c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.subObj(siN=[1], Q=1, addJ=False)
    return c
