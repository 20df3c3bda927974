# -*- coding: utf-8 -*-
"""Drop-in mirror of the reference's LOG-BARRIER satisficing IHMPC controller surface,
/root/reference/sihmpc/ihmpccontroller_log.py (class IHMPCController, log.py:10), so that the loop of
/root/reference/example_distillation_log.py:63-135 runs against it.  `mpc()` runs the damped Newton
kernel of libsihmpc_b200.so (sihmpc_log_step); there is no CPU fallback.

Differences from the weighted class, all taken from the reference:
  * decision vector w = [dU ; syN ; siN ; ss] with ss >= 1 (log.py:375-392);
  * J = Vt + ss^2 - sum_i log(ss*gamma_i - V_i)  (log.py:71-92,196,219,241,262,292); V = [Vt, Vss, ...];
  * `Qt` constructor kwarg, default eye(ny) (log.py:55); `init_Qt()` only recomputes self.Qt -- the
    reference overwrites V[0].V but neither J nor V[0].F (log.py:94-101), and its own example references
    `c.init_Qt` without calling it (example_distillation_log.py:61);
  * mpc(x0, ySP, w0, u0, lam_w0, lam_g0, ViN_ant=None) (log.py:552); `ViN_ant=[]` means zeros (casadi
    call semantics), i.e. siN = 0;
  * warmStart appends ss = 1 (log.py:548); satWeights returns the theoretical weights 1/(gamma - V)
    (log.py:566-574).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.linalg import solve_discrete_lyapunov

from .ihmpccontroller import IHMPCController as _Weighted, DM, _vec
from .plan import SubObjective, compile_plan

__all__ = ["IHMPCController", "DM"]


class IHMPCController(_Weighted):

    def __init__(self, sys, N, **kwargs):                      # log.py:11-92
        super().__init__(sys, N, **kwargs)
        self.Qt = np.asarray(kwargs.get('Qt', np.eye(sys.ny)), dtype=float)       # log.py:55
        self.Q_bar = self._terminal_weight()
        z = np.zeros
        self.Vt = self.fObj(self, SubObjective('Vt', z(self.ny), z(self.ny), z(self.ny), z(self.nu), terminal=True), None, 'Vt')
        self.Vss = self.fObj(self, SubObjective('Vss', z(self.ny), z(self.ny), z(self.ny), z(self.nu)), None, 'Vss')
        self.Vt.weight = 1
        self.Vss.weight = 1
        self.V = [self.Vt, self.Vss]                           # log.py:83
        self.VJ = []
        self.status = 0
        self.newton_iters = 0

    def _terminal_weight(self):                                # log.py:304-317 (F, bilinear)
        F, Psi = np.asarray(self.sys.F, float), np.asarray(self.sys.Psi, float)
        Q_lyap = F.T @ Psi.T @ self.Qt @ Psi @ F
        return solve_discrete_lyapunov(F, Q_lyap, method='bilinear')

    def init_Qt(self):                                         # log.py:94-101
        Qt = []
        for V in self.VJ:
            if V.varType == 'y':
                _, qd = self._qdiag({'Q': V.Q})
                Qt = np.append(Qt, (1 / V.gamma) * qd)
        self.Qt = np.diag(Qt)
        # log.py:100-101 only replaces the *expression* V[0].V; the NLP objective J and V[0].F keep the
        # terminal cost built in __init__, so the compiled plan does not change either.

    def set_terminal_objective(self):
        raise AttributeError("the log controller has no set_terminal_objective (terminal cost is built in __init__)")

    def init_pesos(self):
        raise AttributeError("the log controller has no weights (ihmpccontroller_log.py)")

    def satWeights2(self, *a, **k):
        raise AttributeError("the log controller has no satWeights2 (ihmpccontroller_log.py)")

    def _new(self, kind, inds, Q, qd, kwargs, prefix):
        V = super()._new(kind, inds, Q, qd, kwargs, prefix)
        V.weight = 1                                           # log.py:133
        return V

    def subObjComposed(self, **kwargs):
        V = super().subObjComposed(**kwargs)
        if V is not None:
            V.weight = 1
        return V

    # ---- evaluation: var = [dU ; syN ; siN ; ss]
    def _eval_all(self, x, w, ysp):
        w = np.asarray(w, dtype=float).reshape(-1)
        nwq = self.N * self.nu + 2 * self.ny
        ss = w[nwq] if w.size == nwq + 1 else 0.0
        vals = super()._eval_all(x, w[:nwq], ysp)
        vals[self.V.index(self.Vss)] = ss * ss                 # log.py:76-80
        return vals

    def compile(self):
        if not self._dirty and self._plan is not None:
            return self._plan
        if self._ell:
            raise NotImplementedError("siN sub-objective over several outputs (quadratic row) in the log variant")
        if not self.VJ:
            raise ValueError("no log sub-objective declared")
        self._plan = compile_plan(self.sys, self.N, [V.coeff for V in self.VJ], self._bounds(),
                                  Q_bar=self.Q_bar, vin=self._vin,
                                  rs=np.asarray(self.rslb, float), ri=np.asarray(self.rilb, float),
                                  rs_hi=np.asarray(self.rsub, float), ri_hi=np.asarray(self.riub, float))
        # terminal residual RANGES (rs/ri lb != ub, log.py:413-427): the ranged residual rho is a decision variable after
        # the moves (Plan.n_extra), its range a dense row of the kernel's interior-point phase
        self._eval_plan = None
        self._stepper = None
        self._eval_dp = None
        self._dirty = False
        return self._plan

    def engine(self, batch, **kw):
        """Batched closed loop of the log controller on the device (sihmpc_log_loop)."""
        from .engine import LogBatchedEngine
        return LogBatchedEngine(self.compile(), batch, **kw)

    # ---- step (log.py:552-563)
    def mpc(self, x0, ySP, w0, u0, lam_w0, lam_g0, ViN_ant=None):
        import torch
        from . import _capi
        hs = self._get_stepper()
        dp = hs.dp
        nx, nu, ny, N = self.nx, self.nu, self.ny, self.N
        nv = N * nu
        nw = nv + 2 * ny + 1
        if ViN_ant is None:                                    # log.py:554-555
            ViN_ant = self.ViNant
        vin = _vec(ViN_ant, len(self.ViNant), "ViN_ant") if len(self.ViNant) else np.zeros(0)
        x = _vec(x0, nx, "x0"); u = _vec(u0, nu, "u0"); ysp = _vec(ySP, ny, "ySP")
        w0a = np.asarray(w0, dtype=np.float64).reshape(-1)
        dev = dp.device
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).reshape(1, -1).to(dev)
        xd, ud, yd = t(x), t(u), t(ysp)
        w0d = t(w0a) if w0a.size == nw else None
        wo = torch.zeros(1, nw, dtype=torch.float64, device=dev)
        Jd = torch.zeros(1, dtype=torch.float64, device=dev)
        sd = torch.zeros(1, dtype=torch.int32, device=dev)
        itd = torch.zeros(1, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        # one step of sihmpc_log_loop without SIHMPC_F_ADVANCE = mpc() with the contraction limits ViN_ant of this call
        # (dense rows |siN| <= sqrt(ViN_ant/q) of an integrating plant, log.py:245-270,429-434)
        io = _capi.IO()
        io.x, io.u, io.ysp = xd.data_ptr(), ud.data_ptr(), yd.data_ptr()
        vind = t(vin) if (len(vin) and self._plan.n_vin > 0) else None
        io.vinant = vind.data_ptr() if vind is not None else None
        io.w_opt, io.J, io.status = wo.data_ptr(), Jd.data_ptr(), sd.data_ptr()
        io.flags = 0
        with torch.cuda.device(dev):
            _capi.check(dp.lib.sihmpc_log_loop(
                dp.handle, C.byref(io), C.c_void_p(w0d.data_ptr()) if w0d is not None else None, 1, 1,
                C.c_void_p(itd.data_ptr()), C.c_void_p(st)), "sihmpc_log_loop")
        w = wo.cpu().numpy().reshape(-1)
        J = float(Jd.cpu().numpy()[0])
        self.status = int(sd.cpu().numpy()[0])
        self.newton_iters = int(itd.cpu().numpy()[0])
        if self.status & _capi.ST_ACTIVE_UNSUPPORTED:
            raise NotImplementedError("an explicit inequality of the log NLP (ss >= 1, ss*gamma_i - V_i >= 0.01, "
                                      "input/move boxes) is active and the interior-point solve found no strictly "
                                      "feasible start (u0 outside its box, or 0 outside the move box)")
        x_pred, y_pred, u_pred = self._predict(x, u, w[:nv])
        vals = self._eval_all(x, w, ysp)
        XN = x_pred[-nx:]
        res1 = XN[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi] - w[nv + ny:nv + 2 * ny]
        res2 = XN[:self.nxs] - ysp - w[nv:nv + ny]
        sat = np.array([w[-1] * V.gamma - vals[self.V.index(V)] for V in self.VJ])
        gvec = np.concatenate([x_pred, u_pred, res1, res2, vals, sat])           # log.py:394-440
        lam_w, lam_g = self._log_duals(x, u, ysp, w, vals, vin, gvec.size)
        sol = {
            'J': DM(J), 'du_opt': DM(w[:nv].reshape(N, nu).T), 'w_opt': DM(w),
            'lam_w': DM(lam_w), 'lam_g': DM(lam_g), 'g': DM(gvec),
            'x_pred': DM(x_pred), 'y_pred': DM(y_pred), 'u_pred': DM(u_pred),
        }
        if len(self.ViNant):                                   # log.py:558-561
            warm = self.warmStart(sol, ysp)
            allv = self._eval_all(np.zeros(nx), warm, np.zeros(ny))
            k = 0
            for V in self.V:
                if V.coeff.vin_slot >= 0:
                    self.ViNant[V.coeff.vin_slot] = float(allv[self.V.index(V)])
                    k += 1
        self.du = sol['du_opt'][:, 0].full()
        self.j_value = J
        return sol

    def _log_duals(self, x, u, ysp, w, vals, vin, dim_g):
        """sol['lam_w'] (nw + 1: moves, syN, siN, ss) and sol['lam_g'] (g layout of log.py:394-440: X_pred, U_pred, res1,
        res2, the len(V) rows 0 <= V_i, the m rows ss gamma_i - V_i >= 0.01), casadi's sign convention, which the
        reference feeds back as IPOPT's warm start (example_distillation_log.py:101-103).  The kernel is a primal
        barrier method, so the multipliers are recovered on the host from the optimum it returns: the inequality rows
        that sit on their bound (move box, U_pred rows, ss >= 1, the 0.01 rows, siN contraction) by least squares on
        stationarity  grad phi + sum lam grad g = 0  in (dU, ss), the terminal-residual rows res1 / res2 from
        stationarity in siN / syN like the weighted class (the weights being 1/d_i here)."""
        p = self._plan
        N, nu, ny, nx, m = self.N, self.nu, self.ny, self.nx, len(self.VJ)
        nv = N * nu
        dU, ss = w[:nv], w[-1]
        ne = p.n_extra
        nz = nv + ne
        e = p.Px @ x + p.Py @ ysp + p.e0
        o_sy_, _ = p.row_slices['syN']; o_si_, _ = p.row_slices['siN']
        r_m = p.E[:, :nv] @ dU + e                             # residuals with rho = 0
        rho = np.array([r_m[(o_sy_ if kind == 's' else o_si_) + i] - w[nv + (0 if kind == 's' else ny) + i]
                        for kind, i in p.extra], dtype=float)     # syN = xs(N) - ysp - rho_s, siN = xi(N) - rho_i
        z = np.concatenate([dU, rho])
        Vi = np.array([vals[self.V.index(V)] for V in self.VJ])
        gam = np.array([V.gamma for V in self.VJ], dtype=float)
        d = ss * gam - Vi
        inv = 1.0 / d
        r = p.E @ z + e
        om = p.Wrow @ inv + p.w0row
        grad_V = [2.0 * p.E.T @ (p.Wrow[:, i] * r) + 2.0 * p.Wdu[:, i] * z for i in range(m)]
        g = np.zeros(nz + 1)
        g[:nz] = 2.0 * p.E.T @ (om * r) + 2.0 * (p.Wdu @ inv) * z
        g[nz] = 2.0 * ss - float(gam @ inv)
        rows, keys = [], []                                   # (normal over (dU, ss), key) of the rows on a bound
        tol = 1e-7
        Tc = np.kron(np.tril(np.ones((N, N))), np.eye(nu))
        up = u[None, :] + np.cumsum(dU.reshape(N, nu), axis=0)
        for k in range(nv):
            j = k % nu
            if abs(dU[k] - np.asarray(self.dulb, float)[j]) <= tol or abs(dU[k] - np.asarray(self.duub, float)[j]) <= tol:
                n_ = np.zeros(nz + 1); n_[k] = 1.0; rows.append(n_); keys.append(('w', k))
            uk = up[k // nu, j]
            if abs(uk - np.asarray(self.ulb, float)[j]) <= tol or abs(uk - np.asarray(self.uub, float)[j]) <= tol:
                n_ = np.zeros(nz + 1); n_[:nv] = Tc[k]; rows.append(n_); keys.append(('U', k))
        if ss - 1.0 <= tol:
            n_ = np.zeros(nz + 1); n_[nz] = 1.0; rows.append(n_); keys.append(('ss',))
        for i in range(m):
            if d[i] - 0.01 <= tol:
                n_ = np.zeros(nz + 1); n_[:nz] = -grad_V[i]; n_[nz] = gam[i]; rows.append(n_); keys.append(('d', i))
        siN = w[nv + ny:nv + 2 * ny]
        for rr, kind in enumerate(p.row_kind):
            if kind[0] == 'vin' and len(vin) and np.isfinite(vin[kind[1]]):
                ind, q = self._vin[kind[1]]
                if abs(abs(siN[ind]) - np.sqrt(max(vin[kind[1]], 0.0) / q)) <= tol:
                    n_ = np.zeros(nz + 1); n_[:nz] = 2.0 * q * siN[ind] * p.Gc[rr][:nz]; rows.append(n_); keys.append(('vin', kind[1]))
            elif kind[0] == 'rho':                                # a ranged residual on an end of its range: its multiplier is
                c_ = kind[1]                                      # the one of the res1 / res2 row, set from stationarity below
                if abs(rho[c_] - p.Lc[rr]) <= tol or abs(rho[c_] - p.Uc[rr]) <= tol:
                    n_ = np.zeros(nz + 1); n_[nv + c_] = 1.0; rows.append(n_); keys.append(('rho', c_))
        lam = np.zeros(0)
        if rows:
            lam = np.linalg.lstsq(np.array(rows).T, -g, rcond=None)[0]
        lam_w = np.zeros(nv + 2 * ny + 1)
        lam_g = np.zeros(dim_g)
        o_res1, o_res2, o_V, o_sat = N * nx + nv, N * nx + nv + ny, N * nx + nv + 2 * ny, N * nx + nv + 2 * ny + len(self.V)
        extra_si = np.zeros(ny)
        for key, la in zip(keys, lam):
            if key[0] == 'w':
                lam_w[key[1]] = la
            elif key[0] == 'U':
                lam_g[N * nx + key[1]] = la
            elif key[0] == 'ss':
                lam_w[-1] = la
            elif key[0] == 'd':
                lam_g[o_sat + key[1]] = la
            elif key[0] == 'vin':
                slot = key[1]
                owner = [i for i, (kd, k) in enumerate(self._vin_of) if kd == 'row' and k == slot][0]
                V = next(V for V in self.V if V.coeff.vin_slot == owner)
                lam_g[o_V + self.V.index(V)] = la
                ind, q = self._vin[slot]
                extra_si[ind] += 2.0 * q * siN[ind] * la
        # res2 / res1: stationarity in syN / siN with the weights 1/d_i plus the multipliers of the 0.01 rows
        wts = inv.copy()
        for key, la in zip(keys, lam):
            if key[0] == 'd':
                wts[key[1]] -= la                               # d/dsyN of lam (ss gamma - V_i) = -lam dV_i/dsyN
        omr = p.Wrow @ wts + p.w0row
        syN = w[nv:nv + ny]
        Y = (p.Myx @ x + p.Myu[:, :nv] @ dU).reshape(N, ny)
        r_y = Y - ysp[None, :] - syN[None, :] - p.tau[:, None] * siN[None, :]
        om_y = omr[:N * ny].reshape(N, ny)
        a, b = p.row_slices['syN']; om_sy = omr[a:b]
        a, b = p.row_slices['siN']; om_si = omr[a:b]
        lam_g[o_res2:o_res2 + ny] = -2.0 * np.sum(om_y * r_y, axis=0) + 2.0 * om_sy * syN
        lam_g[o_res1:o_res1 + ny] = -2.0 * np.sum(om_y * r_y * p.tau[:, None], axis=0) + 2.0 * om_si * siN + extra_si
        return lam_w, lam_g

    def warmStart(self, sol, ysp):                             # log.py:518-549
        w0 = np.asarray(sol['w_opt']).reshape(-1)
        ny, nu, nx = self.ny, self.nu, self.nx
        dustart = np.concatenate([w0[:-2 * ny - 1][nu:], np.zeros(nu)])
        xN = np.asarray(sol['x_pred']).reshape(-1)[-nx:]
        res = self.dynF(x0=xN, u0=np.zeros(nu), du0=np.zeros(nu))
        Xknext = res['xkp1'].full().reshape(-1)
        siNnext = Xknext[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        # log.py:543 slices w0[-2*ny:-ny] on the vector that ends with ss, i.e. [syN[1:], siN[0]] -- kept as the
        # reference has it (the value only seeds the next solve and the satWeights diagnostic of
        # example_distillation_log.py:131; the optimum does not depend on it).  The device loop (sihmpc_log_loop)
        # seeds its Newton iteration with the unshifted syN instead: same optimum, DESIGN.md "parity notes".
        syNnext = w0[-2 * ny:-ny]
        return np.concatenate([dustart, syNnext, siNnext, [1.0]])                # log.py:548

    def satWeights(self, x, u, w0, ysp):                       # log.py:566-574
        vals = self._eval_all(x, w0, ysp)
        return np.array([1 / (V.gamma - vals[self.V.index(V)]) for V in self.VJ])
