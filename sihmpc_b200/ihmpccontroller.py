# -*- coding: utf-8 -*-
"""Drop-in mirror of the reference's weighted satisficing IHMPC controller surface.

Same names, argument meaning and error behaviour as
/root/reference/sihmpc/ihmpccontroller.py (class IHMPCController, ctl.py:11), so
/root/reference/example_distillation.py:42-113 runs unchanged against it; the arithmetic of
``mpc()`` runs on the GPU through libsihmpc_b200.so.  What the reference rebuilds with
casadi on every call (``_MPC()``, ctl.py:569) is compiled ONCE into a plan
(sihmpc_b200.plan.compile_plan) the first time ``mpc``/``compile`` is called after the last
declaration.

For thousands of scenarios use ``controller.engine(batch)`` (sihmpc_b200.engine.BatchedEngine);
this class keeps the reference's single-scenario calling convention.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .opom import OPOM
from .plan import SubObjective, compile_plan, fixed_row_violation, terminal_weight

__all__ = ["IHMPCController", "DM"]


class DM(np.ndarray):
    """Minimal stand-in for casadi.DM: always 2-D, ``.full()``, ``float()``, casadi-style slices."""

    def __new__(cls, a):
        arr = np.array(a, dtype=np.float64)
        if arr.ndim == 0:
            arr = arr.reshape(1, 1)
        elif arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        return arr.view(cls)

    def full(self):
        return np.array(self, dtype=np.float64)

    def __getitem__(self, key):
        base = np.asarray(self)
        if isinstance(key, tuple):
            key = tuple(slice(k, k + 1 if k != -1 else None) if isinstance(k, (int, np.integer)) else k
                        for k in key)
            out = base[key]
        else:
            if isinstance(key, (int, np.integer)):
                key = slice(key, key + 1 if key != -1 else None)
            out = base[key] if base.shape[1] == 1 or base.shape[0] != 1 else base[:, key]
        return DM(out) if np.ndim(out) else DM(out)

    def __float__(self):
        if self.size != 1:
            raise TypeError("only a 1x1 DM converts to float")
        return float(np.asarray(self).reshape(-1)[0])


def _vec(v, n, name):
    """list / ndarray / column vector / [] (= zeros, casadi call semantics) -> flat float64[n]."""
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 0:
        return np.zeros(n)
    if a.size != n:
        raise ValueError("%s must have %d entries, got %d" % (name, n, a.size))
    return a.copy()


class _Weight:
    def __init__(self, name):
        self._name = name

    def name(self):
        return self._name


class _DynF:
    """c.dynF(x0=, u0=, du0=) -> {'xkp1','ykp1','ukp1'} (ctl.py:93-113)."""

    def __init__(self, ctrl):
        self._c = ctrl

    def __call__(self, x0=None, u0=None, du0=None):
        c = self._c
        x = _vec(x0, c.nx, "x0"); u = _vec(u0, c.nu, "u0"); du = _vec(du0, c.nu, "du0")
        xk = c.sys.A @ x + c.sys.B @ du
        yk = c.sys.C @ xk + c.sys.D @ du
        return {"xkp1": DM(xk), "ykp1": DM(yk), "ukp1": DM(u + du)}


class IHMPCController(object):

    class fObj:                                               # ctl.py:116-149
        def __init__(self, ctrl, coeff: SubObjective, weight_name, name=""):
            self._ctrl = ctrl
            self.coeff = coeff
            self.name = name
            self.V = coeff                                    # "expression": the coefficient form
            self.min = 0
            self.max = np.inf
            self.gamma = np.inf
            self.weight = _Weight(weight_name) if weight_name is not None else 1
            self.Vi = []
            self.peso = 0
            self.v_hist = []
            self.peso_hist = []

        def F(self, x0, u0, var, ysp):
            """Value at explicit var = [dU ; syN ; siN] (ctl.py:124-126)."""
            vals = self._ctrl._eval_all(x0, var, ysp)
            return DM(vals[self._ctrl.V.index(self)])

        def lim(self, min, max):
            self.min = min
            self.max = max

        def satLim(self, gamma):
            self.gamma = gamma
            self.coeff.gamma = gamma

        def setName(self, name):
            self.name = name

        def setType(self, type='simple'):
            self.type = type

        def setVarType(self, varType):
            self.varType = varType

        def setIndex(self, indx):
            self.index = indx

        def setQ(self, Q):
            self.Q = Q

    # ------------------------------------------------------------------ construction
    def __init__(self, sys, N, **kwargs):                     # ctl.py:12-85
        assert issubclass(type(sys), OPOM), 'the system (sys) must be of type OPOM'
        self.sys = sys
        self.Ts = sys.Ts
        self.N = int(N)
        self.theta_max = sys.theta_max
        assert (self.N > self.theta_max), \
            'the control horizon should be greater than the maximal system dead time: {}' \
            .format(self.theta_max)
        self.nx = sys.nx
        self.nu = sys.nu
        self.ny = sys.ny
        self.nxs = sys.ny
        self.nxd = sys.nd
        self.nxi = sys.ny
        self.nxz = sys.nz
        assert (self.nx == self.nxs + self.nxd + self.nxi + self.nxz)

        g = kwargs.get
        self.xlb = g('xlb', np.ones(sys.nx) * -np.inf); self.xub = g('xub', np.ones(sys.nx) * np.inf)
        self.ulb = g('ulb', np.ones(sys.nu) * -np.inf); self.uub = g('uub', np.ones(sys.nu) * np.inf)
        self.dulb = g('dulb', np.ones(sys.nu) * -np.inf); self.duub = g('duub', np.ones(sys.nu) * np.inf)
        self.sulb = g('sulb', np.ones(sys.nu) * -np.inf); self.suub = g('suub', np.ones(sys.nu) * np.inf)
        self.sylb = g('sylb', np.ones(sys.ny) * -np.inf); self.syub = g('syub', np.ones(sys.ny) * np.inf)
        self.silb = g('silb', np.ones(sys.ny) * -np.inf); self.siub = g('siub', np.ones(sys.ny) * np.inf)
        self.rslb = g('rslb', np.zeros(sys.ny)); self.rsub = g('rsub', np.zeros(sys.ny))
        self.rilb = g('rilb', np.zeros(sys.ny)); self.riub = g('riub', np.zeros(sys.ny))

        self.dynF = _DynF(self)
        self.V = []
        self.VJ = []
        self.F_ViN = []
        self.Pesos = []
        self.ViN_ant = []
        self.J = 0
        self._s = []
        self.ViNant = []
        self.pesos = []
        self.j_value = 0
        self.du = np.zeros((self.nu, 1))
        self.j_hist = []
        self.Qt = None
        self.Q_bar = None
        self._vin = []              # (output index, q) per contraction slot of the PLAN (dense row |siN| <= sqrt(ViN/q))
        self._vin_of = []           # per entry of self.ViNant: ('row', plan slot) or ('ell', index into self._ell)
        self._ell = []              # multi-index siN rows (ellipsoids): dict(V, inds, q, slots = plan slots of its box)
        self._dirty = True
        self._plan = None
        self._eval_plan = None
        self._stepper = None
        self._eval_dp = None
        self._hot_dp = None

    def init_pesos(self):                                     # ctl.py:87-91
        self.pesos = np.array([V.peso for V in self.VJ], dtype=float)
        return self.pesos

    # ------------------------------------------------------------------ declarations
    def _qdiag(self, kwargs):
        if 'Q' in kwargs:
            Q = kwargs['Q']
            if isinstance(Q, (int, np.integer)):
                Q = [Q]                                       # ctl.py:164
        else:
            Q = np.eye(self.sys.ny)                           # ctl.py:166
        qa = np.asarray(Q, dtype=float)
        return Q, (qa.reshape(-1) if qa.ndim <= 1 else np.diag(qa))

    def _new(self, kind, inds, Q, qd, kwargs, prefix):
        ny, nu = self.ny, self.nu
        z = lambda n: np.zeros(n)
        co = SubObjective(name=prefix + str(inds), cy=z(ny), csy=z(ny), csi=z(ny), cdu=z(nu))
        for j, ind in enumerate(inds):
            getattr(co, {'y': 'cy', 'du': 'cdu', 'syN': 'csy', 'siN': 'csi'}[kind])[ind] += qd[j]
        V = self.fObj(self, co, {'y': 'w_y', 'du': 'w_du', 'syN': 'w_syN', 'siN': 'w_siN'}[kind] + str(inds))
        V.setName(prefix + str(inds))
        V.setVarType(kind)
        V.setIndex(inds)
        V.setQ(Q)
        if 'sat' in kwargs:
            V.satLim(kwargs['sat'])
            V.peso = 1 / kwargs['sat']                        # ctl.py:184
            co.peso = V.peso
        self.V.append(V)
        if kind == 'siN':
            self.F_ViN.append(V.F)                            # ctl.py:259
        if 'addJ' not in kwargs or kwargs['addJ'] != False:
            self.Pesos.append(V.weight)
            self.VJ.append(V)
        if kind == 'siN':                                     # contraction (ctl.py:265-269)
            V.lim(0, 'ViN_ant_' + str(inds))
            co.vin_slot = len(self.ViNant)
            self.ViN_ant.append('ViN_ant_' + str(inds))
            self.ViNant.append(np.inf)
            if len(inds) == 1:
                self._vin_of.append(('row', len(self._vin)))
                self._vin.append((int(inds[0]), float(qd[0])))
            else:
                # sum_j q_j siN_j^2 <= ViN_ant over several outputs (ctl.py:246-247,266): ONE convex quadratic row.
                # It enters the plan as one more weighted sub-objective whose weight is the row's MULTIPLIER (found by
                # mpc() in an outer search, every trial a device QP solve), plus its box |siN_j| <= sqrt(ViN_ant/q_j)
                # as ordinary contraction rows (implied by the ellipsoid; exact when ViN_ant = 0)
                slots = []
                for j, ind in enumerate(inds):
                    if qd[j] > 0:
                        slots.append(len(self._vin))
                        self._vin.append((int(ind), float(qd[j])))
                self._vin_of.append(('ell', len(self._ell)))
                self._ell.append(dict(V=V, inds=[int(i) for i in inds], q=[float(v) for v in qd[:len(inds)]], slots=slots))
        self._dirty = True
        return V

    def subObj(self, **kwargs):                               # ctl.py:151-271
        Q, qd = self._qdiag(kwargs)
        if 'y' in kwargs:
            return self._new('y', kwargs['y'], Q, qd, kwargs, 'Vy_')
        if 'du' in kwargs:
            return self._new('du', kwargs['du'], Q, qd, kwargs, 'Vdu_')
        if 'syN' in kwargs:
            return self._new('syN', kwargs['syN'], Q, qd, kwargs, 'VsyN_')
        if 'siN' in kwargs:
            return self._new('siN', kwargs['siN'], Q, qd, kwargs, 'VsiN_')

    def subObjComposed(self, **kwargs):                       # ctl.py:273-305
        if 'y' in kwargs:
            inds = kwargs['y']
            Vy = self.subObj(y=inds, Q=kwargs['Q'], addJ=False)
            VyN = self.subObj(syN=inds, Q=kwargs['Q'], addJ=False)
            co = SubObjective(name='VyC_' + str(inds), cy=Vy.coeff.cy.copy(),
                              csy=self.N * VyN.coeff.csy, csi=np.zeros(self.ny), cdu=np.zeros(self.nu))
            Vc = self.fObj(self, co, 'w_yC' + str(inds))
            if 'addJ' not in kwargs or kwargs['addJ'] != False:
                self.Pesos.append(Vc.weight)
                self.VJ.append(Vc)
            Vc.setType('composed')
            Vc.setName('VyC_' + str(inds))
            Vc.setVarType('y')
            Vc.setIndex(inds)
            Vc.Vi = [Vy, VyN]
            Vc.setQ(Vy.Q)
            if 'sat' in kwargs:
                Vc.satLim(kwargs['sat'])
                Vc.peso = 1 / kwargs['sat']
                co.peso = Vc.peso
            self.V.append(Vc)
            self._dirty = True
            return Vc

    def set_terminal_objective(self):                         # ctl.py:307-343
        Qt = []
        for V in self.VJ:
            if V.varType == 'y':
                _, qd = self._qdiag({'Q': V.Q})
                Qt = np.append(Qt, (1 / V.gamma) * qd)
        self.Qt = np.diag(Qt)
        self.Q_bar = terminal_weight(self.sys, self.Qt)
        co = SubObjective(name='Vt', cy=np.zeros(self.ny), csy=np.zeros(self.ny), csi=np.zeros(self.ny),
                          cdu=np.zeros(self.nu), terminal=True)
        Vt = self.fObj(self, co, None, 'Vt')
        self.V.append(Vt)
        self._dirty = True

    # ------------------------------------------------------------------ plan / device objects
    def _bounds(self):
        return dict(xlb=self.xlb, xub=self.xub, ulb=self.ulb, uub=self.uub, dulb=self.dulb, duub=self.duub,
                    sylb=self.sylb, syub=self.syub, silb=self.silb, siub=self.siub,
                    rslb=self.rslb, rilb=self.rilb)

    def compile(self):
        """Condense the declarations into the constant blocks of the device step (once)."""
        if not self._dirty and self._plan is not None:
            return self._plan
        if not self.VJ:
            raise ValueError("no weighted sub-objective declared")
        cos = [V.coeff for V in self.VJ]
        for e in self._ell:                                   # multiplier slots of the ellipsoid rows: weights m.. of the plan
            c0 = e['V'].coeff
            cos.append(SubObjective(name='lam_' + c0.name, cy=np.zeros(self.ny), csy=np.zeros(self.ny),
                                    csi=c0.csi.copy(), cdu=np.zeros(self.nu)))
        self._plan = compile_plan(self.sys, self.N, cos, self._bounds(),
                                  Q_bar=self.Q_bar, vin=self._vin,
                                  rs=np.asarray(self.rslb, float), ri=np.asarray(self.rilb, float),
                                  rs_hi=np.asarray(self.rsub, float), ri_hi=np.asarray(self.riub, float))
        self._eval_plan = None
        self._stepper = None
        self._eval_dp = None
        self._hot_dp = None
        self._dirty = False
        return self._plan

    def _evaluation_plan(self):
        """Second plan whose 'weighted' list is ALL sub-objectives (fObj.F / v_hist / sol['g'])."""
        self.compile()
        if self._eval_plan is None:
            cos = []
            for V in self.V:
                c = V.coeff
                cos.append(SubObjective(name=c.name, cy=c.cy, csy=c.csy, csi=c.csi, cdu=c.cdu))
            ep = compile_plan(self.sys, self.N, cos, {}, Q_bar=self.Q_bar, vin=(),
                              rs=np.asarray(self.rslb, float), ri=np.asarray(self.rilb, float))
            for i, V in enumerate(self.V):                   # terminal rows carry Vt
                if V.coeff.terminal:
                    a, b = ep.row_slices['term']
                    ep.Wrow[a:b, i] = 1.0
            self._eval_plan = ep
        return self._eval_plan

    def engine(self, batch, **kw):
        from .engine import BatchedEngine
        if self._ell:
            raise NotImplementedError("a siN sub-objective over several outputs (quadratic row) is solved by mpc()'s outer "
                                      "multiplier search; the batched closed-loop engine does not carry it")
        return BatchedEngine(self.compile(), batch, **kw)

    def _get_stepper(self):
        from .engine import HostStepper
        self.compile()
        if self._stepper is None:
            self._stepper = HostStepper(self._plan, 1)
        return self._stepper

    def _eval_all(self, x, w, ysp):
        """All len(V) sub-objective values at explicit w (device: sihmpc_subobjectives)."""
        import torch
        from .engine import DevicePlan
        ep = self._evaluation_plan()
        if self._eval_dp is None:
            self._eval_dp = DevicePlan(ep)
        dp = self._eval_dp
        nw = self.N * self.nu + 2 * self.ny
        xv = torch.as_tensor(_vec(x, self.nx, "x0")).reshape(1, -1).to(dp.device)
        wv = torch.as_tensor(_vec(w, nw, "var")).reshape(1, -1).to(dp.device)
        yv = torch.as_tensor(_vec(ysp, self.ny, "ysp")).reshape(1, -1).to(dp.device)
        out = torch.zeros(1, len(self.V), dtype=torch.float64, device=dp.device)
        st = torch.cuda.current_stream(dp.device).cuda_stream
        from . import _capi
        with torch.cuda.device(dp.device):
            _capi.check(dp.lib.sihmpc_subobjectives(dp.handle, C.c_void_p(xv.data_ptr()), C.c_void_p(wv.data_ptr()),
                                                    C.c_void_p(yv.data_ptr()), 1, C.c_void_p(out.data_ptr()),
                                                    C.c_void_p(st)), "sihmpc_subobjectives")
        return out.cpu().numpy().reshape(-1)

    def _predict(self, x, u, dU):
        import torch
        from . import _capi
        st_ = self._get_stepper()
        dp = st_.dp
        N, nx, ny, nu = self.N, self.nx, self.ny, self.nu
        dev = dp.device
        xv = torch.as_tensor(x).reshape(1, -1).to(dev)
        uv = torch.as_tensor(u).reshape(1, -1).to(dev)
        dv = torch.as_tensor(dU).reshape(1, -1).to(dev)
        xp = torch.zeros(1, N * nx, dtype=torch.float64, device=dev)
        yp = torch.zeros(1, N * ny, dtype=torch.float64, device=dev)
        up = torch.zeros(1, N * nu, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _capi.check(dp.lib.sihmpc_predict(dp.handle, C.c_void_p(xv.data_ptr()), C.c_void_p(uv.data_ptr()),
                                              C.c_void_p(dv.data_ptr()), 1, C.c_void_p(xp.data_ptr()),
                                              C.c_void_p(yp.data_ptr()), C.c_void_p(up.data_ptr()),
                                              C.c_void_p(st)), "sihmpc_predict")
        return xp.cpu().numpy().reshape(-1), yp.cpu().numpy().reshape(-1), up.cpu().numpy().reshape(-1)

    # ------------------------------------------------------------------ step (ctl.py:567-604)
    def mpc(self, x0, ySP, w0, u0, pesos, lam_w0=[], lam_g0=[], ViN_ant=None):
        from . import _capi
        hs = self._get_stepper()
        nx, nu, ny, N = self.nx, self.nu, self.ny, self.N
        if ViN_ant is None:                                   # ctl.py:571-572
            ViN_ant = self.ViNant
        vin = _vec(ViN_ant, len(self.ViNant), "ViN_ant") if len(self.ViNant) else np.zeros(0)
        l = len(self.VJ)
        if len(pesos) != l:                                   # ctl.py:575-577
            pesos = self.init_pesos()
        pesos = _vec(pesos, l, "pesos")
        x = _vec(x0, nx, "x0"); u = _vec(u0, nu, "u0"); ysp = _vec(ySP, ny, "ySP")
        # limits of the plan's contraction rows: a single-index siN row takes its own ViN_ant, the box of an
        # ellipsoid takes the ellipsoid's
        pvin = np.full(len(self._vin), np.inf)
        for i, (kind, k) in enumerate(self._vin_of):
            for slot in ([k] if kind == 'row' else self._ell[k]['slots']):
                pvin[slot] = vin[i]
        pes_ext = np.concatenate([pesos, np.zeros(len(self._ell))])

        def solve(pes):
            hs.x.numpy()[0] = x
            hs.u.numpy()[0] = u
            hs.ysp.numpy()[0] = ysp
            hs.pesos.numpy()[0] = pes
            if len(pvin):
                hs.vinant.numpy()[0, :len(pvin)] = pvin
            hs.step(flags=_capi.F_UPDATE_VIN, want_lam=True)
            return hs.w_opt.numpy()[0].copy()

        w = solve(pes_ext)
        for k, e in enumerate(self._ell):
            rho = vin[[i for i, (kd, kk) in enumerate(self._vin_of) if kd == 'ell' and kk == k][0]]
            w = self._ellipsoid_search(e, l + k, rho, pes_ext, solve, w)
        J = float(hs.J.numpy()[0])
        for k, e in enumerate(self._ell):                     # sol['J'] is the cost, without the multiplier term
            si = w[N * nu + ny:][e['inds']]
            J -= pes_ext[l + k] * float(np.dot(e['q'], si * si))
        self.status = int(hs.status.numpy()[0])
        if self._plan.fixed_rows and fixed_row_violation(self._plan, x, ysp, pvin if len(pvin) else None) > 1e-8:
            self.status |= _capi.ST_INFEASIBLE            # a bound the moves cannot influence is violated at x0
        self.n_active = int(hs.n_active.numpy()[0])
        x_pred, y_pred, u_pred = self._predict(x, u, w[:N * nu])
        vals = self._eval_all(x, w, ysp)
        XN = x_pred[-nx:]
        res1 = XN[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi] - w[N * nu + ny:]
        res2 = XN[:self.nxs] - ysp - w[N * nu:N * nu + ny]
        gvec = np.concatenate([x_pred, u_pred, res1, res2, vals])
        lam_w, lam_g = self._duals(hs.lam.numpy()[0], x, ysp, w, pes_ext, pvin)
        sol = {
            'J': DM(J),
            'du_opt': DM(w[:N * nu].reshape(N, nu).T),        # nu x N (ctl.py:510-515)
            'w_opt': DM(w),
            'lam_w': DM(lam_w),                               # ctl.py:525-529, casadi's sign convention
            'lam_g': DM(lam_g),
            'g': DM(gvec),
            'x_pred': DM(x_pred), 'y_pred': DM(y_pred), 'u_pred': DM(u_pred),
        }
        if len(vin):                                          # ctl.py:581-584: ViNant[i] = F_ViN[i](warm start)
            newv = hs.vinant.numpy()[0, :len(pvin)]           # q siN_next^2 per plan slot, from the device
            for i, (kind, k) in enumerate(self._vin_of):
                if kind == 'row':
                    self.ViNant[i] = float(newv[k])
                else:
                    e = self._ell[k]
                    self.ViNant[i] = float(sum(newv[slot] for slot in e['slots']))
        self.du = sol['du_opt'][:, 0].full()                  # ctl.py:595
        self.j_value = J
        self.j_hist.append(self.j_value)
        for i in range(len(self.V)):                          # ctl.py:598-599
            self.V[i].v_hist.append(float(vals[i]))
        for i in range(len(self.VJ)):
            self.VJ[i].peso_hist.append(pesos[i])
        return sol

    def _duals(self, lam_rows, x, ysp, w, pesos, vin):
        """sol['lam_w'] (nw) and sol['lam_g'] (dim_g = N nx + N nu + 2 ny + len(V), the g layout of ctl.py:414-456)
        from the device multipliers of the inequality rows (sihmpc_io.lam), casadi's sign convention (> 0: upper
        bound active, < 0: lower).  The multipliers of the terminal residual rows res1 / res2 (equalities by
        default, ctl.py:441-449) follow from stationarity in siN / syN:  grad_w J + J_g^T lam_g + lam_w = 0.
        Rows 0 <= V_i <= inf are vacuous (multiplier 0; where V_i = 0 the reference's IPOPT returns noise, SURVEY
        8c); a contraction row V_siN <= ViN_ant is enforced as |siN| <= sqrt(ViN_ant/q), its multiplier mu maps to
        lam = mu / (2 q |siN|)."""
        p = self._plan
        N, nu, ny, nx = self.N, self.nu, self.ny, self.nx
        nvm, nv = N * nu, p.nv + p.n_extra
        lam_w = np.zeros(nvm + 2 * ny)
        lam_g = np.zeros(N * nx + N * nu + 2 * ny + len(self.V))
        lam_w[:nvm] = lam_rows[:nvm]                                            # move box
        lam_g[N * nx:N * nx + nvm] = lam_rows[nv:nv + nvm]                      # U_pred rows, k-major like the reference
        o_res1, o_res2, o_V = N * nx + nvm, N * nx + nvm + ny, N * nx + nvm + 2 * ny
        dense = lam_rows[2 * nv:2 * nv + p.ncd]
        syN, siN = w[nvm:nvm + ny], w[nvm + ny:nvm + 2 * ny]
        lam_rho = {}
        vin_terms = []                                          # (output, d row / d siN * multiplier) of contraction rows
        for r, kind in enumerate(p.row_kind):
            if kind[0] == 'x':
                lam_g[kind[1] * nx + kind[2]] = dense[r]
            elif kind[0] == 'sy':
                lam_w[nvm + kind[1]] = dense[r]
            elif kind[0] == 'si':
                lam_w[nvm + ny + kind[1]] = dense[r]
            elif kind[0] == 'rho':
                lam_rho[kind[1]] = dense[r]
            elif kind[0] == 'vin':
                slot = kind[1]
                owner = next((i for i, (kd, k) in enumerate(self._vin_of) if kd == 'row' and k == slot), None)
                ind, q = self._vin[slot]
                if owner is not None and dense[r] != 0.0 and abs(siN[ind]) > 0:
                    V = next(V for V in self.V if V.coeff.vin_slot == owner)
                    lam_g[o_V + self.V.index(V)] = abs(dense[r]) / (2.0 * q * abs(siN[ind]))
                    vin_terms.append((ind, 2.0 * q * siN[ind] * lam_g[o_V + self.V.index(V)]))
                elif owner is None:
                    vin_terms.append((ind, dense[r]))           # box row of an ellipsoid: d siN_ind / d siN_ind = 1
        # stationarity in the slacks: d J / d syN_i - lam_res2_i + lam_w[syN_i] = 0, and the same for siN with res1 and
        # the contraction row (d V_siN / d siN_i = 2 q siN_i)
        om = p.Wrow @ pesos + p.w0row
        Y = (p.Myx @ x + p.Myu[:, :nvm] @ w[:nvm]).reshape(N, ny)
        r_y = Y - ysp[None, :] - syN[None, :] - p.tau[:, None] * siN[None, :]   # ctl.py:174
        om_y = om[:N * ny].reshape(N, ny)
        a, b = p.row_slices['syN']; om_sy = om[a:b]
        a, b = p.row_slices['siN']; om_si = om[a:b]
        dJ_sy = -2.0 * np.sum(om_y * r_y, axis=0) + 2.0 * om_sy * syN
        dJ_si = -2.0 * np.sum(om_y * r_y * p.tau[:, None], axis=0) + 2.0 * om_si * siN
        lam_res2 = dJ_sy + lam_w[nvm:nvm + ny]
        lam_res1 = dJ_si + lam_w[nvm + ny:]                   # (om carries the ellipsoid multipliers: pesos m..)
        for ind, t in vin_terms:
            lam_res1[ind] += t
        for k, e in enumerate(self._ell):
            lam_g[o_V + self.V.index(e['V'])] = pesos[len(self.VJ) + k]
        lam_g[o_res1:o_res1 + ny] = lam_res1
        lam_g[o_res2:o_res2 + ny] = lam_res2
        return lam_w, lam_g

    def _ellipsoid_search(self, e, slot, rho, pes_ext, solve, w):
        """Multiplier of the quadratic row  V_e(w) = sum_j q_j siN_j^2 <= rho  (ctl.py:243-271): for a fixed
        multiplier the problem is the QP of J + lam V_e (one device solve), V_e falls monotonically in lam; bracket,
        then Brent on V_e - rho to rounding, keeping the feasible end of the last bracket.  KKT of the QCQP:
        grad J + lam grad V_e + (rows) = 0, lam >= 0, lam (V_e - rho) = 0."""
        from scipy.optimize import brentq
        nvm, ny = self.N * self.nu, self.ny
        q, inds = np.asarray(e['q']), e['inds']

        def val(wv):
            si = wv[nvm + ny:nvm + 2 * ny][inds]
            return float(np.dot(q, si * si))

        if not np.isfinite(rho) or val(w) <= rho * (1.0 + 1e-12):
            return w
        if rho <= 0.0:
            return w            # the box rows already pin every siN_j at 0
        seen = {}

        def phi(lam):
            pes_ext[slot] = lam
            seen[lam] = solve(pes_ext)
            return val(seen[lam]) - rho

        hi = 1.0
        while phi(hi) > 0.0:
            hi *= 8.0
            if hi > 1e16:
                raise RuntimeError("the siN ellipsoid row cannot be met")
        lo = 0.0 if hi == 1.0 else hi / 8.0
        lam = brentq(phi, lo, hi, xtol=1e-300, rtol=8.9e-16, maxiter=200)
        feas = [l for l in seen if val(seen[l]) <= rho * (1.0 + 1e-12)]
        lam = min(feas, key=lambda l: abs(l - lam))
        pes_ext[slot] = lam
        return solve(pes_ext)          # leaves J, status, multipliers of THIS solve in the stepper's buffers

    def warmStart(self, sol, ysp):                            # ctl.py:534-564
        w0 = np.asarray(sol['w_opt']).reshape(-1)
        ny, nu, nx = self.ny, self.nu, self.nx
        dustart = np.concatenate([w0[:-2 * ny][nu:], np.zeros(nu)])
        xN = np.asarray(sol['x_pred']).reshape(-1)[-nx:]
        res = self.dynF(x0=xN, u0=np.zeros(nu), du0=np.zeros(nu))
        Xknext = res['xkp1'].full().reshape(-1)
        siNnext = Xknext[self.nxs + self.nxd:self.nxs + self.nxd + self.nxi]
        syNnext = w0[-2 * ny:-ny]                             # "Opcao 1"
        return np.concatenate([dustart, syNnext, siNnext])

    def _vj_values(self, x, w_start, ysp):
        vals = self._eval_all(x, w_start, ysp)
        return np.array([vals[self.V.index(V)] for V in self.VJ])

    def satWeights(self, x, u, w_start, ysp):                 # ctl.py:607-617
        Vn = self._vj_values(x, w_start, ysp)
        gam = np.array([V.gamma for V in self.VJ], dtype=float)
        return 1 / (gam - np.clip(Vn, 0, 0.99 * gam))

    def satWeights2(self, x, u, w_start, ysp, alfa=0):        # ctl.py:619-636
        Vnext = self._vj_values(x, w_start, ysp)
        gamma = np.array([V.gamma for V in self.VJ], dtype=float)
        s = Vnext / gamma
        smax = max(s)
        gamma = np.maximum(gamma, gamma * smax + 1e-6)
        new_pesos = 1 / (gamma - Vnext)
        old = self.pesos if len(self.pesos) == len(new_pesos) else np.zeros(len(new_pesos))
        self.pesos = alfa * np.asarray(old) + (1 - alfa) * new_pesos
        self._s = s
        return self.pesos, gamma

    # ------------------------------------------------------------------ plots (host only, ctl.py:638-701)
    def _plt(self):
        import matplotlib.pyplot as plt
        return plt

    def plotPesos(self, t):
        plt = self._plt(); fig = plt.figure(); fig.suptitle("Weights")
        nw = len(self.pesos); y = int(round(np.sqrt(nw) + 0.5)); x = int(round(nw / y + 0.5))
        for i in range(nw):
            plt.subplot(x, y, i + 1); plt.step(t, self.VJ[i].peso_hist, label=self.VJ[i].weight.name())
            plt.legend(); plt.grid()
        plt.show()

    def plotPesosNormalizados(self, t):
        plt = self._plt(); fig = plt.figure(); fig.suptitle("Normalized Weights")
        for i in range(len(self.pesos)):
            plt.step(t, np.array(self.VJ[i].peso_hist) * self.VJ[i].gamma, label='n' + self.VJ[i].weight.name())
            plt.legend()
        plt.show()

    def plotJ(self, t):
        plt = self._plt(); fig = plt.figure(); fig.suptitle("Total Cost"); plt.plot(t, self.j_hist); plt.show()

    def plotJi(self, t):
        plt = self._plt(); fig = plt.figure(); fig.suptitle("Weighted Local Costs")
        l = len(self.VJ); y = int(round(np.sqrt(l) + 0.5)); x = int(round(l / y + 0.5))
        for i in range(l):
            plt.subplot(x, y, i + 1)
            plt.step(t, np.array(self.VJ[i].peso_hist) * np.array(self.VJ[i].v_hist),
                     label=self.VJ[i].weight.name() + '*' + self.VJ[i].name)
            plt.legend()
        plt.show()

    def plotV(self, t):
        plt = self._plt(); fig = plt.figure(); fig.suptitle("Local Costs")
        l = len(self.V); y = int(round(np.sqrt(l) + 0.5)); x = int(round(l / y + 0.5))
        for i in range(l):
            plt.subplot(x, y, i + 1); plt.step(t, self.V[i].v_hist, label=self.V[i].name); plt.legend()
        plt.show()
