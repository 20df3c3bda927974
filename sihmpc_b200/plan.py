# -*- coding: utf-8 -*-
"""Plan compiler: controller declaration -> constant FP64 blocks for the device step.

Done once per controller on the host (NumPy FP64), never per step.  It replaces what the
reference redoes on EVERY call of ``mpc()``: ``_MPC()`` -> ``_OptimProbl()`` ->
``casadi.nlpsol`` graph construction and AD (``/root/reference/sihmpc/ihmpccontroller.py:
569,480-531,374-477``).

Math (SURVEY.md section 8a "restated math").  Decision vector of the reference
``w = [dU ; syN ; siN]`` (ctl.py:399-412).  With the terminal residual rows held as
equalities (defaults ``rs* = ri* = 0``, ctl.py:50-53,441-449) the slacks are affine in dU:

    syN = xs(N) - ysp - rs        siN = xi(N) - ri

and every term of J is a weighted square of an affine residual.  Residual stack (R rows):

    y-row (k,i) : Y_pred[k][i] - ysp_i - syN_i - (k+1-N) Ts siN_i          (ctl.py:174)
    syN-row i   : syN_i                                                      (ctl.py:224)
    siN-row i   : siN_i                                                      (ctl.py:247)
    term-row a  : (Lbar^T xd(N))_a ,  Q_bar = Lbar Lbar^T                    (ctl.py:338-340)

    r(dU) = E dU + e ,   e = Px x + Py ysp + e0
    J     = sum_r omega_r r_r^2 + sum_j omega_du_j dU_j^2
    omega = Wrow pesos + w0row ,  omega_du = Wdu pesos

so  H = Ht + sum_i pesos_i HJ_i + 2 diag(omega_du)  and  f = 2 E^T (omega o e).
Inequality rows over dU: move box (ctl.py:399-403), cumulative input box (ctl.py:420-424)
and "dense" rows (state box ctl.py:414-418, slack boxes ctl.py:405-412, siN contraction
ctl.py:265-268).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np
from scipy.linalg import solve_discrete_lyapunov

INF = np.inf


@dataclass
class SubObjective:
    """Coefficient form of one reference fObj (ctl.py:116-149).

    V = sum_i cy[i]  * sum_k yrow(k,i)^2 + sum_i csy[i] * syN_i^2
      + sum_i csi[i] * siN_i^2           + sum_j cdu[j] * sum_k dU_k[j]^2   (+ Vt if terminal)
    """
    name: str
    cy: np.ndarray
    csy: np.ndarray
    csi: np.ndarray
    cdu: np.ndarray
    terminal: bool = False
    gamma: float = INF
    peso: float = 0.0
    vin_slot: int = -1          # >=0: upper limit is the contraction symbol ViN_ant[slot]


@dataclass
class Plan:
    nx: int; nu: int; ny: int; nd: int; N: int; nv: int; m: int; mh: int
    R: int; ncd: int; n_vin: int; Ts: float
    # prediction / residual blocks
    E: np.ndarray; Px: np.ndarray; Py: np.ndarray; e0: np.ndarray
    Wrow: np.ndarray; w0row: np.ndarray; Wdu: np.ndarray
    HJ: np.ndarray; hj_index: np.ndarray; Ht: np.ndarray
    # next-state evaluation
    Myx: np.ndarray; Myu: np.ndarray; Mi1x: np.ndarray; Mi1u: np.ndarray
    tau: np.ndarray
    gamma: np.ndarray; peso0: np.ndarray
    # constraints
    dulb: np.ndarray; duub: np.ndarray; ulb: np.ndarray; uub: np.ndarray
    Gc: np.ndarray; Gcx: np.ndarray; Gcy: np.ndarray; gc0: np.ndarray
    Lc: np.ndarray; Uc: np.ndarray; gc_vin: np.ndarray; gc_q: np.ndarray
    vin_row: np.ndarray; vin_q: np.ndarray
    # plant (nominal), CSR over [x ; du]
    pl_rowptr: np.ndarray; pl_col: np.ndarray; pl_val: np.ndarray
    # dense model (host-side prediction outputs, C ABI predict call)
    A: np.ndarray; B: np.ndarray; C: np.ndarray; D: np.ndarray
    Phi: np.ndarray = field(default=None, repr=False)      # (N*nx, nx)
    Gam: np.ndarray = field(default=None, repr=False)      # (N*nx, nv)
    Q_bar: np.ndarray = None
    row_slices: dict = None
    # terminal residual RANGES (ctl.py:441-449 with lb != ub): one extra decision variable per ranged residual,
    # appended after the N*nu moves (every block with a `nv` dimension is nv + n_extra wide); extra[k] =
    # ('s'|'i', output index)
    n_extra: int = 0
    extra: tuple = ()
    # inequality rows the moves cannot influence (zero normal): not sent to the device, kept here so that the caller
    # can still see them violated -- fixed_rows[k] = (Gcx, Gcy, gc0, lo, hi, vin_slot, q); see fixed_row_violation()
    fixed_rows: tuple = ()
    # origin of every dense row (same order as Gc): ('x', k, c) state box on X_pred[k][c], ('sy', i) / ('si', i) slack
    # boxes, ('vin', slot) siN contraction, ('rho', c) range of a ranged terminal residual -- used to scatter the
    # device multipliers into sol['lam_w'] / sol['lam_g'] (ihmpccontroller.py:525-529)
    row_kind: tuple = ()


def plant_csr(A, B, pattern=None):
    """CSR of [A | B].  ``pattern`` (bool, same shape) fixes the sparsity so that per-scenario
    plants (same dead times => same structure) share rowptr/col and differ in values only."""
    AB = np.hstack([A, B])
    if pattern is None:
        pattern = AB != 0.0
    rowptr = [0]; col = []; val = []
    for i in range(AB.shape[0]):
        js = np.nonzero(pattern[i])[0]
        col.extend(js.tolist()); val.extend(AB[i, js].tolist())
        rowptr.append(len(col))
    return (np.array(rowptr, dtype=np.int32), np.array(col, dtype=np.int32),
            np.array(val, dtype=np.float64), pattern)


def terminal_weight(sys, Qt):
    """Q_bar of ctl.py:338-339:  F^T Q F - Q + F^T Psi^T Qt Psi F = 0."""
    F, Psi = np.asarray(sys.F, float), np.asarray(sys.Psi, float)
    Q_lyap = F.T @ Psi.T @ Qt @ Psi @ F
    return solve_discrete_lyapunov(F.T, Q_lyap)


def _psd_factor(Q):
    """Lbar with Q = Lbar Lbar^T for a symmetric PSD matrix (eigen form; Q_bar may be singular).
    Q_bar of an OPOM model is block diagonal (one block per output: F is (block) diagonal and Psi^T Qt Psi couples
    only the modes of one output), so the factor is taken block by block and its columns -- the terminal residual
    rows -- stay local to one output's modes: the tile-CSR products of the device skip the zero tiles (an eigen
    factor of the whole matrix orders the columns by eigenvalue and scatters every 8-row tile over all outputs)."""
    from scipy.sparse.csgraph import connected_components
    Q = 0.5 * (Q + Q.T)
    n = Q.shape[0]
    if n == 0:
        return np.zeros((0, 0))
    scale = np.max(np.abs(Q))
    ncomp, lab = connected_components(np.abs(Q) > 1e-14 * scale, directed=False) if scale > 0 else (1, np.zeros(n, int))
    L = np.zeros((n, n))
    col = 0
    for comp in sorted(range(ncomp), key=lambda k: int(np.nonzero(lab == k)[0][0])):
        idx = np.nonzero(lab == comp)[0]
        lam, V = np.linalg.eigh(Q[np.ix_(idx, idx)])
        order = np.argsort(-lam)
        lam = np.where(lam[order] > 0, lam[order], 0.0)
        L[np.ix_(idx, np.arange(col, col + idx.size))] = V[:, order] * np.sqrt(lam)[None, :]
        col += idx.size
    return L


def bounds_rs(kind, i, rs_hi, ri_hi, bounds):
    """(lb, ub) of the ranged terminal residual (kind 's': res2 = xs(N)-ysp-syN, 'i': res1 = xi(N)-siN)."""
    lo = np.asarray(bounds['rslb' if kind == 's' else 'rilb'], float).reshape(-1)[i]
    hi = (rs_hi if kind == 's' else ri_hi)[i]
    return float(lo), float(hi)


def compile_plan(sys, N, VJ, bounds, Q_bar=None, vin=(), rs=None, ri=None, rs_hi=None, ri_hi=None):
    """Build the Plan.

    sys    : OPOM-like object (attributes A,B,C,D,Ts,nx,nu,ny,nd,nz).
    VJ     : list of SubObjective that enter J with a weight (ctl.py:189-190), in order.
    bounds : dict with the constructor's bound vectors (ctl.py:38-53).
    Q_bar  : terminal weight (nd x nd) or None when no terminal objective was set.
    vin    : list of (output index, q) for the siN contraction rows, slot order (ctl.py:265-269).
    rs, ri : lower ends of the terminal residual bounds rslb, rilb (ctl.py:50-53); rs_hi, ri_hi the upper ends.
             Where they coincide (the default, 0) the residual is an equality and syN / siN are eliminated:
             syN = xs(N) - ysp - rs, siN = xi(N) - ri.  Where lb != ub the residual rho in [lb, ub] becomes a
             decision variable of its own: syN = xs(N) - ysp - rho_s, siN = xi(N) - rho_i (ctl.py:441-449), appended
             after the moves with its range as a dense row.
    """
    A = np.asarray(sys.A, float); B = np.asarray(sys.B, float)
    C = np.asarray(sys.C, float); D = np.asarray(sys.D, float)
    nx, nu, ny, nd = int(sys.nx), int(sys.nu), int(sys.ny), int(sys.nd)
    N = int(N); nv = N * nu; Ts = float(sys.Ts)
    s0, d0, i0 = 0, ny, ny + nd
    rs = np.zeros(ny) if rs is None else np.asarray(rs, float).copy()
    ri = np.zeros(ny) if ri is None else np.asarray(ri, float).copy()
    rs_hi = rs.copy() if rs_hi is None else np.asarray(rs_hi, float)
    ri_hi = ri.copy() if ri_hi is None else np.asarray(ri_hi, float)
    extra = [('s', i) for i in range(ny) if rs_hi[i] != rs[i]] + [('i', i) for i in range(ny) if ri_hi[i] != ri[i]]
    n_extra = len(extra)
    for kind, i in extra:                    # a ranged residual carries no constant offset: it is the variable
        (rs if kind == 's' else ri)[i] = 0.0
    nvm = nv                                 # moves

    # closed-form prediction matrices  X_pred[k] = A^{k+1} x + sum_{j<=k} A^{k-j} B dU_j
    Apow = [np.eye(nx)]
    for _ in range(N + 1):
        Apow.append(A @ Apow[-1])
    Phi = np.vstack([Apow[k + 1] for k in range(N)])
    Gam = np.zeros((N * nx, nv))
    for k in range(N):
        for j in range(k + 1):
            Gam[k * nx:(k + 1) * nx, j * nu:(j + 1) * nu] = Apow[k - j] @ B
    Myx = np.vstack([C @ Phi[k * nx:(k + 1) * nx] for k in range(N)])
    Myu = np.vstack([C @ Gam[k * nx:(k + 1) * nx] for k in range(N)])
    for k in range(N):
        Myu[k * ny:(k + 1) * ny, k * nu:(k + 1) * nu] += D
    PhiN, GamN = Phi[(N - 1) * nx:], Gam[(N - 1) * nx:]
    xsx, xsu = PhiN[s0:s0 + ny], GamN[s0:s0 + ny]
    xdx, xdu = PhiN[d0:d0 + nd], GamN[d0:d0 + nd]
    xix, xiu = PhiN[i0:i0 + ny], GamN[i0:i0 + ny]
    # xi(N+1) with a zero move appended (warm start, ctl.py:543-548)
    Mi1x = (A @ PhiN)[i0:i0 + ny]
    Mi1u = (A @ GamN)[i0:i0 + ny]

    tau = np.array([(k + 1 - N) * Ts for k in range(N)], dtype=float)

    # residual rows in reduced form
    R = N * ny + 2 * ny + nd
    E = np.zeros((R, nv)); Px = np.zeros((R, nx)); Py = np.zeros((R, ny)); e0 = np.zeros(R)
    for k in range(N):
        for i in range(ny):
            r = k * ny + i
            # Y - ysp - (xs(N) - ysp - rs) - tau (xi(N) - ri)
            E[r] = Myu[r] - xsu[i] - tau[k] * xiu[i]
            Px[r] = Myx[r] - xsx[i] - tau[k] * xix[i]
            e0[r] = rs[i] + tau[k] * ri[i]
    o_sy, o_si, o_t = N * ny, N * ny + ny, N * ny + 2 * ny
    for i in range(ny):
        E[o_sy + i] = xsu[i]; Px[o_sy + i] = xsx[i]; Py[o_sy + i, i] = -1.0; e0[o_sy + i] = -rs[i]
        E[o_si + i] = xiu[i]; Px[o_si + i] = xix[i]; e0[o_si + i] = -ri[i]
    if Q_bar is not None:
        Lbar = _psd_factor(np.asarray(Q_bar, float))
        E[o_t:o_t + nd] = Lbar.T @ xdu
        Px[o_t:o_t + nd] = Lbar.T @ xdx
    if n_extra:
        # widen every move-indexed block by the residual variables: r_y += rho_s + tau rho_i, syN -= rho_s, siN -= rho_i
        nv = nvm + n_extra
        Ex = np.zeros((R, n_extra))
        for c, (kind, i) in enumerate(extra):
            for k in range(N):
                Ex[k * ny + i, c] = 1.0 if kind == 's' else tau[k]
            Ex[(o_sy if kind == 's' else o_si) + i, c] = -1.0
        E = np.hstack([E, Ex])
        Myu = np.hstack([Myu, np.zeros((N * ny, n_extra))])
        Mi1u = np.hstack([Mi1u, np.zeros((ny, n_extra))])
        Gam = np.hstack([Gam, np.zeros((N * nx, n_extra))])

    m = len(VJ)
    Wrow = np.zeros((R, m)); w0row = np.zeros(R); Wdu = np.zeros((nv, m))
    for i, V in enumerate(VJ):
        assert not V.terminal
        for ind in range(ny):
            for k in range(N):
                Wrow[k * ny + ind, i] = V.cy[ind]
            Wrow[o_sy + ind, i] = V.csy[ind]
            Wrow[o_si + ind, i] = V.csi[ind]
        for k in range(N):
            for j in range(nu):
                Wdu[k * nu + j, i] = V.cdu[j]                # (rows of the residual variables stay zero)
    if Q_bar is not None:
        w0row[o_t:o_t + nd] = 1.0

    hj_index = [i for i in range(m) if np.any(Wrow[:, i] != 0.0)]
    HJ = np.zeros((len(hj_index), nv, nv))
    for a, i in enumerate(hj_index):
        HJ[a] = 2.0 * E.T @ (Wrow[:, i][:, None] * E)
    Ht = 2.0 * E.T @ (w0row[:, None] * E)

    # ---- inequality rows
    g = lambda k, dflt: np.asarray(bounds.get(k, dflt), float).reshape(-1)
    dulb = g('dulb', np.full(nu, -INF)); duub = g('duub', np.full(nu, INF))
    ulb = g('ulb', np.full(nu, -INF)); uub = g('uub', np.full(nu, INF))
    xlb = g('xlb', np.full(nx, -INF)); xub = g('xub', np.full(nx, INF))
    sylb = g('sylb', np.full(ny, -INF)); syub = g('syub', np.full(ny, INF))
    silb = g('silb', np.full(ny, -INF)); siub = g('siub', np.full(ny, INF))

    rows = []   # (Gc, Gcx, Gcy, g0, lo, hi, vin_slot, q)
    kinds = []
    for k in range(N):
        for c in range(nx):
            if np.isfinite(xlb[c]) or np.isfinite(xub[c]):
                rows.append((Gam[k * nx + c], Phi[k * nx + c], np.zeros(ny), 0.0, xlb[c], xub[c], -1, 0.0))
                kinds.append(('x', k, c))
    for i in range(ny):
        if np.isfinite(sylb[i]) or np.isfinite(syub[i]):
            rows.append((E[o_sy + i], Px[o_sy + i], Py[o_sy + i], e0[o_sy + i], sylb[i], syub[i], -1, 0.0))
            kinds.append(('sy', i))
    for i in range(ny):
        if np.isfinite(silb[i]) or np.isfinite(siub[i]):
            rows.append((E[o_si + i], Px[o_si + i], Py[o_si + i], e0[o_si + i], silb[i], siub[i], -1, 0.0))
            kinds.append(('si', i))
    vin_row = []; vin_q = []
    for slot, (ind, q) in enumerate(vin):
        vin_row.append(o_si + ind); vin_q.append(float(q))
        if q > 0:
            rows.append((E[o_si + ind], Px[o_si + ind], Py[o_si + ind], e0[o_si + ind], -INF, INF, slot, float(q)))
            kinds.append(('vin', slot))
    for c, (kind, i) in enumerate(extra):                       # lb <= rho <= ub
        unit = np.zeros(nv); unit[nvm + c] = 1.0
        lo_, hi_ = (bounds_rs(kind, i, rs_hi, ri_hi, bounds))
        rows.append((unit, np.zeros(nx), np.zeros(ny), 0.0, lo_, hi_, -1, 0.0))
        kinds.append(('rho', c))
    # rows whose normal vanishes cannot be influenced by dU: vacuous or infeasible whatever
    # the solver does (non-integrating plants: xi(N) = xi(0)); they are not sent to the device.
    keep = [np.linalg.norm(r[0]) > 1e-13 * max(1.0, np.linalg.norm(r[1])) for r in rows]
    fixed_rows = tuple((r[1], r[2], r[3], r[4], r[5], r[6], r[7]) for r, k in zip(rows, keep)
                       if not k and (np.isfinite(r[4]) or np.isfinite(r[5]) or r[6] >= 0))
    if any(np.isfinite(fr[3]) or np.isfinite(fr[4]) for fr in fixed_rows):
        import warnings
        warnings.warn("%d bounded row(s) do not depend on the moves (e.g. silb/siub or xlb/xub on a state the inputs "
                      "cannot reach): they are checked on the host (SIHMPC_ST_INFEASIBLE), not enforced by the solver"
                      % sum(1 for fr in fixed_rows if np.isfinite(fr[3]) or np.isfinite(fr[4])), stacklevel=2)
    rows = [r for r, k in zip(rows, keep) if k]
    kinds = [kd for kd, k in zip(kinds, keep) if k]
    ncd = len(rows)
    Gc = np.array([r[0] for r in rows]).reshape(ncd, nv)
    Gcx = np.array([r[1] for r in rows]).reshape(ncd, nx)
    Gcy = np.array([r[2] for r in rows]).reshape(ncd, ny)
    gc0 = np.array([r[3] for r in rows], dtype=float)
    Lc = np.array([r[4] for r in rows], dtype=float)
    Uc = np.array([r[5] for r in rows], dtype=float)
    gc_vin = np.array([r[6] for r in rows], dtype=np.int32)
    gc_q = np.array([r[7] for r in rows], dtype=float)

    rowptr, col, val, _ = plant_csr(A, B)

    return Plan(n_extra=n_extra, extra=tuple(extra), fixed_rows=fixed_rows, row_kind=tuple(kinds),
                nx=nx, nu=nu, ny=ny, nd=nd, N=N, nv=nvm, m=m, mh=len(hj_index), R=R, ncd=ncd,
                n_vin=len(vin), Ts=Ts, E=E, Px=Px, Py=Py, e0=e0, Wrow=Wrow, w0row=w0row, Wdu=Wdu,
                HJ=HJ, hj_index=np.array(hj_index, dtype=np.int32), Ht=Ht,
                Myx=Myx, Myu=Myu, Mi1x=Mi1x, Mi1u=Mi1u, tau=tau,
                gamma=np.array([V.gamma for V in VJ], dtype=float),
                peso0=np.array([V.peso for V in VJ], dtype=float),
                dulb=dulb, duub=duub, ulb=ulb, uub=uub,
                Gc=Gc, Gcx=Gcx, Gcy=Gcy, gc0=gc0, Lc=Lc, Uc=Uc, gc_vin=gc_vin, gc_q=gc_q,
                vin_row=np.array(vin_row, dtype=np.int32), vin_q=np.array(vin_q, dtype=float),
                pl_rowptr=rowptr, pl_col=col, pl_val=val,
                A=A, B=B, C=C, D=D, Phi=Phi, Gam=Gam, Q_bar=Q_bar,
                row_slices={'y': (0, N * ny), 'syN': (o_sy, o_sy + ny), 'siN': (o_si, o_si + ny),
                            'term': (o_t, o_t + nd)})


def fixed_row_violation(plan: Plan, x, ysp, vinant=None) -> float:
    """Largest violation of the rows the moves cannot influence (Plan.fixed_rows) at state x: the reference's IPOPT
    would report such a problem infeasible (ctl.py:414-418,405-412 with a state or slack the inputs cannot reach);
    here the step itself ignores them and the caller raises SIHMPC_ST_INFEASIBLE when this is > 1e-8."""
    x = np.asarray(x, float).reshape(-1); ysp = np.asarray(ysp, float).reshape(-1)
    worst = 0.0
    for gx, gy, g0, lo, hi, slot, q in plan.fixed_rows:
        v = float(gx @ x + gy @ ysp + g0)
        if slot >= 0:
            vin = np.inf if vinant is None else float(np.asarray(vinant, float).reshape(-1)[slot])
            b = np.sqrt(max(vin, 0.0) / q) if np.isfinite(vin) and q > 0 else np.inf
            lo, hi = -b, b
        worst = max(worst, lo - v, v - hi)
    return worst


# ---------------------------------------------------------------------------------------
# Host mirror of what the device assembles from a Plan (used to unit-test the compiler on a
# CPU-only box; the product step never runs through this).
def assemble_qp(plan: Plan, x, u, ysp, pesos, vinant=None):
    x = np.asarray(x, float).reshape(-1); u = np.asarray(u, float).reshape(-1)
    ysp = np.asarray(ysp, float).reshape(-1); pesos = np.asarray(pesos, float).reshape(-1)
    e = plan.Px @ x + plan.Py @ ysp + plan.e0
    om = plan.Wrow @ pesos + plan.w0row
    omdu = plan.Wdu @ pesos
    H = plan.Ht.copy()
    for a, i in enumerate(plan.hj_index):
        H += pesos[i] * plan.HJ[a]
    H += 2.0 * np.diag(omdu)
    f = 2.0 * plan.E.T @ (om * e)
    c = float(om @ (e * e))
    nv, nu, N = plan.nv, plan.nu, plan.N
    T = np.kron(np.tril(np.ones((N, N))), np.eye(nu))
    pad = np.zeros((nv, getattr(plan, 'n_extra', 0)))        # residual variables: no move / input rows
    G = [np.hstack([np.eye(nv), pad]), np.hstack([T, pad])]
    lo = [np.tile(plan.dulb, N), np.tile(plan.ulb - u, N)]
    hi = [np.tile(plan.duub, N), np.tile(plan.uub - u, N)]
    if plan.ncd:
        off = plan.Gcx @ x + plan.Gcy @ ysp + plan.gc0
        L = plan.Lc.copy(); U = plan.Uc.copy()
        for r in range(plan.ncd):
            s = plan.gc_vin[r]
            if s >= 0:
                v = np.inf if vinant is None else vinant[s]
                b = np.sqrt(max(v, 0.0) / plan.gc_q[r]) if np.isfinite(v) else np.inf
                L[r], U[r] = -b, b
        G.append(plan.Gc); lo.append(L - off); hi.append(U - off)
    return H, f, c, np.vstack(G), np.concatenate(lo), np.concatenate(hi), e, om, omdu
