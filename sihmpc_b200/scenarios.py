# -*- coding: utf-8 -*-
"""Named plants, controllers and synthetic scenario sets of BASELINE.json.

Random draws are generated in fixed chunks of CHUNK scenarios from a Philox stream keyed by
(seed, chunk index), so scenario s gets the same numbers whatever the batch size, the rank
layout or the subset asked for -- the oracle regenerates any subset independently.
"""
from __future__ import annotations

import numpy as np

from .opom import OPOM, TransferFunction
from .plan import plant_csr

CHUNK = 4096
SEED_DISTILLATION = 20261018
SEED_MIMO = 7


# ------------------------------------------------------------------------- plants
WOOD_BERRY = dict(K=[[12.8, -18.9], [6.6, -19.4]], tau=[[16.7, 21.0], [10.9, 14.4]],
                  theta=[[1, 3], [7, 3]])     # example_distillation.py:16-30


def fopdt_system(K, tau, theta, Ts=1.0):
    K = np.asarray(K, float); tau = np.asarray(tau, float); theta = np.asarray(theta, float)
    h = [[TransferFunction([K[i, j]], [tau[i, j], 1.0], delay=theta[i, j]) for j in range(K.shape[1])]
         for i in range(K.shape[0])]
    return OPOM(h, Ts)


def distillation_system():
    """Wood & Berry 2x2 column (example_distillation.py:16-38): nx = 22."""
    return fopdt_system(WOOD_BERRY["K"], WOOD_BERRY["tau"], WOOD_BERRY["theta"], 1.0)


def mimo10_parameters(seed=SEED_MIMO, n=10):
    """Config 4 plant: K_ii~U(1,2), K_ij~N(0,0.3^2), tau~U(5,25), theta~U{0..4}."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    K = rng.normal(0.0, 0.3, size=(n, n))
    K[np.arange(n), np.arange(n)] = rng.uniform(1.0, 2.0, size=n)
    tau = rng.uniform(5.0, 25.0, size=(n, n))
    theta = rng.integers(0, 5, size=(n, n)).astype(float)
    return K, tau, theta


def mimo10_system(seed=SEED_MIMO, n=10):
    K, tau, theta = mimo10_parameters(seed, n)
    return fopdt_system(K, tau, theta, 1.0)


# ------------------------------------------------------------------------- controllers
def declare_distillation(ctrl_cls, sys=None, N=10, **bounds):
    """The controller of example_distillation.py:41-59 on any class with the reference surface."""
    sys = distillation_system() if sys is None else sys
    if not bounds:
        bounds = dict(ulb=[0, 0])
    c = ctrl_cls(sys, N, **bounds)
    c.subObjComposed(y=[0], Q=1, sat=N * 1.0 ** 2)
    c.subObjComposed(y=[1], Q=1, sat=N * 0.5 ** 2)
    c.subObj(du=[0], Q=1, sat=N * 0.15 ** 2)
    c.subObj(du=[1], Q=1, sat=N * 0.15 ** 2)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.subObj(siN=[1], Q=1, addJ=False)
    c.set_terminal_objective()
    return c


def declare_distillation_log(ctrl_cls, sys=None, N=10, **bounds):
    """The log-barrier controller of example_distillation_log.py:43-61 (`c.init_Qt` is referenced there without
    being called, so Qt stays eye(ny))."""
    sys = distillation_system() if sys is None else sys
    if not bounds:
        bounds = dict(ulb=[0, 0])
    c = ctrl_cls(sys, N, **bounds)
    c.subObjComposed(y=[0], Q=1, sat=N * 0.5 ** 2)
    c.subObjComposed(y=[1], Q=1, sat=N * 2.0 ** 2)
    c.subObj(du=[0], Q=1, sat=N * 0.3 ** 2)
    c.subObj(du=[1], Q=1, sat=N * 0.3 ** 2)
    c.subObj(siN=[0], Q=1, addJ=False)
    c.subObj(siN=[1], Q=1, addJ=False)
    return c


def declare_mimo10(ctrl_cls, sys=None, N=5, **bounds):
    """Config 4 controller: one composed y per output, one du per input, siN rows, terminal cost."""
    sys = mimo10_system() if sys is None else sys
    n = sys.ny
    if not bounds:
        bounds = dict(ulb=[-5.0] * sys.nu, uub=[5.0] * sys.nu)
    c = ctrl_cls(sys, N, **bounds)
    for i in range(n):
        c.subObjComposed(y=[i], Q=1, sat=N * 1.0 ** 2)
    for j in range(sys.nu):
        c.subObj(du=[j], Q=1, sat=N * 0.2 ** 2)
    for i in range(n):
        c.subObj(siN=[i], Q=1, addJ=False)
    c.set_terminal_objective()
    return c


# ------------------------------------------------------------------------- scenario draws
def _chunked(seed, first, count, draw):
    """draw(rng, n) -> dict of arrays with leading dim n; concatenated over [first, first+count)."""
    out = None
    c0, c1 = first // CHUNK, (first + count - 1) // CHUNK
    for c in range(c0, c1 + 1):
        rng = np.random.Generator(np.random.Philox(key=[seed, c]))
        d = draw(rng, CHUNK)
        lo = max(first, c * CHUNK) - c * CHUNK
        hi = min(first + count, (c + 1) * CHUNK) - c * CHUNK
        d = {k: v[lo:hi] for k, v in d.items()}
        out = d if out is None else {k: np.concatenate([out[k], d[k]]) for k in d}
    return out


def distillation_scenarios(count, first=0, seed=SEED_DISTILLATION, n_steps=200):
    """Config 3: set-point steps at k>10 and k>100, one output-step disturbance in k_d~U{20..180}."""
    def draw(rng, n):
        a = rng.uniform(-1, 1, size=(n, 2))
        b = rng.uniform(-1, 1, size=(n, 2))
        d = rng.normal(0.0, 1.0, size=(n, 2)) * np.array([0.2, 0.1])
        kd = rng.integers(20, 181, size=n)
        return dict(a=a, b=b, d=d, kd=kd)
    r = _chunked(seed, first, count, draw)
    ysp0 = np.tile([96.0, 0.5], (count, 1))
    ysp1 = ysp0 + r["a"] * np.array([1.0, 0.5])
    ysp2 = ysp1 + r["b"] * np.array([0.5, 0.25])
    seq = np.stack([ysp0, ysp1, ysp2], axis=1)                       # [B][3][ny]
    start = np.tile(np.array([0, 11, 101], dtype=np.int32), (count, 1))  # k>10, k>100 (ex.py:86-90)
    nx = 22
    x0 = np.zeros((count, nx)); x0[:, 0] = 96.0; x0[:, 1] = 0.5       # ex.py:68
    u0 = np.tile([1.95, 1.71], (count, 1))                            # ex.py:67
    return dict(x0=x0, u0=u0, ysp_seq=seq, seg_start=start, dist_step=r["kd"].astype(np.int32),
                dist=r["d"], n_steps=n_steps)


def mimo10_scenarios(count, sys, first=0, seed=SEED_MIMO + 1000, n_steps=200):
    """Config 4: set-point steps ~U(-1,1)^ny at k>10 and k>100, disturbance as config 3."""
    ny, nu, nx = sys.ny, sys.nu, sys.nx

    def draw(rng, n):
        return dict(a=rng.uniform(-1, 1, size=(n, ny)), b=rng.uniform(-1, 1, size=(n, ny)),
                    d=rng.normal(0.0, 0.1, size=(n, ny)), kd=rng.integers(20, 181, size=n))
    r = _chunked(seed, first, count, draw)
    ysp0 = np.zeros((count, ny))
    seq = np.stack([ysp0, ysp0 + r["a"], ysp0 + r["a"] + r["b"]], axis=1)
    start = np.tile(np.array([0, 11, 101], dtype=np.int32), (count, 1))
    return dict(x0=np.zeros((count, nx)), u0=np.zeros((count, nu)), ysp_seq=seq, seg_start=start,
                dist_step=r["kd"].astype(np.int32), dist=r["d"], n_steps=n_steps)


def mismatch_scenarios(count, first=0, seed=SEED_DISTILLATION + 5, n_steps=200, spread=0.2):
    """Config 5: plant K_ij(1+d), tau_ij(1+d'), d,d'~U(-spread,spread); controller = nominal model;
    set-point schedule of example_distillation.py:69,86-90.  Returns per-scenario plant CSR values
    (pattern of the nominal model: same dead times => same structure)."""
    nominal = distillation_system()
    _, _, _, pattern = plant_csr(nominal.A, nominal.B)

    def draw(rng, n):
        return dict(dk=rng.uniform(-spread, spread, size=(n, 2, 2)), dt=rng.uniform(-spread, spread, size=(n, 2, 2)))
    r = _chunked(seed, first, count, draw)
    K = np.asarray(WOOD_BERRY["K"]); tau = np.asarray(WOOD_BERRY["tau"])
    nnz = int(pattern.sum())
    vals = np.empty((count, nnz))
    # vectorised structured fill: only Bs (gains) and F/Bd (poles) entries change
    base_rowptr, base_col, base_val, _ = plant_csr(nominal.A, nominal.B, pattern)
    vals[:] = base_val
    ny, nu, nd = nominal.ny, nominal.nu, nominal.nd
    nx = nominal.nx
    rows = np.repeat(np.arange(nx), np.diff(base_rowptr))
    z0 = 2 * ny + nd
    Kp = K[None] * (1 + r["dk"]); taup = tau[None] * (1 + r["dt"])
    fp = np.exp(-nominal.Ts / taup)                                   # [B][ny][nu]
    for p in range(nnz):
        i, c = int(rows[p]), int(base_col[p])
        # which delay block does column c belong to?
        if c >= nx:
            l, j = 0, c - nx
        elif c >= z0:
            l, j = (c - z0) // nu + 1, (c - z0) % nu
        else:
            l, j = -1, -1
        if i < ny and l >= 0:                                         # Bs_l[i,j] = d0 = K
            vals[:, p] = Kp[:, i, j]
        elif ny <= i < ny + nd:
            a = i - ny
            oi, oj = a // nu, a % nu
            if l >= 0:                                                # Bd_l = dd * f = -K f
                vals[:, p] = -Kp[:, oi, oj] * fp[:, oi, oj]
            elif c == i:                                              # F diagonal
                vals[:, p] = fp[:, oi, oj]
    count_ = count
    ysp0 = np.tile([96.0, 0.5], (count_, 1))
    seq = np.stack([ysp0, np.tile([96.0, 1.0], (count_, 1)), np.tile([95.5, 1.0], (count_, 1))], axis=1)
    start = np.tile(np.array([0, 11, 101], dtype=np.int32), (count_, 1))
    x0 = np.zeros((count_, nx)); x0[:, 0] = 96.0; x0[:, 1] = 0.5
    u0 = np.tile([1.95, 1.71], (count_, 1))
    return dict(x0=x0, u0=u0, ysp_seq=seq, seg_start=start, plant_val=vals, K=Kp, tau=taup, n_steps=n_steps)


def plant_matrices_from_values(sys, vals):
    """Dense (A_p, B_p) of one scenario from its CSR values (for the oracle)."""
    rowptr, col, _, pattern = plant_csr(sys.A, sys.B)
    AB = np.zeros((sys.nx, sys.nx + sys.nu))
    rows = np.repeat(np.arange(sys.nx), np.diff(rowptr))
    AB[rows, col] = vals
    return AB[:, :sys.nx], AB[:, sys.nx:]
