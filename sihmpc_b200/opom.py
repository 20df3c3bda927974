# -*- coding: utf-8 -*-
"""OPOM (Output Prediction Oriented Model) builder -- host side, NumPy FP64.

Stands in for the un-vendored third-party package ``opom`` the reference imports
(``/root/reference/example_distillation.py:11``: ``from opom import OPOM,
TransferFunction``; README.md:1-3 installs it unpinned from
github.com/mllima/OPOM).  Its source is NOT under /root/reference, so this is a
restatement of the published model form (Odloak 2004; Rodrigues & Odloak 2003;
Santoro & Odloak 2012 for the dead-time states), shaped to expose exactly the
attributes the reference controller reads (``sihmpc/ihmpccontroller.py:17-34,
101-104,333-334``): ``Ts, theta_max, nx, nu, ny, nd, nz, A, B, C, D, Psi, F``.

Model form.  For every channel (output i, input j) the step response is
``S_ij(t) = d0 + sum_a dd_a * exp(r_a t) + di * t`` (t measured after the dead
time theta_ij).  State ``x = [xs (ny) | xd (nd) | xi (ny) | z (nu*theta_max)]``
with ``z_l(k) = du(k-l)`` (l = 1..theta_max, most recent first -- the ordering
the reference's own scratch vectors pin: ``teste.py:80-101`` has x[8:10] equal
to the previous move).  One step:

    xs+ = xs + Ts*xi + sum_l Bs_l z_l      Bs_l[i,j] = (d0_ij + Ts*di_ij)  if theta_ij == l
    xd+ = F xd       + sum_l Bd_l z_l      Bd_l[(i,j,a), j] = dd_ija * F_ija if theta_ij == l
    xi+ = xi         + sum_l Bi_l z_l      Bi_l[i,j] = di_ij              if theta_ij == l
    z_1+ = du,  z_{l+1}+ = z_l             (z_0 := du, the move being applied)
    y   = xs + Psi xd                       (C = [I, Psi, 0, 0], D = 0)

``F = diag(exp(r*Ts))`` ordered output-major: index ``(i*nu + j)*na + a``.

Under-damped channels: a complex-conjugate pole pair ``p, conj(p)`` with residues ``c, conj(c)``
contributes ``2 Re(c exp(p t))`` to the step response.  It takes TWO real states (Re, Im of the
complex mode ``xi+ = f xi + c f du``, ``f = exp(p Ts)``): ``F`` gets the block
``[[Re f, -Im f], [Im f, Re f]]``, the input column ``(Re(c f), Im(c f))`` and ``Psi`` the weights
``(2, 0)``.  ``F`` is then block diagonal instead of diagonal; the controller only needs it in the
Lyapunov equation of the terminal weight (``ihmpccontroller.py:338-339``), which is general.

Repeated real poles: a pole r of multiplicity m contributes ``sum_k c_k t^k exp(r t)/k!``; its m states are
``w_k(t) = sum_{q>=k} c_q t^(q-k) exp(r t)/(q-k)!`` (``w_0`` = the response term, ``Psi`` weight 1 on it only) and
``F`` gets the upper-triangular block ``exp(r Ts) Ts^(q-k)/(q-k)!``.

Dead times that are not a multiple of Ts: ``theta = (l + f) Ts``; a move first shows ``l + 1`` samples later, when
``tau_e = (l + 1) Ts - theta`` has elapsed since the dead time ran out.  It enters the states when it is ``l`` samples
old (``theta_ij := l`` in the formulas above) with every step-response term evaluated at ``tau_e`` instead of ``Ts``:
``Bs = d0 + tau_e di``, ``Bd = dd exp(r tau_e)`` (``tau_e = Ts`` for whole multiples).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import signal

__all__ = ["TransferFunction", "OPOM"]


class TransferFunction(signal.TransferFunction):
    """Continuous-time SISO transfer function with a pure dead time.

    Call form of the reference: ``TransferFunction(num, den, delay=1)``
    (``example_distillation.py:16-30``).
    """

    def __new__(cls, num, den, delay=0.0):
        return signal.TransferFunction.__new__(cls, num, den)

    def __init__(self, num, den, delay=0.0):
        signal.TransferFunction.__init__(self, num, den)
        self.delay = float(delay)

    def step_coefficients(self, pole_tol=1e-9):
        """Partial fractions of G(s)/s -> (d0, [(dd_a, r_a)...], di).

        Supported pole sets: stable poles, real -- distinct (``(dd, r)`` floats) or REPEATED
        (``(('jordan', [c_0 .. c_{m-1}]), r)``: the terms c_k t^k e^{rt}/k! of the step response) -- or distinct
        complex-conjugate pairs (``(c, p)`` complex, one entry per pair, Im p > 0), plus at most one pole at the
        origin (integrating channel); repeated complex pairs as ``(('cjordan', [c_k complex]), p)``.  Unstable
        poles raise NotImplementedError.
        """
        num = np.atleast_1d(np.asarray(self.num, dtype=float))
        den = np.atleast_1d(np.asarray(self.den, dtype=float))
        # G(s)/s
        r, p, k = signal.residue(num, np.polymul(den, [1.0, 0.0]))
        if len(np.atleast_1d(k)) and np.any(np.abs(k) > 0):
            raise NotImplementedError("improper transfer function (direct term in G(s)/s)")
        zero = np.abs(p) <= pole_tol
        nzero = int(np.sum(zero))
        if nzero > 2:
            raise NotImplementedError("more than one integrator per channel")
        stable = p[~zero]
        if np.any(np.real(stable) >= 0):
            raise NotImplementedError("unstable pole in channel")
        # scipy.signal.residue orders repeated roots by ascending power: r/(s-p), r/(s-p)^2, ...
        rz = np.real(r[zero])
        d0 = float(rz[0]) if nzero >= 1 else 0.0
        di = float(rz[1]) if nzero == 2 else 0.0
        modes = []
        rs, ps = list(r[~zero]), list(stable)
        rep_tol = 1e-6                              # residue() groups roots closer than its own tolerance (1e-3)
        k = 0
        while k < len(ps):
            pp = ps[k]
            mult = 1
            while k + mult < len(ps) and abs(ps[k + mult] - pp) <= rep_tol * max(1.0, abs(pp)):
                mult += 1
            if mult > 1:
                if abs(pp.imag) > pole_tol:
                    if pp.imag > 0:                 # one entry per conjugate pair; complex coefficients
                        modes.append((('cjordan', [complex(c) for c in rs[k:k + mult]]), complex(pp)))
                else:
                    modes.append((('jordan', [float(np.real(c)) for c in rs[k:k + mult]]), float(np.real(pp))))
            elif abs(pp.imag) <= pole_tol:
                modes.append((float(np.real(rs[k])), float(np.real(pp))))
            elif pp.imag > 0:                       # one entry per conjugate pair
                modes.append((complex(rs[k]), complex(pp)))
            k += mult
        return d0, modes, di


class OPOM(object):
    """OPOM state-space realisation of a matrix ``H`` of TransferFunction objects.

    ``OPOM(h, Ts)`` with ``h = [[h11, h12], [h21, h22]]``
    (``example_distillation.py:33-38``).
    """

    def __init__(self, H, Ts):
        Hm = np.array(H, dtype=object)
        if Hm.ndim == 0:
            Hm = Hm.reshape(1, 1)
        if Hm.ndim == 1:
            Hm = Hm.reshape(1, -1)
        self.H = Hm
        self.ny, self.nu = Hm.shape
        self.Ts = Ts
        ny, nu = self.ny, self.nu

        coeffs = [[Hm[i, j].step_coefficients() for j in range(nu)] for i in range(ny)]

        def n_states(modes):
            return sum((2 if isinstance(mm[1], complex) else 1) * (len(mm[0][1]) if isinstance(mm[0], tuple) else 1)
                       for mm in modes)
        self.na = max(1, max(n_states(coeffs[i][j][1]) for i in range(ny) for j in range(nu)))
        na = self.na
        self.nd = ny * nu * na

        # Dead time theta = (l + f) Ts, 0 <= f < 1: a move made at sample k first shows at sample k + l + 1, by
        # then tau_e = (l + 1) Ts - theta in (0, Ts] has elapsed since the dead time ran out (tau_e = Ts for a
        # dead time that is a multiple of Ts).  The move enters the states when it is l samples old, with the
        # step-response terms evaluated at tau_e.
        theta = np.zeros((ny, nu), dtype=int)
        self.tau_e = np.full((ny, nu), float(Ts))
        for i in range(ny):
            for j in range(nu):
                th = Hm[i, j].delay / Ts
                if th < -1e-12:
                    raise ValueError("negative dead time")
                l = int(np.floor(th + 1e-9))
                theta[i, j] = l
                self.tau_e[i, j] = (l + 1) * Ts - Hm[i, j].delay if abs(th - l) > 1e-9 else float(Ts)
        self.theta = theta
        self.theta_max = int(theta.max())
        self.nz = self.theta_max * nu
        self.nx = 2 * ny + self.nd + self.nz

        # coefficient tables
        self.D0 = np.zeros((ny, nu))
        self.Di = np.zeros((ny, nu))
        self.Dd = np.zeros(self.nd)          # diagonal of the Dd matrix (real modes)
        self.R = np.zeros(self.nd)           # poles (real modes; real part of complex ones)
        self.F = np.zeros((self.nd, self.nd))    # padded (non-existent) modes get F = 0
        self._bd = np.zeros(self.nd)         # what a move adds to each xd state when its dead time is over
        self.Psi = np.zeros((ny, self.nd))
        for i in range(ny):
            for j in range(nu):
                d0, modes, di = coeffs[i][j]
                self.D0[i, j] = d0
                self.Di[i, j] = di
                a = 0
                te = self.tau_e[i, j]
                for dd, r in modes:
                    idx = (i * nu + j) * na + a
                    if isinstance(r, complex) and isinstance(dd, tuple):
                        # repeated complex pair: the complex Jordan states of the real case, each as (Re, Im)
                        c = dd[1]
                        mlt = len(c)
                        for k in range(mlt):
                            bk = 0.0 + 0.0j
                            for q in range(k, mlt):
                                fkq = np.exp(r * Ts) * Ts ** (q - k) / math.factorial(q - k)
                                a0, b0 = idx + 2 * k, idx + 2 * q
                                self.F[a0:a0 + 2, b0:b0 + 2] = [[fkq.real, -fkq.imag], [fkq.imag, fkq.real]]
                                bk += c[q] * te ** (q - k) * np.exp(r * te) / math.factorial(q - k)
                            self._bd[idx + 2 * k], self._bd[idx + 2 * k + 1] = bk.real, bk.imag
                            self.R[idx + 2 * k] = self.R[idx + 2 * k + 1] = r.real
                        self.Psi[i, idx] = 2.0
                        a += 2 * mlt
                    elif isinstance(r, complex):
                        f = np.exp(r * Ts)
                        self.F[idx:idx + 2, idx:idx + 2] = [[f.real, -f.imag], [f.imag, f.real]]
                        g = dd * np.exp(r * te)
                        self._bd[idx], self._bd[idx + 1] = g.real, g.imag
                        self.Psi[i, idx] = 2.0
                        self.R[idx] = self.R[idx + 1] = r.real
                        a += 2
                    elif isinstance(dd, tuple):
                        # repeated real pole: states w_k(t) = sum_{q>=k} c_q t^(q-k) e^{rt}/(q-k)!  (w_0 = the
                        # response term); over one sample w <- e^{r Ts} sum_n (Ts^n/n!) shift^n w
                        c = dd[1]
                        mlt = len(c)
                        for k in range(mlt):
                            for q in range(k, mlt):
                                self.F[idx + k, idx + q] = np.exp(r * Ts) * Ts ** (q - k) / math.factorial(q - k)
                                self._bd[idx + k] += c[q] * te ** (q - k) * np.exp(r * te) / math.factorial(q - k)
                            self.R[idx + k] = r
                        self.Dd[idx] = c[0]
                        self.Psi[i, idx] = 1.0
                        a += mlt
                    else:
                        f = np.exp(r * Ts)
                        self.Dd[idx] = dd
                        self.R[idx] = r
                        self.F[idx, idx] = f
                        self._bd[idx] = dd * np.exp(r * te)
                        self.Psi[i, idx] = 1.0
                        a += 1
                # states of this channel that no mode uses keep F = 0 (they stay at zero)
                for a2 in range(a, na):
                    self.Psi[i, (i * nu + j) * na + a2] = 1.0
        self.N = np.zeros((self.nd, nu))
        for i in range(ny):
            for j in range(nu):
                for a in range(na):
                    self.N[(i * nu + j) * na + a, j] = 1.0

        self.A, self.B, self.C, self.D = self._state_space()
        self.X = np.zeros(self.nx)

    # ---- structured blocks (also used by the batched plant update on the device)
    def input_blocks(self, l):
        """(Bs_l, Bd_l, Bi_l): how a move that is ``l`` samples old enters xs, xd, xi."""
        ny, nu, na, Ts = self.ny, self.nu, self.na, self.Ts
        mask = (self.theta == l)
        Bs = np.where(mask, self.D0 + self.tau_e * self.Di, 0.0)
        Bi = np.where(mask, self.Di, 0.0)
        Bd = np.zeros((self.nd, nu))
        for i in range(ny):
            for j in range(nu):
                if mask[i, j]:
                    for a in range(na):
                        idx = (i * nu + j) * na + a
                        Bd[idx, j] = self._bd[idx]
        return Bs, Bd, Bi

    def _state_space(self):
        ny, nu, nd, nz, nx, Ts = self.ny, self.nu, self.nd, self.nz, self.nx, self.Ts
        s0, d0, i0, z0 = 0, ny, ny + nd, 2 * ny + nd
        A = np.zeros((nx, nx))
        B = np.zeros((nx, nu))
        A[s0:s0 + ny, s0:s0 + ny] = np.eye(ny)
        A[s0:s0 + ny, i0:i0 + ny] = Ts * np.eye(ny)
        A[d0:d0 + nd, d0:d0 + nd] = self.F
        A[i0:i0 + ny, i0:i0 + ny] = np.eye(ny)
        for l in range(0, self.theta_max + 1):
            Bs, Bd, Bi = self.input_blocks(l)
            if l == 0:
                B[s0:s0 + ny, :] = Bs
                B[d0:d0 + nd, :] = Bd
                B[i0:i0 + ny, :] = Bi
            else:
                c = z0 + (l - 1) * nu
                A[s0:s0 + ny, c:c + nu] = Bs
                A[d0:d0 + nd, c:c + nu] = Bd
                A[i0:i0 + ny, c:c + nu] = Bi
        if nz:
            B[z0:z0 + nu, :] = np.eye(nu)
            for l in range(1, self.theta_max):
                A[z0 + l * nu:z0 + (l + 1) * nu, z0 + (l - 1) * nu:z0 + l * nu] = np.eye(nu)
        C = np.zeros((ny, nx))
        C[:, s0:s0 + ny] = np.eye(ny)
        C[:, d0:d0 + nd] = self.Psi
        D = np.zeros((ny, nu))
        return A, B, C, D

    # ---- convenience simulation (host only)
    def output(self, du):
        """Advance the internal state by one move ``du`` and return y."""
        du = np.asarray(du, dtype=float).reshape(self.nu)
        self.X = self.A @ self.X + self.B @ du
        return self.C @ self.X + self.D @ du

    def step_response(self, n, j=0):
        """Unit step in input j held for n samples -> (n+1, ny) outputs, y[0] at t=0."""
        x = np.zeros(self.nx)
        out = [self.C @ x]
        for k in range(n):
            du = np.zeros(self.nu)
            if k == 0:
                du[j] = 1.0
            x = self.A @ x + self.B @ du
            out.append(self.C @ x + self.D @ du)
        return np.array(out)
