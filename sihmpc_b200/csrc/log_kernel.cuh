// log_kernel.cuh -- log-barrier satisficing IHMPC step (ihmpccontroller_log.py), one warp per scenario.
//
//   minimise  phi(dU, ss) = Vt(dU) + ss^2 - sum_{i in VJ} log( ss*gamma_i - V_i(dU) )
//   (ihmpccontroller_log.py:71-92,196,219,241,262,292; w = [dU ; syN ; siN ; ss], :375-392)
// with the terminal equalities eliminated exactly as in the weighted variant.  With
// d_i = ss*gamma_i - V_i the Newton system is the weighted assembly with pesos_i := 1/d_i plus a
// rank-m term:
//   grad_dU = 2 E^T((w0row + Wrow (1/d)) o r) + 2 (Wdu (1/d)) o dU          grad_ss = 2 ss - sum gamma_i/d_i
//   H_dUdU  = Ht + sum (1/d_i) HJ_i + 2 diag(Wdu (1/d)) + sum a_i a_i^T / d_i^2 ,  a_i = grad V_i
//   H_dU,ss = - sum gamma_i a_i / d_i^2                                       H_ss,ss = 2 + sum gamma_i^2 / d_i^2
// Damped Newton (in-shared-memory Cholesky of the (nv+1)x(nv+1) system, backtracking that keeps every
// d_i > 0).  The explicit inequalities of the reference NLP (ss >= 1, ss*gamma_i - V_i >= 0.01,
// input and move boxes; log.py:390-392,420-424,435-440) are checked at that minimiser (they are inactive in the
// reference's own example: ss* = sqrt(m/2) > 1).  If one is not strictly inactive the step is solved again as an
// INTERIOR-POINT problem, the way the reference's IPOPT treats them: every explicit inequality c_j(v) >= 0 gets
// its own barrier  -mu log c_j  (bounds relaxed by 1e-10 max(1,|b|) -- IPOPT relaxes by 1e-8, bound_relax_factor -- so a point on
// a bound is strictly interior), Newton on phi_mu with the same machinery, mu from 1e-1 down to 1e-14 (a weakly
// active row with multiplier 1e-4 is then within 1e-10 of its bound).  The extra
// terms: ss >= 1 and the shifted d_i - 0.01 change the weights 1/d_i, 1/d_i^2 of the existing blocks; a row of the
// input box (cumulative moves of one input up to step k) adds t1 = -mu/s_lo + mu/s_hi to the gradient of every
// move of that input up to k and t2 = mu/s_lo^2 + mu/s_hi^2 to every pair of them (suffix sums S1, S2 over k);
// the move box adds diagonal terms.  SIHMPC_ST_ACTIVE_UNSUPPORTED is only left set when that solve fails
// (no strictly feasible start: u outside its box, 0 outside the move box).
#pragma once
#include "step_kernel.cuh"

namespace sih {


// Cholesky of the lower triangle of M (n <= NMAX <= 32) with lane i holding ROW i in registers: the Crout dot products
// s_i = M_ik - sum_{j<k} L_ij L_kj run over compile-time j (both loops fully unrolled), L_kj comes as a broadcast LDS
// from the finished row k in shared memory, every finished column goes back to shared memory for the substitutions.
// Same operation order as warp_cholesky (even / odd j on two chains): bitwise the same factor, a third of its instructions
// (the in-shared-memory loop spent them on addresses, loop control and the second LDS per product).
template <int NMAX>
__device__ int warp_cholesky_rows(double *M, const int ld, const int n, double *dinv, const int lane) {
    double row[NMAX];
    {
        const double *Mi = M + (size_t)lane * ld;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) row[j] = (lane < n && j <= lane) ? Mi[j] : 0.0;
    }
    int bad = 0;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
        if (k < n) {                                        // (uniform)
            const double *Mk = M + (size_t)k * ld;
            double a0 = row[k], a1 = 0.0;
#pragma unroll
            for (int j = 0; j < k; ++j) {
                if (j & 1) a1 = fma(-row[j], Mk[j], a1);
                else a0 = fma(-row[j], Mk[j], a0);
            }
            const double sv = a0 + a1;
            double d = __shfl_sync(SIH_FULL, sv, k);
            if (!(d > 0.0)) { bad = 1; d = 1e-300; }
            const double inv = rsqrt(d);
            const double l = sv * inv;                      // lane k: d*inv = sqrt(d)
            row[k] = l;
            if (lane >= k && lane < n) M[(size_t)lane * ld + k] = l;
            if (lane == 0) dinv[k] = inv;
            __syncwarp();
        }
    }
    return bad;
}

struct LogLayout { int H, dinv, x, xn, uu, ysp, e, r, v, g, p, gi, d, Vi, vt, s1, s2, b1, b2, glo, ghi, t1, t2, vw, w12, total; };

__host__ __device__ inline LogLayout make_log_layout(const DevPlan &P) {
    LogLayout w; int o = 0;
    const int n = P.nv + 1, ld = n | 1;
#define SIH_TAKE(field, cnt) do { w.field = o; o += ((cnt) + 1) & ~1; } while (0)
    SIH_TAKE(H, n * ld); SIH_TAKE(dinv, n); SIH_TAKE(x, P.nx + P.nu); SIH_TAKE(xn, P.nx); SIH_TAKE(uu, P.nu); SIH_TAKE(ysp, P.ny); SIH_TAKE(e, P.R);
    SIH_TAKE(r, P.R); SIH_TAKE(v, n); SIH_TAKE(g, n); SIH_TAKE(p, n); SIH_TAKE(gi, P.m * P.nv);
    SIH_TAKE(d, P.m); SIH_TAKE(Vi, P.m); SIH_TAKE(vt, n);
    SIH_TAKE(s1, n); SIH_TAKE(s2, n); SIH_TAKE(b1, n); SIH_TAKE(b2, n);
    const int nc = P.ncd > 0 ? P.ncd : 1;           // dense rows: bounds on Gc dU, barrier coefficients
    SIH_TAKE(glo, nc); SIH_TAKE(ghi, nc); SIH_TAKE(t1, nc); SIH_TAKE(t2, nc);
    SIH_TAKE(vw, n);                                // the point this step started from (warm start / shifted plan)
    SIH_TAKE(w12, 2 * P.m);                         // first / second derivative weights of the barrier terms at the current point
#undef SIH_TAKE
    w.total = o;
    return w;
}

// residual r = ET v + e, sub-objective values, d_i; returns phi (inf when some d_i <= 0).
// MM > 0 (= m, compile time): every lane keeps the m partial sums of ITS residual rows / moves in registers (no second
// pass over r by m lanes out of 32), one warp sum per sub-objective, lane i takes log d_i (the m logarithms were
// evaluated one after the other by every lane).  The result is bitwise the same on every lane (xor butterfly).
template <int MM>
__device__ double log_eval(const DevPlan &P, const double *vv, const double *e, double *r, double *Vi,
                           double *d, int lane) {
    const int R = P.R, nv = P.nv, m = P.m;
    if constexpr (MM > 0) {
        double vt = 0.0, acc[MM];
#pragma unroll
        for (int i = 0; i < MM; ++i) acc[i] = 0.0;
        for (int k = lane; k < R; k += 32) {
            double a = e[k];
            for (int c = 0; c < nv; ++c) a = fma(P.ET[(size_t)c * R + k], vv[c], a);
            r[k] = a;
            const double a2 = a * a;
            vt = fma(P.w0row[k], a2, vt);
            const double *wr = P.Wrow + (size_t)k * MM;
#pragma unroll
            for (int i = 0; i < MM; ++i) acc[i] = fma(wr[i], a2, acc[i]);
        }
        for (int c = lane; c < nv; c += 32) {
            const double v2 = vv[c] * vv[c];
            const double *wd = P.Wdu + (size_t)c * MM;
#pragma unroll
            for (int i = 0; i < MM; ++i) acc[i] = fma(wd[i], v2, acc[i]);
        }
        vt = warp_sum(vt);
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < MM; ++i) { const double t = warp_sum(acc[i]); if (lane == i) mine = t; }
        const double ss = vv[nv];
        const bool own = lane < MM;
        const double di = ss * P.gamma[own ? lane : 0] - mine;
        if (own) { Vi[lane] = mine; d[lane] = di; }
        const bool bad = __any_sync(SIH_FULL, own && !(di > 0.0));
        const double lg = warp_sum((own && di > 0.0) ? log(di) : 0.0);
        __syncwarp();
        return bad ? SIH_INF : (vt + ss * ss) - lg;
    }
    double vt = 0.0;
    for (int k = lane; k < R; k += 32) {
        double a = e[k];
        for (int c = 0; c < nv; ++c) a = fma(P.ET[(size_t)c * R + k], vv[c], a);
        r[k] = a;
        vt = fma(P.w0row[k], a * a, vt);
    }
    vt = warp_sum(vt);
    __syncwarp();
    const double ss = vv[nv];
    for (int i = lane; i < m; i += 32) {
        double a = 0.0;
        for (int k = 0; k < R; ++k) a = fma(P.Wrow[(size_t)k * m + i], r[k] * r[k], a);
        for (int c = 0; c < nv; ++c) a = fma(P.Wdu[(size_t)c * m + i], vv[c] * vv[c], a);
        Vi[i] = a;
        d[i] = ss * P.gamma[i] - a;
    }
    __syncwarp();
    double phi = vt + ss * ss, bad = 0.0;
    for (int i = 0; i < m; ++i) { const double di = d[i]; if (!(di > 0.0)) bad = 1.0; else phi -= log(di); }
    return bad > 0.0 ? SIH_INF : phi;
}

__device__ __forceinline__ double log_relax(const double b, const double sgn) {       // cf. IPOPT bound_relax_factor (1e-8 there)
    return b + sgn * 1e-10 * fmax(1.0, fabs(b));
}

// barrier of the explicit inequalities at vv (d[] from log_eval at the same point):  -mu sum log c_j, inf outside.
// With `deriv` the first/second derivative coefficients are left in shared memory: S1/S2[k*nu+j] = suffix sums
// over the horizon of the input-box rows of input j, b1/b2 = move-box terms per variable.
// Dense rows lo <= Gc dU <= hi (state / slack boxes, the siN contraction row |siN| <= sqrt(ViN_ant/q) of an
// integrating plant, log.py:245-270,429-434): glo/ghi hold the bounds on Gc dU of THIS step, T1/T2 receive
// mu (-1/s_lo + 1/s_hi) and mu (1/s_lo^2 + 1/s_hi^2) per row.
__device__ double log_barrier(const DevPlan &P, const double *vv, const double *uu, const double *d, const double mu,
                              const bool deriv, double *S1, double *S2, double *b1, double *b2, int lane,
                              const double *glo = nullptr, const double *ghi = nullptr, double *T1 = nullptr,
                              double *T2 = nullptr) {
    const int nv = P.nv, nu = P.nu, m = P.m, N = P.N;
    double acc = 0.0, bad = 0.0;
    for (int r = lane; r < P.ncd; r += 32) {
        double val = 0.0;
        for (int c = 0; c < nv; ++c) val = fma(P.GcT[(size_t)c * P.ncd + r], vv[c], val);
        double t1 = 0.0, t2 = 0.0;
        if (glo[r] > -SIH_INF) { const double sl = val - log_relax(glo[r], -1.0); if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); t1 -= 1.0 / sl; t2 += 1.0 / (sl * sl); } }
        if (ghi[r] < SIH_INF) { const double sl = log_relax(ghi[r], 1.0) - val; if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); t1 += 1.0 / sl; t2 += 1.0 / (sl * sl); } }
        if (deriv) { T1[r] = mu * t1; T2[r] = mu * t2; }
    }
    if (lane == 0) {
        const double s = vv[nv] - log_relax(1.0, -1.0);
        if (!(s > 0.0)) bad = 1.0; else acc -= log(s);
        for (int i = 0; i < m; ++i) {
            const double si = d[i] - log_relax(0.01, -1.0);
            if (!(si > 0.0)) bad = 1.0; else acc -= log(si);
        }
    }
    for (int j = lane; j < nu; j += 32) {
        const bool hl = P.ulb[j] > -SIH_INF, hh = P.uub[j] < SIH_INF, dl = P.dulb[j] > -SIH_INF, dh = P.duub[j] < SIH_INF;
        const double lo = hl ? log_relax(P.ulb[j], -1.0) : 0.0, hi = hh ? log_relax(P.uub[j], 1.0) : 0.0;
        const double dlo = dl ? log_relax(P.dulb[j], -1.0) : 0.0, dhi = dh ? log_relax(P.duub[j], 1.0) : 0.0;
        double cs = uu[j];
        for (int k = 0; k < N; ++k) {
            const int c = k * nu + j;
            const double dv = vv[c];
            cs += dv;
            double t1 = 0.0, t2 = 0.0, u1 = 0.0, u2 = 0.0;
            if (hl) { const double sl = cs - lo; if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); t1 -= 1.0 / sl; t2 += 1.0 / (sl * sl); } }
            if (hh) { const double sl = hi - cs; if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); t1 += 1.0 / sl; t2 += 1.0 / (sl * sl); } }
            if (dl) { const double sl = dv - dlo; if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); u1 -= 1.0 / sl; u2 += 1.0 / (sl * sl); } }
            if (dh) { const double sl = dhi - dv; if (!(sl > 0.0)) bad = 1.0; else { acc -= log(sl); u1 += 1.0 / sl; u2 += 1.0 / (sl * sl); } }
            if (deriv) { S1[c] = mu * t1; S2[c] = mu * t2; b1[c] = mu * u1; b2[c] = mu * u2; }
        }
        if (deriv)
            for (int k = N - 2; k >= 0; --k) { S1[k * nu + j] += S1[(k + 1) * nu + j]; S2[k * nu + j] += S2[(k + 1) * nu + j]; }
    }
    acc = warp_sum(acc);
    bad = warp_max(bad);
    __syncwarp();
    return bad > 0.0 ? SIH_INF : mu * acc;
}

// One launch = n_steps control steps of every scenario.  n_steps == 1 without SIHMPC_F_ADVANCE is mpc() of
// the log controller (ihmpccontroller_log.py:552-563); with SIHMPC_F_ADVANCE the body of the driver loop
// (example_distillation_log.py:85-135: mpc, first move to the plant, warmStart with ss = 1, log.py:518-549)
// runs in place with the state resident in shared memory.
#ifndef SIH_LOG_MINB
#define SIH_LOG_MINB 4
#endif
template <int NVR, int MM>
__global__ void __launch_bounds__(128, SIH_LOG_MINB)
ihmpc_log_kernel(const DevPlan P, const sihmpc_io io, const double *__restrict__ w0, double *__restrict__ w_warm,
                 const long long batch, const int n_steps, int *iters_out, const int max_iter) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * wpb + wib, nwarps = (long long)gridDim.x * wpb;
    const LogLayout lay = make_log_layout(P);
    double *W = smem + (size_t)wib * lay.total;
    double *Hm = W + lay.H, *dinv = W + lay.dinv, *xs = W + lay.x, *ys = W + lay.ysp, *e = W + lay.e;
    double *r = W + lay.r, *v = W + lay.v, *g = W + lay.g, *gi = W + lay.gi, *d = W + lay.d, *Vi = W + lay.Vi;
    double *vt = W + lay.vt, *xn = W + lay.xn, *uu = W + lay.uu;
    double *S1 = W + lay.s1, *S2 = W + lay.s2, *b1 = W + lay.b1, *b2 = W + lay.b2;
    double *glo = W + lay.glo, *ghi = W + lay.ghi, *T1 = W + lay.t1, *T2 = W + lay.t2, *vw = W + lay.vw;
    double *w1s = W + lay.w12, *w2s = w1s + P.m;
    double *vfree = W + lay.p;                              // minimiser of phi without the explicit inequalities
    const int ncd = P.ncd;
    const int nx = P.nx, nu = P.nu, ny = P.ny, nv = P.nv, m = P.m, R = P.R;
    const int nvm = P.nvm;                                  // moves; v[nvm..nv) = ranged terminal residuals rho (log.py:413-427)
    const int n = nv + 1, ld = n | 1, nw = nvm + 2 * ny + 1;
    const int nstat = ny + SIHMPC_NSTATS_EXTRA;
    const bool advance = (io.flags & SIHMPC_F_ADVANCE) != 0;
    double *const w_opt = io.w_opt;
    double *const J = io.J;
    int *const status = io.status;

    for (long long s = gw; s < batch; s += nwarps) {
        for (int i = lane; i < nx; i += 32) xs[i] = io.x[s * nx + i];
        for (int i = lane; i < nu; i += 32) uu[i] = io.u[s * nu + i];
        for (int i = lane; i < nv; i += 32) v[i] = (w0 && i < nvm) ? w0[s * nw + i] : 0.0;
        if (lane == 0) v[nv] = w0 ? w0[s * nw + nw - 1] : 1.0;
        // ss of the last optimum of this scenario (0: none).  The warm start of the reference resets ss to 1 (log.py:548); the
        // Newton iteration of the NEXT step may start wherever it likes -- the minimiser of the strictly convex phi is the
        // same -- and started at the last ss* (sqrt(m/2) near steady state) it needs one iteration less (5 -> 4; 1 at rest).
        // Carried between launches in io.pesos [batch] (unused otherwise by the log variant; NULL: within the launch only).
        double ss_hint = ((io.flags & SIHMPC_F_ADVANCE) && io.pesos) ? io.pesos[s] : 0.0;
        double st_ise = 0.0, st_du2 = 0.0, st_J = 0.0;      // lane i < ny holds the ISE of output i
        int st_or = 0, it_total = 0;
        if (io.stats) {
            if (lane < ny) st_ise = io.stats[s * nstat + lane];
            st_du2 = io.stats[s * nstat + ny]; st_or = (int)io.stats[s * nstat + ny + 3];
            it_total = (int)io.stats[s * nstat + ny + 4];
        }
      for (int step = 0; step < n_steps; ++step) {
        const int kabs = io.k0 + step;
        if (io.ysp_seq) {
            int seg = 0;
            for (int a = 1; a < io.n_seg; ++a)
                if (io.seg_start[s * io.n_seg + a] <= kabs) seg = a;
            for (int i = lane; i < ny; i += 32) ys[i] = io.ysp_seq[(s * io.n_seg + seg) * ny + i];
        } else {
            for (int i = lane; i < ny; i += 32) ys[i] = io.ysp[s * ny + i];
        }
        __syncwarp();
        for (int k = lane; k < R; k += 32) {
            double a = P.e0[k];
            for (int j = 0; j < nx; ++j) a = fma(P.PxT[(size_t)j * R + k], xs[j], a);
            for (int c = 0; c < ny; ++c) a = fma(P.PyT[(size_t)c * R + k], ys[c], a);
            e[k] = a;
        }
        // dense rows: bounds on Gc dU for this state (contraction rows from ViN_ant: +inf = not yet set)
        for (int rr = lane; rr < ncd; rr += 32) {
            double off = P.gc0[rr];
            for (int j = 0; j < nx; ++j) off = fma(P.GcxT[(size_t)j * ncd + rr], xs[j], off);
            for (int c = 0; c < ny; ++c) off = fma(P.GcyT[(size_t)c * ncd + rr], ys[c], off);
            double lo = P.Lc[rr], hi = P.Uc[rr];
            const int slot = P.gc_vin[rr];
            if (slot >= 0) {
                const double vin = io.vinant ? io.vinant[s * P.n_vin + slot] : SIH_INF;
                const double b = (vin < SIH_INF) ? sqrt(fmax(vin, 0.0) / P.gc_q[rr]) : SIH_INF;
                lo = -b; hi = b;
            }
            glo[rr] = lo - off; ghi[rr] = hi - off;
        }
        for (int i = lane; i < nv; i += 32) vw[i] = v[i];
        __syncwarp();
        int st = 0, it = 0;
        double phi = log_eval<MM>(P, v, e, r, Vi, d, lane);
        {   // make the start strictly feasible for the barrier: ss >= 1 and ss*gamma_i > V_i
            double smax = 0.0;
            for (int i = 0; i < m; ++i) smax = fmax(smax, Vi[i] / P.gamma[i]);
            const double ss0 = fmax(fmax(fmax(v[nv], 1.0), 1.1 * smax + 0.1), (ss_hint < 1e6) ? ss_hint : 0.0);
            __syncwarp();
            if (lane == 0) v[nv] = ss0;
            __syncwarp();
            phi = log_eval<MM>(P, v, e, r, Vi, d, lane);
        }
        // Newton on phi (mu = 0) or phi_mu = phi - mu sum log c_j (interior-point rounds), at most `lim` iterations
        auto newton = [&](const double mu, const int lim) {
            // weights of the barrier terms, ONCE per iteration (they were two FP64 divisions inside every term of every
            // Hessian element: 17 % of the kernel's samples)
            const int m_ = MM > 0 ? MM : m;                  // compile-time trip count of the sub-objective loops when MM > 0
            constexpr int MR = MM > 0 ? MM : 1;
            bool near = false;                               // the last step was a full Newton step with a small decrement
        for (int itn = 1; itn <= lim; ++itn) {
            ++it;
            if (mu > 0.0) log_barrier(P, v, uu, d, mu, true, S1, S2, b1, b2, lane, glo, ghi, T1, T2);     // derivative coefficients at v
            for (int i = lane; i < m; i += 32) {
                const double di = d[i], si = di - log_relax(0.01, -1.0);
                w1s[i] = 1.0 / di + (mu > 0.0 ? mu / si : 0.0);
                w2s[i] = 1.0 / (di * di) + (mu > 0.0 ? mu / (si * si) : 0.0);
            }
            __syncwarp();
            // (MM > 0: the weights and the weights of the mh Hessian blocks live in registers for the whole iteration)
            double w1r[MR], w2r[MR], wq[MR];
            if constexpr (MM > 0) {
#pragma unroll
                for (int i = 0; i < MM; ++i) { w1r[i] = w1s[i]; w2r[i] = w2s[i]; wq[i] = (i < P.mh) ? w1s[P.hj_index[i]] : 0.0; }
            }
            auto w1f = [&](const int i) { if constexpr (MM > 0) return w1r[i]; else return w1s[i]; };
            auto w2f = [&](const int i) { if constexpr (MM > 0) return w2r[i]; else return w2s[i]; };
            // ---- gradients a_i = grad_dU V_i (rows of gi), lanes over variables; E and r read once per four sub-objectives
            for (int c = lane; c < nv; c += 32) {
                double a = 0.0;
                const double vc = v[c];
                if constexpr (MM > 0) {
                    double ai[MM];
#pragma unroll
                    for (int i = 0; i < MM; ++i) ai[i] = 0.0;
                    for (int k = 0; k < R; ++k) {
                        const double er = P.E[(size_t)k * nv + c] * r[k];
                        const double *wr = P.Wrow + (size_t)k * MM;
#pragma unroll
                        for (int i = 0; i < MM; ++i) ai[i] = fma(er, wr[i], ai[i]);
                        a = fma(er, P.w0row[k], a);
                    }
                    double gc = 2.0 * a;
#pragma unroll
                    for (int i = 0; i < MM; ++i) {
                        const double gic = 2.0 * fma(P.Wdu[(size_t)c * MM + i], vc, ai[i]);
                        gi[i * nv + c] = gic;
                        gc = fma(gic, w1r[i], gc);
                    }
                    a = gc;
                } else {
                    for (int i0 = 0; i0 < m; i0 += 4) {
                        const int mi = (m - i0 < 4) ? m - i0 : 4;
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        for (int k = 0; k < R; ++k) {
                            const double er = P.E[(size_t)k * nv + c] * r[k];
                            const double *wr = P.Wrow + (size_t)k * m + i0;
                            a0 = fma(er, wr[0], a0);
                            if (mi > 1) a1 = fma(er, wr[1], a1);
                            if (mi > 2) a2 = fma(er, wr[2], a2);
                            if (mi > 3) a3 = fma(er, wr[3], a3);
                            if (i0 == 0) a = fma(er, P.w0row[k], a);
                        }
                        gi[i0 * nv + c] = 2.0 * (a0 + P.Wdu[(size_t)c * m + i0] * vc);
                        if (mi > 1) gi[(i0 + 1) * nv + c] = 2.0 * (a1 + P.Wdu[(size_t)c * m + i0 + 1] * vc);
                        if (mi > 2) gi[(i0 + 2) * nv + c] = 2.0 * (a2 + P.Wdu[(size_t)c * m + i0 + 2] * vc);
                        if (mi > 3) gi[(i0 + 3) * nv + c] = 2.0 * (a3 + P.Wdu[(size_t)c * m + i0 + 3] * vc);
                    }
                    double gc = 2.0 * a;
                    for (int i = 0; i < m; ++i) gc += gi[i * nv + c] * w1f(i);
                    a = gc;
                }
                if (mu > 0.0) {
                    if (c < nvm) a += S1[c] + b1[c];
                    for (int rr = 0; rr < ncd; ++rr) a = fma(T1[rr], P.GcT[(size_t)c * ncd + rr], a);
                }
                g[c] = a;
            }
            if (lane == 0) {
                double gs = 2.0 * v[nv];
#pragma unroll
                for (int i = 0; i < m_; ++i) gs -= P.gamma[i] * w1f(i);
                if (mu > 0.0) gs -= mu / (v[nv] - log_relax(1.0, -1.0));
                g[nv] = gs;
            }
            __syncwarp();
            const double dec_tol = (mu > 0.0) ? fmin(1e-16 * (1.0 + fabs(phi)), 1e-2 * mu) : 1e-18 * (1.0 + fabs(phi));
            if (near) {
                // the factor of the LAST iteration is still in Hm: after a full step with a small decrement the Hessian has
                // moved by a relative ~sqrt(dec), so the decrement through the old factor decides "converged" as well as the
                // new one would -- the iteration that only confirms convergence skips its Hessian and factorisation
                double accc[NVR];
                load_acc<NVR>(accc, g, n, lane, -1.0);
                warp_trsv_lower<NVR>(Hm, ld, dinv, n, accc, lane);
                warp_trsv_lower_T<NVR>(Hm, ld, dinv, n, accc, lane);
                double decc = 0.0;
#pragma unroll
                for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < n) decc = fma(-g[i], accc[t], decc); }
                decc = warp_sum(decc);
                if (!(decc > 0.5 * dec_tol)) break;
                __syncwarp();
            }
            // ---- Hessian (lower triangle)
            for (int idx = lane; idx < (n * (n + 1)) / 2; idx += 32) {        // lanes over the lower triangle itself: a(a+1)/2 + b
                int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
                if ((a * (a + 1)) / 2 > idx) --a;
                if (((a + 1) * (a + 2)) / 2 <= idx) ++a;
                const int b = idx - (a * (a + 1)) / 2;
                double h;
                if (a < nv) {
                    const size_t ab = (size_t)a * nv + b;
                    h = P.Ht[ab];
                    if constexpr (MM > 0) {
#pragma unroll
                        for (int q = 0; q < MM; ++q) if (q < P.mh) h = fma(wq[q], P.HJ[(size_t)q * nv * nv + ab], h);
                    } else {
                        for (int q = 0; q < P.mh; ++q) h = fma(w1f(P.hj_index[q]), P.HJ[(size_t)q * nv * nv + ab], h);
                    }
#pragma unroll
                    for (int i = 0; i < m_; ++i) h = fma(gi[i * nv + a] * w2f(i), gi[i * nv + b], h);
                    if (a == b) {
#pragma unroll
                        for (int i = 0; i < m_; ++i) h = fma(2.0 * P.Wdu[(size_t)a * m_ + i], w1f(i), h);
                    }
                    if (mu > 0.0) {
                        if (a < nvm && (a % nu) == (b % nu)) h += S2[a];          // a >= b: rows k >= k_a cover both moves
                        if (a == b && a < nvm) h += b2[a];
                        for (int rr = 0; rr < ncd; ++rr)
                            h = fma(T2[rr] * P.GcT[(size_t)a * ncd + rr], P.GcT[(size_t)b * ncd + rr], h);
                    }
                } else if (b < nv) {
                    h = 0.0;
#pragma unroll
                    for (int i = 0; i < m_; ++i) h -= P.gamma[i] * gi[i * nv + b] * w2f(i);
                } else {
                    h = 2.0;
#pragma unroll
                    for (int i = 0; i < m_; ++i) h += P.gamma[i] * P.gamma[i] * w2f(i);
                    if (mu > 0.0) { const double sl = v[nv] - log_relax(1.0, -1.0); h += mu / (sl * sl); }
                }
                Hm[a * ld + b] = h;
            }
            __syncwarp();
#ifdef SIH_LOG_DEBUG
            double dbg_diag_min = SIH_INF, dbg_diag_max = 0.0;
            for (int i = 0; i < n; ++i) { dbg_diag_min = fmin(dbg_diag_min, Hm[i * ld + i]); dbg_diag_max = fmax(dbg_diag_max, Hm[i * ld + i]); }
            const bool dbg_np = warp_cholesky<NVR>(Hm, ld, n, dinv, lane);
            if (dbg_np) st |= SIHMPC_ST_NOTPD;
            if (s == SIH_LOG_DEBUG && lane == 0)
                printf("k=%d mu=%.1e it=%d phi=%.10g ss=%.6g notpd=%d diag[%.3g,%.3g] d=[%.3g %.3g %.3g %.3g] rho=[%.6g %.6g]\n", kabs, mu, itn, phi,
                       v[nv], (int)dbg_np, dbg_diag_min, dbg_diag_max, d[0], d[1], d[2], d[3], nv > nvm ? v[nvm] : 0.0, nv > nvm + 1 ? v[nvm + 1] : 0.0);
#else
            if (NVR == 1 && n <= 24) { if (warp_cholesky_rows<24>(Hm, ld, n, dinv, lane)) st |= SIHMPC_ST_NOTPD; }
            else if (warp_cholesky<NVR>(Hm, ld, n, dinv, lane)) st |= SIHMPC_ST_NOTPD;
#endif
            double acc[NVR];
            load_acc<NVR>(acc, g, n, lane, -1.0);
            warp_trsv_lower<NVR>(Hm, ld, dinv, n, acc, lane);
            warp_trsv_lower_T<NVR>(Hm, ld, dinv, n, acc, lane);       // acc = Newton direction p
            double dec = 0.0;
#pragma unroll
            for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < n) dec = fma(-g[i], acc[t], dec); }
            dec = warp_sum(dec);
            if (!(dec > dec_tol)) break;
            // ---- backtracking on phi (keeps every d_i > 0)
            double t_ = 1.0, phin = SIH_INF;
            bool ok = false;
            for (int ls = 0; ls < 45; ++ls) {
#pragma unroll
                for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < n) vt[i] = fma(t_, acc[t], v[i]); }
                __syncwarp();
                phin = log_eval<MM>(P, vt, e, r, Vi, d, lane);
                if (mu > 0.0 && phin < SIH_INF) phin += log_barrier(P, vt, uu, d, mu, false, S1, S2, b1, b2, lane, glo, ghi, T1, T2);
                if (phin <= phi - 1e-4 * t_ * dec) { ok = true; break; }
                t_ *= 0.5;
                __syncwarp();
            }
            if (!ok) {                     // rounding level: restore the state of v and stop
                __syncwarp();
                phi = log_eval<MM>(P, v, e, r, Vi, d, lane);
                if (mu > 0.0) phi += log_barrier(P, v, uu, d, mu, false, S1, S2, b1, b2, lane, glo, ghi, T1, T2);
                break;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) v[i] = vt[i];
            const double flat_tol = (mu > 0.0) ? fmin(1e-15 * (1.0 + fabs(phi)), 1e-3 * mu) : 1e-15 * (1.0 + fabs(phi));
            const bool flat = (phi - phin) <= flat_tol;                      // decrease at rounding level: converged
            phi = phin;
            // (phi is self-concordant: after a full step with decrement dec the Hessian is within (1 - sqrt(dec))^-2 of the old one;
            // the barrier terms of the interior-point rounds carry the factor mu)
            near = (t_ == 1.0) && dec <= 1e-4 * (mu > 0.0 ? fmin(mu, 1.0) : 1.0);
            __syncwarp();
            if (flat) break;
        }
        };
        newton(0.0, max_iter);
        if (it >= max_iter) st |= SIHMPC_ST_MAXITER;
        // ---- explicit inequalities of the reference NLP must be strictly inactive
        {
            double marg = SIH_INF;
            if (lane == 0) {
                marg = v[nv] - 1.0;
                for (int i = 0; i < m; ++i) marg = fmin(marg, d[i] - 0.01);
            }
            for (int j = lane; j < nu; j += 32) {
                double cs = uu[j];
                for (int k = 0; k < P.N; ++k) {
                    const double dv = v[k * nu + j];
                    cs += dv;
                    marg = fmin(marg, fmin(cs - P.ulb[j], P.uub[j] - cs));
                    marg = fmin(marg, fmin(dv - P.dulb[j], P.duub[j] - dv));
                }
            }
            for (int rr = lane; rr < ncd; rr += 32) {
                double val = 0.0;
                for (int c = 0; c < nv; ++c) val = fma(P.GcT[(size_t)c * ncd + rr], v[c], val);
                marg = fmin(marg, fmin(val - glo[rr], ghi[rr] - val));
            }
            marg = -warp_max(-marg);
            if (!(marg > 1e-9)) {
                // ---- interior point on the explicit inequalities (see the header): strictly feasible start = no moves
                // (or, when a dense row excludes that, the warm start the caller passed: the shifted plan of the last
                // step sits ON its contraction row, inside the relaxed one), ss large enough for every barrier
                __syncwarp();
                // (a dense row with ONE non-zero, on a ranged residual, is a simple bound on that variable: its anchor is the
                // centre of the range whatever the moves do, and it takes no part in the choice of the moves' anchor)
                auto simple_col = [&](const int rr) {
                    int cnt = 0, col = -1;
                    for (int c = 0; c < nv; ++c) if (P.GcT[(size_t)c * ncd + rr] != 0.0) { ++cnt; col = c; }
                    return (cnt == 1 && col >= nvm) ? col : -1;
                };
                double feas0 = SIH_INF;
                for (int rr = lane; rr < ncd; rr += 32)
                    if (simple_col(rr) < 0) feas0 = fmin(feas0, fmin(0.0 - log_relax(glo[rr], -1.0), log_relax(ghi[rr], 1.0) - 0.0));
                feas0 = -warp_max(-feas0);
                const bool zero_ok = !(feas0 <= 0.0);
                const double ss_free = v[nv];
                __syncwarp();
                for (int i = lane; i < nv; i += 32) { vfree[i] = v[i]; v[i] = (zero_ok && i < nvm) ? 0.0 : vw[i]; }
                if (lane == 0) v[nv] = 2.0;
                __syncwarp();
                for (int rr = lane; rr < ncd; rr += 32) {
                    const int col = simple_col(rr);
                    if (col >= 0) {
                        const double gcv = P.GcT[(size_t)col * ncd + rr], lo = glo[rr], hi = ghi[rr];
                        const double centre = (lo > -SIH_INF && hi < SIH_INF) ? 0.5 * (lo + hi)
                                            : (lo > -SIH_INF ? lo + fmax(1.0, fabs(lo)) : hi - fmax(1.0, fabs(hi)));
                        v[col] = centre / gcv;
                    }
                }
                __syncwarp();
                if (!zero_ok && lane == 0) {
                    // the shifted plan sits ON its contraction row (slack = the 1e-10 relaxation: a barrier curvature of
                    // 1e19 that no factorisation survives): move it, by the least change, to the CENTRES of the bounded
                    // dense rows --  v -= Gc_b^T (Gc_b Gc_b^T)^-1 (Gc_b v - centre)  over the (few) bounded rows b
                    constexpr int MB = 8;
                    int rows[MB], nb = 0;
                    double G[MB][MB + 1];
                    for (int rr = 0; rr < ncd && nb < MB; ++rr)
                        if ((glo[rr] > -SIH_INF || ghi[rr] < SIH_INF) && simple_col(rr) < 0) rows[nb++] = rr;
                    for (int a = 0; a < nb; ++a) {
                        const int ra = rows[a];
                        const double lo = glo[ra], hi = ghi[ra];
                        const double centre = (lo > -SIH_INF && hi < SIH_INF) ? 0.5 * (lo + hi)
                                            : (lo > -SIH_INF ? lo + fmax(1.0, fabs(lo)) : hi - fmax(1.0, fabs(hi)));
                        double val = 0.0;
                        for (int c = 0; c < nv; ++c) val = fma(P.GcT[(size_t)c * ncd + ra], v[c], val);
                        G[a][nb] = val - centre;
                        for (int b = 0; b < nb; ++b) {
                            double acc = 0.0;
                            for (int c = 0; c < nv; ++c) acc = fma(P.GcT[(size_t)c * ncd + ra], P.GcT[(size_t)c * ncd + rows[b]], acc);
                            G[a][b] = acc;
                        }
                    }
                    for (int a = 0; a < nb; ++a) {                 // Gauss-Jordan on the (SPD) Gram matrix
                        const double piv = G[a][a];
                        if (!(fabs(piv) > 1e-300)) continue;
                        for (int b = a; b <= nb; ++b) G[a][b] /= piv;
                        for (int r2 = 0; r2 < nb; ++r2)
                            if (r2 != a) { const double f_ = G[r2][a]; for (int b = a; b <= nb; ++b) G[r2][b] = fma(-f_, G[a][b], G[r2][b]); }
                    }
                    for (int c = 0; c < nv; ++c) {
                        double corr = 0.0;
                        for (int a = 0; a < nb; ++a) corr = fma(P.GcT[(size_t)c * ncd + rows[a]], G[a][nb], corr);
                        v[c] -= corr;
                    }
                }
                __syncwarp();
                // ---- start of the interior-point rounds: from that strictly feasible anchor TOWARDS a target, 95 % of the way
                // to the first bound (all rows linear: the segment stays strictly inside).  Two targets, the better start
                // (lower phi) wins: the minimiser found without the inequalities, and the point this step was entered with
                // (in a closed loop the shifted plan of the last step: near the optimum, but ON its active bounds).  The
                // anchor itself (zero moves) can be far off right after a set-point change (V_i ~ 1e2 gamma_i, ss ~ 1e2) and a
                // barrier method started there creeps along the d_i >= 0.01 wall.
                // SIMPLE bounds first (a ranged residual's row has one non-zero; the move box): the violated coordinate
                // of the target is pulled 5 % inside its range, the others keep their values -- the segment then only
                // shortens for the coupled rows (cumulative inputs, contraction, state boxes)
                double *va = vt;                                // the anchor
                for (int i = lane; i < nv; i += 32) va[i] = v[i];
                __syncwarp();
                auto towards = [&](double *tg, const bool always_shrink) {
                    auto inside = [](const double x, const double lo, const double hi) {
                        const bool fl = lo > -SIH_INF, fh = hi < SIH_INF;
                        const double wl = (fl && fh) ? 0.05 * (hi - lo) : 0.05 * fmax(1.0, fabs(fl ? lo : hi));
                        if (fl && x < lo + wl) return lo + wl;
                        if (fh && x > hi - wl) return hi - wl;
                        return x;
                    };
                    for (int c = lane; c < nvm; c += 32) tg[c] = inside(tg[c], P.dulb[c % nu], P.duub[c % nu]);
                    for (int rr = lane; rr < ncd; rr += 32) {
                        const int col = simple_col(rr);
                        if (col >= 0) {
                            const double gcv = P.GcT[(size_t)col * ncd + rr];
                            const double a_ = glo[rr] / gcv, b_ = ghi[rr] / gcv;
                            tg[col] = inside(tg[col], fmin(a_, b_), fmax(a_, b_));
                        }
                    }
                    __syncwarp();
                    double th = 1.0;
                    auto ratio = [&](const double ca, const double cf, const double b) {
                        const double q = (b - ca) / (cf - ca);
                        if (q >= 0.0 && q < th) th = q;
                    };
                    for (int j = lane; j < nu; j += 32) {
                        const bool hl = P.ulb[j] > -SIH_INF, hh = P.uub[j] < SIH_INF, dl = P.dulb[j] > -SIH_INF, dh = P.duub[j] < SIH_INF;
                        const double lo = hl ? log_relax(P.ulb[j], -1.0) : 0.0, hi = hh ? log_relax(P.uub[j], 1.0) : 0.0;
                        const double dlo = dl ? log_relax(P.dulb[j], -1.0) : 0.0, dhi = dh ? log_relax(P.duub[j], 1.0) : 0.0;
                        double ca = uu[j], cf = uu[j];
                        for (int k = 0; k < P.N; ++k) {
                            const double da = va[k * nu + j], df = tg[k * nu + j];
                            ca += da; cf += df;
                            if (hl && cf <= lo) ratio(ca, cf, lo);
                            if (hh && cf >= hi) ratio(ca, cf, hi);
                            if (dl && df <= dlo) ratio(da, df, dlo);
                            if (dh && df >= dhi) ratio(da, df, dhi);
                        }
                    }
                    for (int rr = lane; rr < ncd; rr += 32) {
                        double ca = 0.0, cf = 0.0;
                        for (int c = 0; c < nv; ++c) { const double gcv = P.GcT[(size_t)c * ncd + rr]; ca = fma(gcv, va[c], ca); cf = fma(gcv, tg[c], cf); }
                        if (glo[rr] > -SIH_INF && cf <= log_relax(glo[rr], -1.0)) ratio(ca, cf, log_relax(glo[rr], -1.0));
                        if (ghi[rr] < SIH_INF && cf >= log_relax(ghi[rr], 1.0)) ratio(ca, cf, log_relax(ghi[rr], 1.0));
                    }
                    double nonfin = 0.0;
                    for (int i = lane; i < nv; i += 32) if (!(fabs(tg[i]) < SIH_INF)) nonfin = 1.0;
                    th = -warp_max(-th);
                    if (th < 1.0 || always_shrink) th *= 0.95;
                    if (warp_max(nonfin) > 0.0 || !(th >= 0.0)) th = 0.0;          // (no usable target: stay at the anchor)
                    __syncwarp();
                    for (int i = lane; i < nv; i += 32) v[i] = fma(th, tg[i] - va[i], va[i]);
                    __syncwarp();
                    // bound push (cf. IPOPT bound_push / bound_frac): a start whose slack is the 1e-10 relaxation -- u ON its
                    // bound with no moves, a shifted plan riding a bound -- gives barrier curvatures of 1e18 that no
                    // factorisation survives; every move and every cumulative input is pushed kappa inside, step by step
                    for (int j = lane; j < nu; j += 32) {
                        const bool hl = P.ulb[j] > -SIH_INF, hh = P.uub[j] < SIH_INF, dl = P.dulb[j] > -SIH_INF, dh = P.duub[j] < SIH_INF;
                        double ku = 1e-3 * fmax(1.0, fmax(hl ? fabs(P.ulb[j]) : 0.0, hh ? fabs(P.uub[j]) : 0.0));
                        if (hl && hh) ku = fmin(ku, 0.25 * (P.uub[j] - P.ulb[j]));
                        double kd = 1e-3 * fmax(1.0, fmax(dl ? fabs(P.dulb[j]) : 0.0, dh ? fabs(P.duub[j]) : 0.0));
                        if (dl && dh) kd = fmin(kd, 0.25 * (P.duub[j] - P.dulb[j]));
                        double cs = uu[j];
                        for (int k = 0; k < P.N; ++k) {
                            double dv = v[k * nu + j];
                            if (dl && dv < P.dulb[j] + kd) dv = P.dulb[j] + kd;
                            if (dh && dv > P.duub[j] - kd) dv = P.duub[j] - kd;
                            const double cn = cs + dv;
                            if (hl && cn < P.ulb[j] + ku) dv += P.ulb[j] + ku - cn;
                            else if (hh && cn > P.uub[j] - ku) dv -= cn - (P.uub[j] - ku);
                            cs += dv;
                            v[k * nu + j] = dv;
                        }
                    }
                    __syncwarp();
                    // ss large enough for every barrier; returns phi there
                    log_eval<MM>(P, v, e, r, Vi, d, lane);
                    double smax = 0.0;
                    for (int i = 0; i < m; ++i) smax = fmax(smax, (Vi[i] + 0.02) / P.gamma[i]);
                    __syncwarp();
                    if (lane == 0) v[nv] = fmax(fmax(ss_free, 1.01), 1.1 * smax + 0.1);
                    __syncwarp();
                    return log_eval<MM>(P, v, e, r, Vi, d, lane);
                };
                {
                    const double m1 = towards(vfree, false);
                    for (int i = lane; i < n; i += 32) g[i] = v[i];
                    __syncwarp();
                    const double m2 = towards(vw, true);
                    if (!(m2 < m1)) {
                        __syncwarp();
                        for (int i = lane; i < n; i += 32) v[i] = g[i];
                        __syncwarp();
                    }
                }
                bool fail_ip = false;
                for (double mu = 1e-1; mu >= 0.99e-14; mu *= 0.1) {
                    phi = log_eval<MM>(P, v, e, r, Vi, d, lane);
                    const double bar = log_barrier(P, v, uu, d, mu, false, S1, S2, b1, b2, lane, glo, ghi, T1, T2);
                    if (!(phi < SIH_INF) || !(bar < SIH_INF)) { fail_ip = true; break; }
                    phi += bar;
                    newton(mu, 60);
                }
                phi = log_eval<MM>(P, v, e, r, Vi, d, lane);         // the cost the reference reports: without the barriers
                if (fail_ip) st |= SIHMPC_ST_ACTIVE_UNSUPPORTED;
            }
        }
        st_or |= st; it_total += it; st_J = phi;
        ss_hint = (st == 0) ? v[nv] : 0.0;
        const bool last_step = (step == n_steps - 1);
        if (last_step) {
            if (io.du) for (int j = lane; j < nu; j += 32) io.du[s * nu + j] = v[j];
            if (w_opt) {
                for (int i = lane; i < nvm; i += 32) w_opt[s * nw + i] = v[i];
                for (int i = lane; i < ny; i += 32) {
                    w_opt[s * nw + nvm + i] = r[P.o_sy + i];
                    w_opt[s * nw + nvm + ny + i] = r[P.o_si + i];
                }
                if (lane == 0) w_opt[s * nw + nw - 1] = v[nv];
            }
            if (lane == 0) {
                if (J) J[s] = phi;
                if (status) status[s] = st;
                if (iters_out) iters_out[s] = it;
            }
        }
        if ((io.flags & SIHMPC_F_UPDATE_VIN) && io.vinant) {          // ViNant = q siN_next^2 (log.py:558-561), every step:
            for (int k = lane; k < P.n_vin; k += 32) {                // the next step's contraction row reads it
                const int i = P.vin_row[k] - P.o_si;
                double a = 0.0;
                for (int j = 0; j < nx; ++j) a = fma(P.Mi1xT[(size_t)j * ny + i], xs[j], a);
                for (int c = 0; c < nv; ++c) a = fma(P.Mi1uT[(size_t)c * ny + i], v[c], a);
                io.vinant[s * P.n_vin + k] = P.vin_q[k] * a * a;
            }
            __syncwarp();
        }
        if (io.traj_du) for (int j = lane; j < nu; j += 32) io.traj_du[((long long)step * batch + s) * nu + j] = v[j];
        if (io.traj_J && lane == 0) io.traj_J[(long long)step * batch + s] = phi;
        __syncwarp();
        if (!advance) continue;
        // ---- plant x+ = [A|B][x ; du], u+ = u + du, measured output, statistics
        for (int j = lane; j < nu; j += 32) xs[nx + j] = v[j];
        __syncwarp();
        {
            const double *pv = io.plant_val ? io.plant_val + (size_t)s * P.pl_nnz : P.pl_val;
            for (int i = lane; i < nx; i += 32) {
                double a = 0.0;
                const int p1 = P.pl_rowptr[i + 1];
                for (int pp = P.pl_rowptr[i]; pp < p1; ++pp) a = fma(pv[pp], xs[P.pl_col[pp]], a);
                xn[i] = a;
            }
        }
        __syncwarp();
        if (io.dist_step && io.dist_step[s] == kabs)
            for (int i = lane; i < ny; i += 32) xn[i] += io.dist[s * ny + i];
        {
            double d2 = 0.0;
            for (int j = lane; j < nu; j += 32) { const double dj = v[j]; uu[j] += dj; d2 = fma(dj, dj, d2); }
            st_du2 += warp_sum(d2);
        }
        __syncwarp();
        if (lane < ny) {
            double a = 0.0;
            for (int j = 0; j < nx; ++j) a = fma(P.CT[(size_t)j * ny + lane], xn[j], a);
            for (int j = 0; j < nu; ++j) a = fma(P.DT[(size_t)j * ny + lane], v[j], a);
            const double err = a - ys[lane];
            st_ise = fma(err, err, st_ise);
            if (io.traj_y) io.traj_y[((long long)step * batch + s) * ny + lane] = a;
        }
        __syncwarp();
        for (int i = lane; i < nx; i += 32) xs[i] = xn[i];
        // ---- warm start of the next step (log.py:518-549): shifted moves, ss = 1
        {
            double sh[NVR];
#pragma unroll
            for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; sh[t] = (i + nu < nvm) ? v[i + nu] : 0.0; }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < nvm) v[i] = sh[t]; }   // (rho keeps its value)
            if (lane == 0) v[nv] = 1.0;
        }
        __syncwarp();
      }  // steps
        if (advance) {
            for (int i = lane; i < nx; i += 32) io.x[s * nx + i] = xs[i];
            for (int i = lane; i < nu; i += 32) io.u[s * nu + i] = uu[i];
            if (w_warm) {
                for (int i = lane; i < nvm; i += 32) w_warm[s * nw + i] = v[i];
                if (lane == 0) w_warm[s * nw + nw - 1] = v[nv];
            }
            if (io.pesos && lane == 0) io.pesos[s] = ss_hint;
        }
        if (io.stats) {
            if (lane < ny) io.stats[s * nstat + lane] = st_ise;
            if (lane == 0) {
                io.stats[s * nstat + ny] = st_du2; io.stats[s * nstat + ny + 2] = st_J;
                io.stats[s * nstat + ny + 3] = (double)st_or; io.stats[s * nstat + ny + 4] = (double)it_total;
            }
        }
        __syncwarp();
    }
}

}  // namespace sih
