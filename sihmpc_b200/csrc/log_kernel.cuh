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
// input and move boxes; log.py:390-392,420-424,435-440) are CHECKED at the minimiser: if one is
// not strictly inactive the scenario is flagged SIHMPC_ST_ACTIVE_UNSUPPORTED instead of returning an
// approximate point (they are inactive in the reference's own example: ss* = sqrt(m/2) > 1).
#pragma once
#include "step_kernel.cuh"

namespace sih {

struct LogLayout { int H, dinv, x, ysp, e, r, v, g, p, gi, d, Vi, vt, total; };

__host__ __device__ inline LogLayout make_log_layout(const DevPlan &P) {
    LogLayout w; int o = 0;
    const int n = P.nv + 1, ld = n | 1;
#define SIH_TAKE(field, cnt) do { w.field = o; o += ((cnt) + 1) & ~1; } while (0)
    SIH_TAKE(H, n * ld); SIH_TAKE(dinv, n); SIH_TAKE(x, P.nx); SIH_TAKE(ysp, P.ny); SIH_TAKE(e, P.R);
    SIH_TAKE(r, P.R); SIH_TAKE(v, n); SIH_TAKE(g, n); SIH_TAKE(p, n); SIH_TAKE(gi, P.m * P.nv);
    SIH_TAKE(d, P.m); SIH_TAKE(Vi, P.m); SIH_TAKE(vt, n);
#undef SIH_TAKE
    w.total = o;
    return w;
}

// residual r = ET v + e, sub-objective values, d_i; returns phi (inf when some d_i <= 0)
__device__ double log_eval(const DevPlan &P, const double *vv, const double *e, double *r, double *Vi,
                           double *d, int lane) {
    const int R = P.R, nv = P.nv, m = P.m;
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

template <int NVR>
__global__ void __launch_bounds__(128)
ihmpc_log_kernel(const DevPlan P, const double *__restrict__ x, const double *__restrict__ u,
                 const double *__restrict__ ysp, const double *__restrict__ w0, const long long batch,
                 double *w_opt, double *J, int *status, int *iters_out, const int max_iter) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * wpb + wib, nwarps = (long long)gridDim.x * wpb;
    const LogLayout lay = make_log_layout(P);
    double *W = smem + (size_t)wib * lay.total;
    double *Hm = W + lay.H, *dinv = W + lay.dinv, *xs = W + lay.x, *ys = W + lay.ysp, *e = W + lay.e;
    double *r = W + lay.r, *v = W + lay.v, *g = W + lay.g, *gi = W + lay.gi, *d = W + lay.d, *Vi = W + lay.Vi;
    double *vt = W + lay.vt;
    const int nx = P.nx, nu = P.nu, ny = P.ny, nv = P.nv, m = P.m, R = P.R;
    const int n = nv + 1, ld = n | 1, nw = nv + 2 * ny + 1;

    for (long long s = gw; s < batch; s += nwarps) {
        for (int i = lane; i < nx; i += 32) xs[i] = x[s * nx + i];
        for (int i = lane; i < ny; i += 32) ys[i] = ysp[s * ny + i];
        for (int i = lane; i < nv; i += 32) v[i] = w0 ? w0[s * nw + i] : 0.0;
        if (lane == 0) v[nv] = w0 ? w0[s * nw + nw - 1] : 1.0;
        __syncwarp();
        for (int k = lane; k < R; k += 32) {
            double a = P.e0[k];
            for (int j = 0; j < nx; ++j) a = fma(P.PxT[(size_t)j * R + k], xs[j], a);
            for (int c = 0; c < ny; ++c) a = fma(P.PyT[(size_t)c * R + k], ys[c], a);
            e[k] = a;
        }
        __syncwarp();
        int st = 0, it = 0;
        double phi = log_eval(P, v, e, r, Vi, d, lane);
        {   // make the start strictly feasible for the barrier: ss >= 1 and ss*gamma_i > V_i
            double smax = 0.0;
            for (int i = 0; i < m; ++i) smax = fmax(smax, Vi[i] / P.gamma[i]);
            const double ss0 = fmax(fmax(v[nv], 1.0), 1.1 * smax + 0.1);
            __syncwarp();
            if (lane == 0) v[nv] = ss0;
            __syncwarp();
            phi = log_eval(P, v, e, r, Vi, d, lane);
        }
        for (it = 1; it <= max_iter; ++it) {
            // ---- gradients a_i = grad_dU V_i (rows of gi), lanes over variables
            for (int c = lane; c < nv; c += 32) {
                for (int i = 0; i < m; ++i) {
                    double a = 0.0;
                    for (int k = 0; k < R; ++k) a = fma(P.E[(size_t)k * nv + c], P.Wrow[(size_t)k * m + i] * r[k], a);
                    gi[i * nv + c] = 2.0 * (a + P.Wdu[(size_t)c * m + i] * v[c]);
                }
                double a = 0.0;
                for (int k = 0; k < R; ++k) a = fma(P.E[(size_t)k * nv + c], P.w0row[k] * r[k], a);
                double gc = 2.0 * a;
                for (int i = 0; i < m; ++i) gc += gi[i * nv + c] / d[i];
                g[c] = gc;
            }
            if (lane == 0) {
                double gs = 2.0 * v[nv];
                for (int i = 0; i < m; ++i) gs -= P.gamma[i] / d[i];
                g[nv] = gs;
            }
            __syncwarp();
            // ---- Hessian (lower triangle)
            for (int idx = lane; idx < n * n; idx += 32) {
                const int a = idx / n, b = idx % n;
                if (b > a) continue;
                double h;
                if (a < nv) {
                    h = P.Ht[a * nv + b];
                    for (int q = 0; q < P.mh; ++q) h = fma(1.0 / d[P.hj_index[q]], P.HJ[((size_t)q * nv + a) * nv + b], h);
                    for (int i = 0; i < m; ++i) {
                        const double di = d[i];
                        h = fma(gi[i * nv + a], gi[i * nv + b] / (di * di), h);
                        if (a == b) h = fma(2.0 * P.Wdu[(size_t)a * m + i], 1.0 / di, h);
                    }
                } else if (b < nv) {
                    h = 0.0;
                    for (int i = 0; i < m; ++i) h -= P.gamma[i] * gi[i * nv + b] / (d[i] * d[i]);
                } else {
                    h = 2.0;
                    for (int i = 0; i < m; ++i) h += P.gamma[i] * P.gamma[i] / (d[i] * d[i]);
                }
                Hm[a * ld + b] = h;
            }
            __syncwarp();
            if (warp_cholesky<NVR>(Hm, ld, n, dinv, lane)) st |= SIHMPC_ST_NOTPD;
            double acc[NVR];
            load_acc<NVR>(acc, g, n, lane, -1.0);
            warp_trsv_lower<NVR>(Hm, ld, dinv, n, acc, lane);
            warp_trsv_lower_T<NVR>(Hm, ld, dinv, n, acc, lane);       // acc = Newton direction p
            double dec = 0.0;
#pragma unroll
            for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < n) dec = fma(-g[i], acc[t], dec); }
            dec = warp_sum(dec);
            if (!(dec > 1e-18 * (1.0 + fabs(phi)))) break;
            // ---- backtracking on phi (keeps every d_i > 0)
            double t_ = 1.0, phin = SIH_INF;
            bool ok = false;
            for (int ls = 0; ls < 45; ++ls) {
#pragma unroll
                for (int t = 0; t < NVR; ++t) { const int i = lane + 32 * t; if (i < n) vt[i] = fma(t_, acc[t], v[i]); }
                __syncwarp();
                phin = log_eval(P, vt, e, r, Vi, d, lane);
                if (phin <= phi - 1e-4 * t_ * dec) { ok = true; break; }
                t_ *= 0.5;
                __syncwarp();
            }
            if (!ok) {                     // rounding level: restore the state of v and stop
                __syncwarp();
                phi = log_eval(P, v, e, r, Vi, d, lane);
                break;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) v[i] = vt[i];
            phi = phin;
            __syncwarp();
        }
        if (it > max_iter) st |= SIHMPC_ST_MAXITER;
        // ---- explicit inequalities of the reference NLP must be strictly inactive
        {
            double marg = SIH_INF;
            if (lane == 0) {
                marg = v[nv] - 1.0;
                for (int i = 0; i < m; ++i) marg = fmin(marg, d[i] - 0.01);
            }
            for (int j = lane; j < nu; j += 32) {
                double cs = u[s * nu + j];
                for (int k = 0; k < P.N; ++k) {
                    const double dv = v[k * nu + j];
                    cs += dv;
                    marg = fmin(marg, fmin(cs - P.ulb[j], P.uub[j] - cs));
                    marg = fmin(marg, fmin(dv - P.dulb[j], P.duub[j] - dv));
                }
            }
            marg = -warp_max(-marg);
            if (!(marg > 1e-9)) st |= SIHMPC_ST_ACTIVE_UNSUPPORTED;
        }
        for (int i = lane; i < nv; i += 32) w_opt[s * nw + i] = v[i];
        for (int i = lane; i < ny; i += 32) {
            w_opt[s * nw + nv + i] = r[P.o_sy + i];
            w_opt[s * nw + nv + ny + i] = r[P.o_si + i];
        }
        if (lane == 0) {
            w_opt[s * nw + nw - 1] = v[nv];
            if (J) J[s] = phi;
            if (status) status[s] = st;
            if (iters_out) iters_out[s] = it;
        }
        __syncwarp();
    }
}

}  // namespace sih
