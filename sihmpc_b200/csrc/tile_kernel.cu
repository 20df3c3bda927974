// tile_kernel.cu -- fast path v4 for MID-SIZE controllers (N*nu + 1 <= 64, any nx / R / m):
// CTA-cooperative, everything that multiplies on the FP64 TENSOR pipe (DMMA.8x8x4).
//
// One CTA = 8 warps = 8 scenarios (an "octet"), state resident in shared memory for all n_steps.
//
//   * every product with a SHARED controller matrix is  out^T[scenario][row] = in^T[scenario][k] M^T
//     with M = 8 scenarios: mma.sync.m8n8k4.f64, A fragment = the scenarios' vectors in shared
//     memory ([scenario][k], stride == 4 mod 16 doubles: conflict free), B fragment = the constant
//     matrix, stored on the host as TILE-CSR (only the non-zero 8x4 tiles, already in B-fragment
//     order: one coalesced 256-byte load per MMA, served by L1/L2 -- all CTAs stream the same
//     constants).  Output tiles are dealt round-robin to the 8 warps.  The Hessian
//     H = Ht + sum_a pesos_a HJ_a + 2 diag(Wdu pesos) is ONE such product whose output rows are
//     the packed 8x8 tiles of the factorisation (k-space [pesos ; 1]);
//   * the factorisation is per scenario (warp w <-> scenario w): right-looking Cholesky on 8x8
//     tiles in shared memory, panels of 4 columns.  The 4x4 pivot block is factorised and inverted
//     redundantly by every lane (W = L11^-1); the panel solve L21 = A21 W^T and every trailing
//     update C -= L21 L21^T is a DMMA whose A and B fragments are the SAME register (one LDS.64 per
//     row tile), the C fragment is the tile in its natural row-major order.  The right-hand side
//     -f rides along as row nv of the matrix, so the forward substitution is free; the backward
//     substitution uses the stored W blocks (no sequential chain inside a block);
//   * plant update (CSR, optional per-scenario values), bound check, weights: warp per scenario.
//
// Like the other fast paths it computes the UNCONSTRAINED optimum and hands a scenario whose
// optimum violates a bound over to the general kernel (start_step[]), so the loop stays exact.
//
// Reference semantics replaced: the same as step_kernel.cuh (ihmpccontroller.py:567-604, 534-564,
// 619-636, 93-113; example_distillation.py:81-113).
#include <cuda_runtime.h>
#include <vector>
#include <functional>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "devplan.h"
#include <math.h>
#include "tile_api.h"


namespace sih {

namespace {
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SIH_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SIH_FULL, v, o));
    return v;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 1/sqrt(d) for a positive normal d: MUFU.RSQ64H seed + two Newton steps (see dist_kernel.cuh)
__device__ __forceinline__ double fast_rsqrt(const double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double t = d * y;
    double e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = d * y;
    e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}
}  // namespace

__host__ __device__ __forceinline__ int tl_tid(int I, int J) { return (I * (I + 1)) / 2 + J; }

// accumulators of output tile T of one tile-ELL matrix; in_q = in + q*S + g (A-fragment base of this lane).
// A row is the list of its non-zero 8x4 tiles: kof[t] = first k of stored tile t (the A address), the tile itself in
// B-fragment order (one coalesced 256-byte load per MMA).  The host pads every list to a multiple of four with zero
// tiles, so the loop is one body of four MMAs on four independent accumulator pairs, no remainder, no run
// bookkeeping; the B tiles and A offsets of the next group are fetched while this one multiplies.
__device__ __forceinline__ void tile_row(const TileCfg &C, const int poff, const int T, const double *in_q,
                                         const int lane, double &r0, double &r1) {
    const int t0 = __ldg(C.ptr + poff + T), n = __ldg(C.ptr + poff + T + 1) - t0;
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
    if (n > 0) {
        const double *Bp = C.B + (size_t)t0 * 32 + lane;
        const int4 *kp = reinterpret_cast<const int4 *>(C.kof + t0);          // lists start on multiples of four
        int4 ko = __ldg(kp);
        double b0 = __ldg(Bp), b1 = __ldg(Bp + 32), b2 = __ldg(Bp + 64), b3 = __ldg(Bp + 96);
        for (int t = 4; ; t += 4) {
            const bool more = t < n;
            int4 nk = ko;
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
            if (more) {
                nk = __ldg(kp + (t >> 2));
                n0 = __ldg(Bp + t * 32); n1 = __ldg(Bp + t * 32 + 32); n2 = __ldg(Bp + t * 32 + 64); n3 = __ldg(Bp + t * 32 + 96);
            }
            const double a0 = in_q[ko.x], a1 = in_q[ko.y], a2 = in_q[ko.z], a3 = in_q[ko.w];
            dmma884(c0, c1, a0, b0);
            dmma884(d0, d1, a1, b1);
            dmma884(e0, e1, a2, b2);
            dmma884(f0, f1, a3, b3);
            if (!more) break;
            ko = nk; b0 = n0; b1 = n1; b2 = n2; b3 = n3;
        }
    }
    r0 = (c0 + d0) + (e0 + f0); r1 = (c1 + d1) + (e1 + f1);
}

__device__ __forceinline__ double tl_rsqrt(const double d) { return fast_rsqrt(d); }

// Right-looking BLOCK elimination (block LDL^T with 4x4 pivot blocks) of the augmented matrix
// [H ; b^T] of ONE scenario by ONE warp.  A: packed lower 8x8 tiles (row-major 64 doubles each) in
// shared memory.  For every panel of four columns:  M = A11^-1 (closed form through two 2x2 blocks:
// two reciprocals, no square roots),  X = A21 M  (one DMMA per row tile),  C -= X A21^T (one DMMA per
// trailing tile: A fragment = X, B fragment = the panel itself).  H = L D L^T with L = X unit block
// lower and D = blockdiag(A11), so with the right-hand side riding along as row nv:
//     w = D^-1 L^-1 b = the X values of row nv,      z_blk = w_blk - sum_{i below} X[i][blk] z_i.
// On exit the strictly-lower part of A holds X and yv[0..nv) = w.
//
// Everything is compile-time (NT tile rows, panel P): the trailing matrix lives in REGISTERS as MMA C
// fragments (c[tile]), so a trailing update is one DMMA and nothing else; only the four columns of the
// current panel go through shared memory (C layout -> A/B layout).
__device__ __forceinline__ double tl_rcp(const double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

template <int NT, int P>
__device__ __forceinline__ void tfs_panel(double2 (&c)[NT * (NT + 1) / 2], double *A, double *yv, const int nv,
                                          const int rb, const int lane, int &notpd) {
    if constexpr (P < 2 * NT) {
        constexpr int J = P >> 1, c0 = (P & 1) * 4, Ib = NT - 1;
        if (4 * P < nv) {                                              // uniform
            const int q = lane >> 2, g = lane & 3;
            // (a) the panel's four columns of every tile (I,J): registers -> home
            if ((g >> 1) == (c0 >> 2)) {
#pragma unroll
                for (int I = J; I < NT; ++I)
                    *reinterpret_cast<double2 *>(A + (I * (I + 1) / 2 + J) * 64 + q * 8 + 2 * g) = c[I * (I + 1) / 2 + J];
            }
            __syncwarp();
            const int nreal = (nv - 4 * P < 4) ? (nv - 4 * P) : 4;        // columns >= nreal: identity
            // the panel in A/B layout (rows of the pivot block itself do not take part)
            double pl[NT];
#pragma unroll
            for (int I = J; I < NT; ++I) {
                double v = A[(I * (I + 1) / 2 + J) * 64 + q * 8 + c0 + g];
                if (I == J && q < c0 + nreal) v = 0.0;
                pl[I] = v;
            }
            // (b) M = A11^-1, redundantly in every lane
            const double *D = A + (J * (J + 1) / 2 + J) * 64 + c0 * 8 + c0;   // element (r,c) at D[r*8+c]
            double a00 = D[0], a10 = D[8], a11 = D[9], a20 = D[16], a21 = D[17], a22 = D[18];
            double a30 = D[24], a31 = D[25], a32 = D[26], a33 = D[27];
            if (nreal < 4) { a30 = a31 = a32 = 0.0; a33 = 1.0; }
            if (nreal < 3) { a20 = a21 = 0.0; a22 = 1.0; }
            if (nreal < 2) { a10 = 0.0; a11 = 1.0; }
            double detp = fma(a00, a11, -a10 * a10);
            if (!(a00 > 0.0) || !(detp > 0.0)) { notpd = 1; detp = 1.0; }
            const double ip = tl_rcp(detp);
            const double p00 = a11 * ip, p10 = -a10 * ip, p11 = a00 * ip;
            const double t00 = fma(a20, p00, a21 * p10), t01 = fma(a20, p10, a21 * p11);
            const double t10 = fma(a30, p00, a31 * p10), t11 = fma(a30, p10, a31 * p11);
            const double s00 = a22 - fma(t00, a20, t01 * a21);
            const double s10 = a32 - fma(t10, a20, t11 * a21);
            const double s11 = a33 - fma(t10, a30, t11 * a31);
            double dets = fma(s00, s11, -s10 * s10);
            if (!(s00 > 0.0) || !(dets > 0.0)) { notpd = 1; dets = 1.0; }
            const double is = tl_rcp(dets);
            const double m22 = s11 * is, m32 = -s10 * is, m33 = s00 * is;
            const double u00 = fma(m22, t00, m32 * t10), u01 = fma(m22, t01, m32 * t11);
            const double u10 = fma(m32, t00, m33 * t10), u11 = fma(m32, t01, m33 * t11);
            const double m00 = p00 + fma(t00, u00, t10 * u10);
            const double m10 = p10 + fma(t01, u00, t11 * u10);
            const double m11 = p11 + fma(t01, u01, t11 * u11);
            double mqg;                                                // M[q][g] (symmetric), zero for q >= 4
            {
                const double r0 = (g == 0) ? m00 : ((g == 1) ? m10 : ((g == 2) ? -u00 : -u10));
                const double r1 = (g == 0) ? m10 : ((g == 1) ? m11 : ((g == 2) ? -u01 : -u11));
                const double r2 = (g == 0) ? -u00 : ((g == 1) ? -u01 : ((g == 2) ? m22 : m32));
                const double r3 = (g == 0) ? -u10 : ((g == 1) ? -u11 : ((g == 2) ? m32 : m33));
                mqg = (q == 0) ? r0 : ((q == 1) ? r1 : ((q == 2) ? r2 : ((q == 3) ? r3 : 0.0)));
            }
            // (c) X = A21 M : results (C layout, lanes g<2) back home
            {
                double o[NT][2];
#pragma unroll
                for (int I = J; I < NT; ++I) { o[I][0] = o[I][1] = 0.0; dmma884(o[I][0], o[I][1], pl[I], mqg); }
                if (g < 2) {
#pragma unroll
                    for (int I = J; I < NT; ++I)
                        if (!(I == J && q < c0 + nreal))
                            *reinterpret_cast<double2 *>(A + (I * (I + 1) / 2 + J) * 64 + q * 8 + c0 + 2 * g) = make_double2(o[I][0], o[I][1]);
                }
            }
            __syncwarp();
            if (lane < nreal) yv[4 * P + lane] = A[(Ib * (Ib + 1) / 2 + J) * 64 + rb * 8 + c0 + lane];
            // (d) trailing update on the register tiles: c(I,K) -= X_I A21_K^T
            if constexpr (P + 1 < 2 * NT) {
                double xa[NT];
#pragma unroll
                for (int I = J; I < NT; ++I) {
                    double v = A[(I * (I + 1) / 2 + J) * 64 + q * 8 + c0 + g];
                    if (I == J && q < c0 + 4) v = 0.0;
                    xa[I] = -v;
                }
#pragma unroll
                for (int K = J + (c0 ? 1 : 0); K < NT; ++K) {
#pragma unroll
                    for (int I = K; I < NT; ++I)
                        dmma884(c[I * (I + 1) / 2 + K].x, c[I * (I + 1) / 2 + K].y, xa[I], pl[K]);
                }
            }
        }
        tfs_panel<NT, P + 1>(c, A, yv, nv, rb, lane, notpd);
    }
}

template <int NT, bool ADD_DIAG = true>
__device__ __forceinline__ int tile_factor(double *A, const double *omdu_w, const int nv, const int rb,
                                           const int lane, double *yv) {
    constexpr int NTILES = NT * (NT + 1) / 2;
    const int q = lane >> 2, g = lane & 3;
    double2 c[NTILES];
#pragma unroll
    for (int t = 0; t < NTILES; ++t) c[t] = *reinterpret_cast<const double2 *>(A + t * 64 + q * 8 + 2 * g);
    // + 2 omega_du on the diagonal (ihmpccontroller.py:200,212: the du sub-objectives)
    if constexpr (ADD_DIAG) {
#pragma unroll
        for (int J = 0; J < NT; ++J) {
            const double od = 2.0 * omdu_w[8 * J + q];               // zero beyond nv
            if (q == 2 * g) c[J * (J + 1) / 2 + J].x += od;
            if (q == 2 * g + 1) c[J * (J + 1) / 2 + J].y += od;
        }
    }
    int notpd = 0;
    tfs_panel<NT, 0>(c, A, yv, nv, rb, lane, notpd);
    return notpd;
}

// backward substitution  z_blk = w_blk - sum_{i below} X[i][blk] z_i ; zv[nv..8NT) must be zero
template <int NT, int P>
__device__ __forceinline__ void tbs_panel(const double *A, const int nv, const int lane, const double *yv, double *zv) {
    if constexpr (P >= 0) {
        constexpr int J = P >> 1, c0 = (P & 1) * 4;
        if (4 * P < nv) {
            const int rq = lane >> 2, cg = lane & 3;
            double acc = 0.0;
            constexpr int NJ = c0 ? (NT - J - 1) : (NT - J);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                // rows i = 4P + 4 + rq + 8j
                int T, row;
                if (c0) { T = J + 1 + j; row = rq; }
                else { T = J + j + (rq >> 2); row = (rq + 4) & 7; }
                if (T < NT) acc = fma(A[(T * (T + 1) / 2 + J) * 64 + row * 8 + c0 + cg], zv[8 * T + row], acc);
            }
            acc += __shfl_xor_sync(SIH_FULL, acc, 4);
            acc += __shfl_xor_sync(SIH_FULL, acc, 8);
            acc += __shfl_xor_sync(SIH_FULL, acc, 16);
            if (rq == 0 && 4 * P + cg < nv) zv[4 * P + cg] = yv[4 * P + cg] - acc;
            __syncwarp();
        }
        tbs_panel<NT, P - 1>(A, nv, lane, yv, zv);
    }
}

// Hessian product: packed lower tiles of  sum_k a_k X_k  for the eight scenarios of the CTA (k-space [1 ; pesos]),
// BHm = the blocks in B-fragment order (du space: Ht, HJ_a; bound space: S^T (Ht | HJ_a + 2 diag Wdu_a) S)
template <int KH>
__device__ __forceinline__ void tile_hessian_k(const TileCfg &C, const double *__restrict__ BHm,
                                               const double *pes, double *Hs, const int w, const int lane) {
    const int q = lane >> 2, g = lane & 3;
    for (int u = w; u < 2 * C.ntile; u += TL_WARPS) {
        const int tid = u >> 1, r0 = (u & 1) * 4;
        const int nrow = __ldg(C.hrows + tid);
        if (r0 >= nrow) continue;
        // every row of a tile is stored (rows >= nrow hold zeros), so all 4 KH loads go out before the first MMA
        const double *Bp = BHm + ((size_t)(tid * 8 + r0) * KH) * 32 + lane;
        double b[4][KH], a[KH];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int t = 0; t < KH; ++t) b[rr][t] = __ldg(Bp + (rr * KH + t) * 32);
#pragma unroll
        for (int t = 0; t < KH; ++t) a[t] = pes[q * C.S_p + 4 * t + g];
        double h[4][2];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            h[rr][0] = h[rr][1] = 0.0;
#pragma unroll
            for (int t = 0; t < KH; ++t) dmma884(h[rr][0], h[rr][1], a[t], b[rr][t]);
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
            if (r0 + rr < nrow)
                *reinterpret_cast<double2 *>(Hs + (size_t)q * C.HS + tid * 64 + (r0 + rr) * 8 + 2 * g) = make_double2(h[rr][0], h[rr][1]);
    }
}

__device__ __forceinline__ void tile_hessian(const TileCfg &C, const double *__restrict__ BHm, const int KHm,
                                             const double *pes, double *Hs, const int w, const int lane) {
    switch (KHm) {                                     // uniform; k-tiles of [1 ; pesos] that carry Hessian blocks
        case 1: tile_hessian_k<1>(C, BHm, pes, Hs, w, lane); break;
        case 2: tile_hessian_k<2>(C, BHm, pes, Hs, w, lane); break;
        case 3: tile_hessian_k<3>(C, BHm, pes, Hs, w, lane); break;
        case 4: tile_hessian_k<4>(C, BHm, pes, Hs, w, lane); break;
        case 5: tile_hessian_k<5>(C, BHm, pes, Hs, w, lane); break;
        default: tile_hessian_k<TL_KHMAX>(C, BHm, pes, Hs, w, lane); break;
    }
}

// MODE 0: first pass (unconstrained optimum + bound check; a scenario that violates a bound is appended to `list`).
// MODE 1: resume pass over the listed scenarios, eight per CTA, each from its own start_step, with the in-kernel
//         active set below; what it cannot resolve is flagged in start_step / n_handover for the general kernel.
template <int NT, int MODE>
__global__ void __launch_bounds__(TL_WARPS * 32, 1)
tile_step_kernel(const DevPlan P, const TileCfg C, const sihmpc_io io, const long long batch, const int n_steps,
                 int *__restrict__ start_step, int *__restrict__ n_handover, int *__restrict__ n_inkernel,
                 int *__restrict__ list, const int *__restrict__ n_list,
                 const unsigned long long *__restrict__ warm_mask) {
    constexpr bool RESUME = (MODE == 1);
    extern __shared__ double sm[];
    __shared__ long long s_sid[TL_WARPS];
    __shared__ int s_stp[TL_WARPS], s_len[TL_WARPS];
    long long n_work = batch;
    if constexpr (RESUME) {
        n_work = (long long)(*n_list) - 0x7f7f7f7f;
        if ((long long)blockIdx.x * TL_WARPS >= n_work) return;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, q = lane >> 2, g = lane & 3;
    const int nx = P.nx, nu = P.nu, ny = P.ny, nv = P.nv, m = P.m, R = P.R, Ny = P.Ny;
    const int nstat = ny + SIHMPC_NSTATS_EXTRA;
    const bool advance = (io.flags & SIHMPC_F_ADVANCE) != 0;
    double *Hs = sm + C.o_H, *sv = sm + C.o_sv, *eb = sm + C.o_e, *omb = sm + C.o_om, *tb = sm + C.o_t;
    double *omdu = sm + C.o_omdu, *pes = sm + C.o_pes, *sq = sm + C.o_sq, *vn = sm + C.o_vn;
    double *scb = sm + C.o_sc;
    // views of the scenario this warp owns in the per-scenario phases
    double *Hw = Hs + (size_t)w * C.HS, *svw = sv + w * C.S_sv, *ew = eb + w * C.S_r, *omw = omb + w * C.S_r;
    double *omduw = omdu + w * C.S_v, *pw = pes + w * C.S_p, *sqw = sq + w * C.S_q, *vnw = vn + w * C.S_m;
    double *zw = sm + C.o_z + w * 64, *yw = sm + C.o_y + w * 64, *scw = scb + w * C.S_sc;
    double *uw = scw + 2, *stw = scw + 2 + nu;
    double *xtw = omw;                                // plant scratch: omega is dead after the cost
    double *tbw = tb + w * C.S_r;
    const bool shared_plant = (io.plant_val == nullptr) && C.n_x > 0;
    const int o_ysp = nx + nv, o_one = nx + nv + ny;
    const int nsq = Ny + 2 * ny;

    // active set scratch of this scenario: the point lives in the squares buffer (only used after the solve), the
    // gradient in the forward-substitution output (dead outside factorisation + back substitution)
    double *cw = sqw, *gw = yw;
    const long long n_oct = (n_work + TL_WARPS - 1) / TL_WARPS;
    for (long long oct = blockIdx.x; oct < n_oct; oct += gridDim.x) {
        // every slot starts from clean buffers: padding entries are never rewritten inside the loop, and a
        // scenario that was handed over keeps stepping unconstrained in its slot and may leave Inf/NaN there
        for (int i = threadIdx.x; i < C.total; i += blockDim.x) sm[i] = 0.0;
        __syncthreads();
        long long s = oct * TL_WARPS + w;
        const bool valid = s < n_work;
        int st0 = 0, n_iter = n_steps;
        if constexpr (RESUME) {
            s = valid ? list[s] : batch - 1;
            st0 = valid ? start_step[s] : n_steps;
            if (st0 < 0 || st0 > n_steps) st0 = n_steps;
            if (lane == 0) s_len[w] = n_steps - st0;
            __syncthreads();
            n_iter = 0;
#pragma unroll
            for (int i = 0; i < TL_WARPS; ++i) n_iter = max(n_iter, s_len[i]);
            if (valid && lane == 0) start_step[s] = 0x7f7f7f7f;      // resolved here unless flagged again below
        }
        const long long sl = valid ? s : batch - 1;
        bool alive = valid;
        unsigned long long mlo = 0ull, mhi = 0ull;     // bound masks of the last step (warm start of the active set)
        if constexpr (RESUME) {                        // scenario sent back by the general kernel: its last active set
            if (warm_mask && valid) { mlo = warm_mask[2 * s]; mhi = warm_mask[2 * s + 1]; }
        }
        if (lane == 0) s_sid[w] = s;
        // ---------------- load the scenario
        for (int i = lane; i < nx; i += 32) svw[i] = io.x[sl * nx + i];
        for (int i = lane; i < nv; i += 32) svw[nx + i] = 0.0;
        for (int i = lane; i < m; i += 32) pw[1 + i] = io.pesos[sl * m + i];     // k-space [1 ; pesos]
        for (int i = lane; i < nu; i += 32) uw[i] = io.u[sl * nu + i];
        for (int i = lane; i < nstat; i += 32) stw[i] = io.stats ? io.stats[sl * nstat + i] : 0.0;
        int seg = 0, next_start = 0x7fffffff, dstep = -1;
        if (io.ysp_seq) {
            for (int a = 1; a < io.n_seg; ++a)
                if (io.seg_start[sl * io.n_seg + a] <= io.k0) seg = a;
            if (seg + 1 < io.n_seg) next_start = io.seg_start[sl * io.n_seg + seg + 1];
        }
        if (io.dist_step) dstep = io.dist_step[sl];
        for (int i = lane; i < ny; i += 32)
            svw[o_ysp + i] = io.ysp_seq ? io.ysp_seq[(sl * io.n_seg + seg) * ny + i] : io.ysp[sl * ny + i];
        if (lane == 0) { svw[o_one] = 1.0; pw[0] = 1.0; scw[1] = alive ? 1.0 : 0.0; }
        const double *pv = io.plant_val ? io.plant_val + (size_t)sl * P.pl_nnz : P.pl_val;

        auto write_state = [&]() {
            for (int i = lane; i < nx; i += 32) io.x[s * nx + i] = svw[i];
            for (int i = lane; i < nu; i += 32) io.u[s * nu + i] = uw[i];
            for (int i = lane; i < m; i += 32) io.pesos[s * m + i] = pw[1 + i];
            if (io.stats) for (int i = lane; i < nstat; i += 32) io.stats[s * nstat + i] = stw[i];
        };

        for (int it = 0; it < n_iter; ++it) {
            const int step = st0 + it;
            if constexpr (RESUME) {
                if (step >= n_steps && alive) { alive = false; if (lane == 0) scw[1] = 0.0; }
            }
            if (lane == 0) s_stp[w] = step;
            const int kabs = io.k0 + step;
            if (kabs >= next_start) {                     // rare: the schedule switches segment
                while (seg + 1 < io.n_seg && io.seg_start[sl * io.n_seg + seg + 1] <= kabs) ++seg;
                next_start = (seg + 1 < io.n_seg) ? io.seg_start[sl * io.n_seg + seg + 1] : 0x7fffffff;
                for (int i = lane; i < ny; i += 32) svw[o_ysp + i] = io.ysp_seq[(sl * io.n_seg + seg) * ny + i];
            }
            __syncthreads();                              // [A] state of all 8 scenarios visible
            // ---------------- P1: e = Px x + Py ysp + e0 (+ xi(N+1) rows), omega, t = omega o e ; omega_du ; H
            for (int T = w; T < C.n_r; T += TL_WARPS) {
                double e0, e1, o0, o1;
                tile_row(C, C.me, T, sv + q * C.S_sv + g, lane, e0, e1);
                tile_row(C, C.mw, T, pes + q * C.S_p + g, lane, o0, o1);
                const int off = q * C.S_r + 8 * T + 2 * g;
                *reinterpret_cast<double2 *>(eb + off) = make_double2(e0, e1);
                *reinterpret_cast<double2 *>(omb + off) = make_double2(o0, o1);
                *reinterpret_cast<double2 *>(tb + off) = make_double2(o0 * e0, o1 * e1);
            }
            for (int T = (w + 5) & 7; T < C.n_v; T += TL_WARPS) {
                double o0, o1;
                tile_row(C, C.mdu, T, pes + q * C.S_p + g, lane, o0, o1);
                *reinterpret_cast<double2 *>(omdu + q * C.S_v + 8 * T + 2 * g) = make_double2(o0, o1);
            }
            // Hessian tiles  H = Ht + sum_a pesos_a HJ_a : dense k-tiles [0,KH) of [1 ; pesos], half a tile per unit
            tile_hessian(C, C.BH, C.KH, pes, Hs, w, lane);
            __syncthreads();                              // [B]
            // ---------------- P2: right-hand side b = -2 E^T t into row nv of the augmented matrix
            for (int T = w; T < C.n_f; T += TL_WARPS) {
                double f0, f1;
                tile_row(C, C.mf, T, tb + q * C.S_r + g, lane, f0, f1);
                *reinterpret_cast<double2 *>(Hs + (size_t)q * C.HS + tl_tid(C.Ib, T) * 64 + C.rb * 8 + 2 * g) = make_double2(f0, f1);
            }
            __syncthreads();                              // [C]
            // ---------------- P3 (warp per scenario): factorise, solve, check the boxes
            const int notpd = tile_factor<NT>(Hw, omduw, nv, C.rb, lane, yw);
            tbs_panel<NT, 2 * NT - 1>(Hw, nv, lane, yw, zw);
            for (int i = lane; i < nv; i += 32) svw[nx + i] = zw[i];
            // viol = max over rows of (violation - 1e-10 (1 + |bound|)): a bound is violated when viol > 0 (the relative
            // test of the general kernel without its divisions)
            double viol = 0.0, viol_abs = 0.0;
            if (P.has_box)
                for (int i = lane; i < nv; i += 32) {
                    const int j = i % nu;
                    const double lo = P.dulb[j], hi = P.duub[j], val = zw[i];
                    if (lo > -SIH_INF) { const double a = lo - val; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(lo), a)); }
                    if (hi < SIH_INF) { const double a = val - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(hi), a)); }
                }
            if (P.has_cum)
                for (int i = lane; i < nv; i += 32) {
                    const int j = i % nu;
                    double val = 0.0;
                    for (int c = j; c <= i; c += nu) val += zw[c];
                    const double lo = P.ulb[j] - uw[j], hi = P.uub[j] - uw[j];
                    if (lo > -SIH_INF) { const double a = lo - val; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(lo), a)); }
                    if (hi < SIH_INF) { const double a = val - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(hi), a)); }
                }
            viol = warp_max(viol);
            viol_abs = warp_max(viol_abs);
            int nact = 0;
            if constexpr (MODE != 0) {
                if (C.cspace >= 0) {
#include "tile_active_set.inc"
                }
            }
            if (alive && (viol > 0.0 || notpd)) {
                alive = false;                            // hand over from this step on (MODE 0: resume pass; else general kernel)
                mlo = mhi = 0ull;
                if (lane == 0) {
                    start_step[s] = step; scw[1] = 0.0;
                    const int slot = atomicAdd(n_handover, 1) - 0x7f7f7f7f;
                    if constexpr (MODE == 0) list[slot] = (int)s;
                }
                write_state();
            }
            if (lane == 0) scw[0] = 0.0;
            __syncthreads();                              // [D]
            // ---------------- P4: r = e + [E ; Mi1u] dU ; weighted squares
            for (int T = w; T < C.n_r; T += TL_WARPS) {
                double r0, r1;
                tile_row(C, C.mr, T, sv + q * C.S_sv + g, lane, r0, r1);
                const int off = q * C.S_r + 8 * T + 2 * g;
                const double2 e2 = *reinterpret_cast<const double2 *>(eb + off);
                const double2 o2 = *reinterpret_cast<const double2 *>(omb + off);
                r0 += e2.x; r1 += e2.y;
                *reinterpret_cast<double2 *>(eb + off) = make_double2(r0, r1);
                double part = fma(o2.x, r0 * r0, o2.y * r1 * r1);
                part += __shfl_xor_sync(SIH_FULL, part, 1);
                part += __shfl_xor_sync(SIH_FULL, part, 2);
                if (g == 0) atomicAdd(scb + q * C.S_sc, part);
            }
            if (shared_plant && advance)                  // x+ = [A|B][x ; du] on the tensor pipe (t is dead)
                for (int T = (w + 2) & 7; T < C.n_x; T += TL_WARPS) {
                    double x0, x1;
                    tile_row(C, C.mpl, T, sv + q * C.S_sv + g, lane, x0, x1);
                    *reinterpret_cast<double2 *>(tb + q * C.S_r + 8 * T + 2 * g) = make_double2(x0, x1);
                }
            __syncthreads();                              // [E]
            // ---------------- P5 (warp per scenario): cost, outputs, plant, shifted-plan inputs
            double Jval;
            {
                double a = 0.0;
                for (int j = lane; j < nv; j += 32) a = fma(omduw[j], zw[j] * zw[j], a);
                Jval = warp_sum(a) + scw[0];
            }
            const bool last_step = (step == n_steps - 1);
            if (alive) {
                if (last_step) {
                    if (io.du) for (int j = lane; j < nu; j += 32) io.du[s * nu + j] = zw[j];
                    if (io.w_opt) {
                        const int nw = nv + 2 * ny;
                        for (int i = lane; i < nv; i += 32) io.w_opt[s * nw + i] = zw[i];
                        for (int i = lane; i < ny; i += 32) {
                            io.w_opt[s * nw + nv + i] = ew[P.o_sy + i];
                            io.w_opt[s * nw + nv + ny + i] = ew[P.o_si + i];
                        }
                    }
                    if (lane == 0) {
                        if (io.J) io.J[s] = Jval;
                        if (io.status) io.status[s] = 0;
                        if (io.n_active) io.n_active[s] = nact;
                    }
                }
                if (io.traj_du) for (int j = lane; j < nu; j += 32) io.traj_du[((long long)step * batch + s) * nu + j] = zw[j];
                if (io.traj_J && lane == 0) io.traj_J[(long long)step * batch + s] = Jval;
                if ((io.flags & SIHMPC_F_UPDATE_VIN) && io.vinant && (last_step || !advance))
                    for (int k = lane; k < P.n_vin; k += 32) {
                        const double sn = ew[R + (P.vin_row[k] - P.o_si)];
                        io.vinant[s * P.n_vin + k] = P.vin_q[k] * sn * sn;
                    }
            }
            if (!advance) continue;                       // uniform: n_steps == 1
            // plant x+ = [A|B][x ; du]  (du = first nu entries of dU, which follow x in the buffer)
            if (shared_plant) {
                for (int i = lane; i < nx; i += 32) svw[i] = tbw[i];
            } else {
                for (int i = lane; i < nx; i += 32) {
                    double a = 0.0;
                    const int p1 = P.pl_rowptr[i + 1];
                    for (int pp = P.pl_rowptr[i]; pp < p1; ++pp) a = fma(pv[pp], svw[P.pl_col[pp]], a);
                    xtw[i] = a;
                }
                __syncwarp();
                for (int i = lane; i < nx; i += 32) svw[i] = xtw[i];
            }
            if (kabs == dstep) {
                __syncwarp();
                for (int i = lane; i < ny; i += 32) svw[i] += io.dist[sl * ny + i];
            }
            {
                double d2 = 0.0;
                for (int j = lane; j < nu; j += 32) { const double d = zw[j]; uw[j] += d; d2 = fma(d, d, d2); }
                d2 = warp_sum(d2);
                if (lane == 0) {
                    stw[ny] += d2;
                    stw[ny + 1] = fmax(stw[ny + 1], viol_abs);
                    stw[ny + 2] = Jval;
                }
            }
            for (int i = lane; i < ny; i += 32) {
                const double a = ew[P.o_sy + i], b = ew[R + i];
                sqw[Ny + i] = a * a; sqw[Ny + ny + i] = b * b;
            }
            for (int j = lane; j < nv; j += 32) sqw[nsq + j] = zw[j] * zw[j];
            __syncthreads();                              // [F]
            // ---------------- P6: measured output (statistics), Y' of the shifted plan -> squares
            for (int T = w; T < C.n_y; T += TL_WARPS) {
                double y0, y1;
                tile_row(C, C.my, T, sv + q * C.S_sv + g, lane, y0, y1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = 8 * T + 2 * g + h;
                    if (i < ny) {
                        const double yv_ = h ? y1 : y0;
                        const double err = yv_ - sv[q * C.S_sv + o_ysp + i];
                        scb[q * C.S_sc + 2 + nu + i] += err * err;
                        if (io.traj_y && scb[q * C.S_sc + 1] != 0.0) io.traj_y[((long long)s_stp[q] * batch + s_sid[q]) * ny + i] = yv_;
                    }
                }
            }
            if (io.flags & (SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1)) {
                for (int T = (w + 3) & 7; T < C.n_yp; T += TL_WARPS) {
                    double y0, y1;
                    tile_row(C, C.myp, T, sv + q * C.S_sv + g, lane, y0, y1);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = 8 * T + 2 * g + h;
                        if (r < Ny) {
                            const int i = r / P.N, k = r - i * P.N;      // output-major rows (tile_configure)
                            const double *eq = eb + q * C.S_r;
                            const double v_ = (h ? y1 : y0) - sv[q * C.S_sv + o_ysp + i] - eq[P.o_sy + i] - P.tau[k] * eq[R + i];
                            sq[q * C.S_q + r] = v_ * v_;
                        }
                    }
                }
                __syncthreads();                          // [G]
                // ---------------- P7: sub-objective values at the new state
                for (int T = w; T < C.n_m; T += TL_WARPS) {
                    double v0, v1;
                    tile_row(C, C.mvn, T, sq + q * C.S_q + g, lane, v0, v1);
                    *reinterpret_cast<double2 *>(vn + q * C.S_m + 8 * T + 2 * g) = make_double2(v0, v1);
                }
                __syncthreads();                          // [H]
                // ---------------- P8 (warp per scenario): satisficing weights
                if (io.flags & SIHMPC_F_WEIGHTS_SAT2) {
                    double smax = -SIH_INF;
                    for (int i = lane; i < m; i += 32) smax = fmax(smax, vnw[i] / P.gamma[i]);
                    smax = warp_max(smax);
                    for (int i = lane; i < m; i += 32) {
                        const double g0 = P.gamma[i];
                        const double g1 = fmax(g0, g0 * smax + 1e-6);
                        pw[1 + i] = io.alfa * pw[1 + i] + (1.0 - io.alfa) * (1.0 / (g1 - vnw[i]));
                    }
                } else {
                    for (int i = lane; i < m; i += 32) {
                        const double g0 = P.gamma[i];
                        pw[1 + i] = 1.0 / (g0 - fmin(fmax(vnw[i], 0.0), 0.99 * g0));
                    }
                }
            } else {
                __syncthreads();                          // the owners may not touch ysp before every y tile is done
            }
            if constexpr (RESUME) {
                if (alive && last_step) { write_state(); alive = false; if (lane == 0) scw[1] = 0.0; }
            }
        }  // steps

        // ---------------- write back
        if (!RESUME && alive && advance) write_state();
        __syncthreads();
    }
}

// ---- host side ----------------------------------------------------------------------------------
// The kernels are instantiated one tile-row count per translation unit (-DSIH_TL_ONLY_NT=k: only
// tile_set_smem_k / tile_launch_k), compiled in parallel by build.py; the unit without the define holds the host
// logic and dispatches over them.
#define SIH_TL_CAT2(a, b) a##b
#define SIH_TL_CAT(a, b) SIH_TL_CAT2(a, b)
#ifdef SIH_TL_ONLY_NT
cudaError_t SIH_TL_CAT(tile_set_smem_, SIH_TL_ONLY_NT)(size_t bytes) {
    cudaError_t e = sih_raise_smem_limit(reinterpret_cast<const void *>(tile_step_kernel<SIH_TL_ONLY_NT, 0>), bytes);
    return e != cudaSuccess ? e : sih_raise_smem_limit(reinterpret_cast<const void *>(tile_step_kernel<SIH_TL_ONLY_NT, 1>), bytes);
}
// mode 0: first pass (appends to `list`, counts in `counter`); mode 1: resume pass over `list` (n_list scenarios,
// known on the device only: blocks beyond it leave at once), failures counted in `counter`
cudaError_t SIH_TL_CAT(tile_launch_, SIH_TL_ONLY_NT)(const TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch,
                                                     int n_steps, int grid, int mode, int *list, const int *n_list,
                                                     int *counter, const unsigned long long *warm, cudaStream_t st) {
    if (mode == 0)
        tile_step_kernel<SIH_TL_ONLY_NT, 0><<<grid, TL_WARPS * 32, tp.smem_bytes, st>>>(P, tp.cfg, io, batch, n_steps, tp.start_step,
                                                                                         counter, tp.n_inkernel, list, nullptr, nullptr);
    else
        tile_step_kernel<SIH_TL_ONLY_NT, 1><<<grid, TL_WARPS * 32, tp.smem_bytes, st>>>(P, tp.cfg, io, batch, n_steps, tp.start_step,
                                                                                         counter, tp.n_inkernel, list, n_list, warm);
    return cudaGetLastError();
}
#else
#define SIH_TL_DECL(N_) \
    cudaError_t tile_set_smem_##N_(size_t bytes); \
    cudaError_t tile_launch_##N_(const TilePlan &, const DevPlan &, const sihmpc_io &, long long, int, int, int, int *, const int *, int *, \
                                 const unsigned long long *, cudaStream_t);
SIH_TL_DECL(1) SIH_TL_DECL(2) SIH_TL_DECL(3) SIH_TL_DECL(4) SIH_TL_DECL(5) SIH_TL_DECL(6) SIH_TL_DECL(7) SIH_TL_DECL(8)
#undef SIH_TL_DECL

struct TileBuilder {
    std::vector<int> ptr, kof;
    std::vector<double> B;
    int n_stored = 0;
    // one output tile: fill(t, tile[32]) writes k-tile t in B-fragment order and returns true if non-zero.  The
    // non-zero k-tiles are stored in order with their A offsets, the list padded to a multiple of four with zero
    // tiles (offset 0: the scenario's own first entries times zero).
    void row(int KT, const std::function<bool(int, double *)> &fill) {
        ptr.push_back(n_stored);
        std::vector<double> tile(32);
        int cnt = 0;
        for (int t = 0; t < KT; ++t) {
            if (!fill(t, tile.data())) continue;
            kof.push_back(4 * t);
            B.insert(B.end(), tile.begin(), tile.end());
            ++cnt;
        }
        for (; cnt % 4; ++cnt) { kof.push_back(0); B.insert(B.end(), 32, 0.0); }
        n_stored += cnt;
    }
    void close() { ptr.push_back(n_stored); }
    // a matrix with nT output tiles; M(r,c) over rows [0, 8 nT) and k-space [0, nc); returns the ptr offset
    int add(int nT, int nc, const std::function<double(int, int)> &M) {
        const int off = (int)ptr.size();
        const int KT = (nc + 3) / 4;
        for (int T = 0; T < nT; ++T)
            row(KT, [&](int t, double *tile) {
                bool nz = false;
                for (int lane = 0; lane < 32; ++lane) {
                    const int n = lane >> 2, k = lane & 3, c = 4 * t + k;
                    const double v = (c < nc) ? M(8 * T + n, c) : 0.0;
                    tile[lane] = v;
                    if (v != 0.0) nz = true;
                }
                return nz; });
        close();
        return off;
    }
};

static cudaError_t tile_set_smem(int NT, size_t bytes) {
#define SIH_TL_ATTR(N_) case N_: return tile_set_smem_##N_(bytes);
    switch (NT) {
        SIH_TL_ATTR(1) SIH_TL_ATTR(2) SIH_TL_ATTR(3) SIH_TL_ATTR(4)
        SIH_TL_ATTR(5) SIH_TL_ATTR(6) SIH_TL_ATTR(7) SIH_TL_ATTR(8)
        default: return cudaErrorInvalidValue;
    }
#undef SIH_TL_ATTR
}

inline int tl_stride(int n) { int s = (n + 3) & ~3; while (s % 16 != 4) s += 4; return s; }

cudaError_t tile_configure(const sihmpc_plan_desc *d, const DevPlan &P, TilePlan *tp) {
    memset(tp, 0, sizeof *tp);
    const int nx = d->nx, nu = d->nu, ny = d->ny, nv = P.nv, m = d->m, R = d->R, Ny = P.Ny, N = d->N;
    if (nv + 1 > 8 * TL_NTMAX || d->ncd > 0 || P.eval_term || P.nv != P.nvm) return cudaSuccess;
    TileCfg &C = tp->cfg;
    C.NT = (nv + 1 + 7) / 8; C.NP = (nv + 3) / 4; C.Ib = nv / 8; C.rb = nv % 8;
    const int ntiles = C.NT * (C.NT + 1) / 2;
    C.HS = ntiles * 64; while (C.HS % 16 != 8) C.HS += 2;
    const int RE = R + ny, K_sv = nx + nv + ny + 1, nsq = Ny + 2 * ny;
    C.n_r = (RE + 7) / 8; C.n_v = (nv + 7) / 8; C.n_f = C.NT; C.n_y = (ny + 7) / 8; C.n_yp = (Ny + 7) / 8; C.n_m = (m + 7) / 8;
    C.n_x = (nx + 7) / 8;
    int SR = 8 * C.n_r; if (SR < 8 * C.n_x) SR = 8 * C.n_x;
    C.S_sv = tl_stride(K_sv); C.S_r = tl_stride(SR); C.S_v = tl_stride(8 * C.NT); C.S_p = tl_stride(m + 1);
    C.S_q = tl_stride(nsq + nv); C.S_m = tl_stride(8 * C.n_m); C.S_sc = (2 + nu + ny + SIHMPC_NSTATS_EXTRA + 1) & ~1;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    C.o_H = take(TL_WARPS * C.HS); C.o_sv = take(TL_WARPS * C.S_sv); C.o_e = take(TL_WARPS * C.S_r);
    C.o_om = take(TL_WARPS * C.S_r); C.o_t = take(TL_WARPS * C.S_r); C.o_omdu = take(TL_WARPS * C.S_v);
    C.o_pes = take(TL_WARPS * C.S_p); C.o_sq = take(TL_WARPS * C.S_q); C.o_vn = take(TL_WARPS * C.S_m);
    C.o_z = take(TL_WARPS * 64); C.o_y = take(TL_WARPS * 64); C.o_sc = take(TL_WARPS * C.S_sc);
    C.total = o;
    tp->smem_bytes = (size_t)o * sizeof(double);
    if (tp->smem_bytes > 227 * 1024) return cudaSuccess;

    TileBuilder tb;
    const int o_v = nx, o_ysp = nx + nv, o_one = nx + nv + ny;
    // residual rows of the prediction in OUTPUT-major order inside this kernel (row i*N + k <- plan row k*ny + i): an
    // 8-row tile then spans two outputs instead of eight, and a row only touches the modes of its own output
    auto yr = [&](int r) { return r < Ny ? (r % N) * ny + (r / N) : r; };
    // e = [Px | Py | e0][x ; ysp ; 1], rows R.. = Mi1x x
    C.me = tb.add(C.n_r, K_sv, [&](int r, int c) -> double {
        if (r >= RE) return 0.0;
        if (c < nx) return r < R ? d->Px[(size_t)yr(r) * nx + c] : d->Mi1x[(size_t)(r - R) * nx + c];
        if (c >= o_ysp && c < o_one) return r < R ? d->Py[(size_t)yr(r) * ny + (c - o_ysp)] : 0.0;
        if (c == o_one) return r < R ? d->e0[yr(r)] : 0.0;
        return 0.0; });
    // omega = [w0row | Wrow][1 ; pesos]
    C.mw = tb.add(C.n_r, m + 1, [&](int r, int c) -> double {
        if (r >= R) return 0.0;
        return c == 0 ? d->w0row[yr(r)] : d->Wrow[(size_t)yr(r) * m + (c - 1)]; });
    // omega_du = Wdu pesos
    C.mdu = tb.add(C.n_v, m + 1, [&](int r, int c) -> double {
        return (r < nv && c >= 1) ? d->Wdu[(size_t)r * m + (c - 1)] : 0.0; });
    // b = -2 E^T t
    C.mf = tb.add(C.n_f, R, [&](int r, int c) -> double {
        return r < nv ? -2.0 * d->E[(size_t)yr(c) * nv + r] : 0.0; });
    // Hessian tiles (lower), k-space [1 ; pesos]: Ht at k = 0, HJ_a at k = 1 + hj_index[a]  (dense k-tiles)
    std::vector<double> BH;
    std::vector<int> hrows;
    {
        int kmax = 0;
        for (int a = 0; a < d->mh; ++a) if (1 + d->hj_index[a] > kmax) kmax = 1 + d->hj_index[a];
        C.KH = (kmax + 1 + 3) / 4;
        if (C.KH > TL_KHMAX) return cudaSuccess;
        C.ntile = ntiles;
        BH.assign((size_t)ntiles * 8 * C.KH * 32, 0.0);
        hrows.assign(ntiles, 0);
        for (int I = 0; I < C.NT; ++I)
            for (int J = 0; J <= I; ++J) {
                const int tid = tl_tid(I, J);
                int nr = nv - 8 * I; if (nr < 0) nr = 0; if (nr > 8) nr = 8;
                hrows[tid] = nr;
                for (int r = 0; r < nr; ++r)
                    for (int t = 0; t < C.KH; ++t)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int n = lane >> 2, k = lane & 3;
                            const int row = 8 * I + r, col = 8 * J + n, kk = 4 * t + k;
                            double v = 0.0;
                            if (col <= row && col < nv && kk <= m) {
                                if (kk == 0) v = d->Ht[(size_t)row * nv + col];
                                else
                                    for (int a = 0; a < d->mh; ++a)
                                        if (d->hj_index[a] == kk - 1) v += d->HJ[((size_t)a * nv + row) * nv + col];
                            }
                            BH[(((size_t)tid * 8 + r) * C.KH + t) * 32 + lane] = v;
                        }
            }
    }
    // Hessian blocks of the bound space (tile_active_set.inc): S^T X_k S with S = difference along the horizon per
    // input (cumulative moves) or I (move box); X_0 = Ht, X_{1+i} = sum_{a: hj_index[a]=i} HJ_a + 2 diag(Wdu[:, i])
    std::vector<double> BHc;
    C.cspace = (P.has_cum && !P.has_box) ? 1 : ((P.has_box && !P.has_cum) ? 0 : -1);
    C.KHc = (m + 1 + 3) / 4;
    if (C.KHc > TL_KHMAX || getenv("SIHMPC_NO_INKERNEL_AS")) C.cspace = -1;
    if (C.cspace >= 0) {
        const bool cum = (C.cspace == 1);
        std::vector<double> Xc((size_t)(m + 1) * nv * nv, 0.0), X((size_t)nv * nv), XS((size_t)nv * nv);
        for (int k = 0; k <= m; ++k) {
            for (int a = 0; a < nv; ++a)
                for (int b = 0; b < nv; ++b) {
                    double x = 0.0;
                    if (k == 0) x = d->Ht[(size_t)a * nv + b];
                    else {
                        for (int h = 0; h < d->mh; ++h)
                            if (d->hj_index[h] == k - 1) x += d->HJ[((size_t)h * nv + a) * nv + b];
                        if (a == b) x += 2.0 * d->Wdu[(size_t)a * m + (k - 1)];
                    }
                    X[(size_t)a * nv + b] = x;
                }
            // (X S)[a][b] = X[a][b] - X[a][b+nu] ; (S^T Y)[a][b] = Y[a][b] - Y[a+nu][b]
            for (int a = 0; a < nv; ++a)
                for (int b = 0; b < nv; ++b)
                    XS[(size_t)a * nv + b] = X[(size_t)a * nv + b] - ((cum && b + nu < nv) ? X[(size_t)a * nv + b + nu] : 0.0);
            for (int a = 0; a < nv; ++a)
                for (int b = 0; b < nv; ++b)
                    Xc[((size_t)k * nv + a) * nv + b] = XS[(size_t)a * nv + b] - ((cum && a + nu < nv) ? XS[(size_t)(a + nu) * nv + b] : 0.0);
        }
        BHc.assign((size_t)ntiles * 8 * C.KHc * 32, 0.0);
        for (int I = 0; I < C.NT; ++I)
            for (int J = 0; J <= I; ++J) {
                const int tid = tl_tid(I, J);
                for (int r = 0; r < hrows[tid]; ++r)
                    for (int t = 0; t < C.KHc; ++t)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int n = lane >> 2, k = lane & 3;
                            const int row = 8 * I + r, col = 8 * J + n, kk = 4 * t + k;
                            if (col <= row && col < nv && kk <= m)
                                BHc[(((size_t)tid * 8 + r) * C.KHc + t) * 32 + lane] = Xc[((size_t)kk * nv + row) * nv + col];
                        }
            }
    }
    // r += [E ; Mi1u] dU
    C.mr = tb.add(C.n_r, K_sv, [&](int r, int c) -> double {
        if (r >= RE || c < o_v || c >= o_v + nv) return 0.0;
        return r < R ? d->E[(size_t)yr(r) * nv + (c - o_v)] : d->Mi1u[(size_t)(r - R) * nv + (c - o_v)]; });
    // y = [C | D][x+ ; du]
    C.my = tb.add(C.n_y, K_sv, [&](int r, int c) -> double {
        if (r >= ny) return 0.0;
        if (c < nx) return d->C[(size_t)r * nx + c];
        if (c < nx + nu) return d->D[(size_t)r * nu + (c - nx)];
        return 0.0; });
    // Y' = Myx x+ + Myu shift(dU)
    C.myp = tb.add(C.n_yp, K_sv, [&](int r, int c) -> double {
        if (r >= Ny) return 0.0;
        if (c < nx) return d->Myx[(size_t)yr(r) * nx + c];
        const int v = c - o_v;
        return (v >= nu && v < nv) ? d->Myu[(size_t)yr(r) * nv + (v - nu)] : 0.0; });
    // V_i = sum_r Wrow[r][i] sq_r + sum_v Wdu[v-nu][i] dU_v^2
    C.mvn = tb.add(C.n_m, nsq + nv, [&](int i, int c) -> double {
        if (i >= m) return 0.0;
        if (c < nsq) return d->Wrow[(size_t)yr(c) * m + i];
        const int v = c - nsq;
        return v >= nu ? d->Wdu[(size_t)(v - nu) * m + i] : 0.0; });
    // x+ = [A | B][x ; du]  (shared plant; per-scenario plant values use the CSR path)
    C.mpl = tb.add(C.n_x, K_sv, [&](int r, int c) -> double {
        if (r >= nx || c >= nx + nu) return 0.0;
        for (int pp = d->pl_rowptr[r]; pp < d->pl_rowptr[r + 1]; ++pp)
            if (d->pl_col[pp] == c) return d->pl_val[pp];
        return 0.0; });

    const size_t nk = tb.kof.size(), np = tb.ptr.size(), nh = hrows.size();      // kof first: 16-byte aligned int4 loads
    std::vector<int> ib(nk + np + nh);
    memcpy(ib.data(), tb.kof.data(), nk * sizeof(int));
    memcpy(ib.data() + nk, tb.ptr.data(), np * sizeof(int));
    memcpy(ib.data() + nk + np, hrows.data(), nh * sizeof(int));
    cudaError_t e = cudaMalloc(&tp->ibuf, ib.size() * sizeof(int));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(tp->ibuf, ib.data(), ib.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    const size_t nB = (tb.B.size() + 31) & ~(size_t)31;
    const size_t nBH = (BH.size() + 31) & ~(size_t)31;
    e = cudaMalloc(&tp->dbuf, (nB + nBH + BHc.size() + 32) * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(tp->dbuf, tb.B.data(), tb.B.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(tp->dbuf + nB, BH.data(), BH.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    if (!BHc.empty()) {
        e = cudaMemcpy(tp->dbuf + nB + nBH, BHc.data(), BHc.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
    }
    C.kof = tp->ibuf; C.ptr = tp->ibuf + nk; C.hrows = tp->ibuf + nk + np; C.B = tp->dbuf; C.BH = tp->dbuf + nB;
    C.BHc = tp->dbuf + nB + nBH;
    tp->n_tiles = tb.n_stored;
    if (getenv("SIHMPC_TILE_DEBUG")) {          // work balance of the product phases: groups of four tiles per warp and matrix
        const int *pt = tb.ptr.data();
        auto grp = [&](int poff, int T) { return (pt[poff + T + 1] - pt[poff + T]) / 4; };
        struct { const char *name; int off, n, rot; } ms[] = {{"me", C.me, C.n_r, 0}, {"mw", C.mw, C.n_r, 0}, {"mdu", C.mdu, C.n_v, 5},
            {"mf", C.mf, C.n_f, 0}, {"mr", C.mr, C.n_r, 0}, {"mpl", C.mpl, C.n_x, 2}, {"my", C.my, C.n_y, 0}, {"myp", C.myp, C.n_yp, 3}, {"mvn", C.mvn, C.n_m, 0}};
        for (auto &mm : ms) {
            fprintf(stderr, "[tile] %-4s tiles %3d groups/warp:", mm.name, mm.n);
            int tot = 0, mx = 0;
            for (int w = 0; w < TL_WARPS; ++w) {
                int gsum = 0;
                for (int T = (w + mm.rot) & 7; T < mm.n; T += TL_WARPS) gsum += grp(mm.off, T);
                fprintf(stderr, " %3d", gsum); tot += gsum; if (gsum > mx) mx = gsum;
            }
            fprintf(stderr, "  total %d max %d (balance %.2f)\n", tot, mx, tot / (8.0 * (mx ? mx : 1)));
        }
        fprintf(stderr, "[tile] KH %d ntile %d NT %d n_stored %d\n", C.KH, C.ntile, C.NT, tb.n_stored);
    }
    e = tile_set_smem(C.NT, tp->smem_bytes);
    if (e != cudaSuccess) return e;
    tp->ok = 1;
    return cudaSuccess;
}

// the hand-over block lives in the caller's workspace (capi.cu: bind_handover): nothing is allocated at step time
void tile_bind(TilePlan *tp, int *ctr, unsigned long long *warm, long long batch) {
    tp->ctr = ctr; tp->warm_mask = warm;
    tp->n_handover = ctr; tp->n_fail = ctr + 1; tp->n_inkernel = ctr + 2;
    tp->start_step = ctr + TL_CTR; tp->list = ctr + TL_CTR + batch; tp->list2 = ctr + TL_CTR + 2 * batch;
    tp->capacity = batch;
}

// counters and start_step back to "nothing handed over" (every counter counts up from the baseline 0x7f7f7f7f)
cudaError_t tile_reset(const TilePlan &tp, long long batch, cudaStream_t st) {
    return cudaMemsetAsync(tp.ctr, 0x7f, (size_t)(TL_CTR + batch) * sizeof(int), st);
}

static int tile_grid(const TilePlan &tp, long long batch, int sm_count) {
    const long long n_oct = (batch + TL_WARPS - 1) / TL_WARPS;
    int per_sm = (int)((227 * 1024) / (tp.smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const long long cap = (long long)sm_count * per_sm;
    return (int)(n_oct < cap ? n_oct : cap);
}

static cudaError_t tile_dispatch(const TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch, int n_steps,
                                 int grid, int mode, int *list, const int *n_list, int *counter,
                                 const unsigned long long *warm, cudaStream_t st) {
#define SIH_TL_LAUNCH(N_) case N_: return tile_launch_##N_(tp, P, io, batch, n_steps, grid, mode, list, n_list, counter, warm, st);
    switch (tp.cfg.NT) {
        SIH_TL_LAUNCH(1) SIH_TL_LAUNCH(2) SIH_TL_LAUNCH(3) SIH_TL_LAUNCH(4)
        SIH_TL_LAUNCH(5) SIH_TL_LAUNCH(6) SIH_TL_LAUNCH(7) SIH_TL_LAUNCH(8)
        default: return cudaErrorInvalidValue;
    }
#undef SIH_TL_LAUNCH
}

// first pass over the whole batch, then (bounds present, in-kernel active set available, enough work to pay for one
// more launch) the resume pass over the scenarios it listed; what that cannot resolve is counted in n_fail with
// start_step set
cudaError_t tile_launch(TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch,
                               int n_steps, int sm_count, cudaStream_t st, bool allow_resume) {
    const int grid = tile_grid(tp, batch, sm_count);
    tp.last_resumed = 0;
    cudaError_t e = tile_dispatch(tp, P, io, batch, n_steps, grid, 0, tp.list, nullptr, tp.n_handover, nullptr, st);
    if (e != cudaSuccess || !allow_resume || !tile_resumes(tp, batch, n_steps) || !(P.has_box || P.has_cum)) return e;
    tp.last_resumed = 1;
    return tile_dispatch(tp, P, io, batch, n_steps, grid, 1, tp.list, tp.n_handover, tp.n_fail, nullptr, st);
}

// resume pass number `round` (0..TL_BOUNCES-1) over the scenarios the general kernel carried through one hard step and
// sent back with their active set (list2 / tile_bounce_counter(round)); failures counted in tile_fail_counter(round)
cudaError_t tile_launch_bounced(const TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch,
                                int n_steps, int sm_count, int round, cudaStream_t st) {
    return tile_dispatch(tp, P, io, batch, n_steps, tile_grid(tp, batch, sm_count), 1, tp.list2, tile_bounce_counter(tp, round),
                         tile_fail_counter(tp, round), tp.warm_mask, st);
}

void tile_destroy(TilePlan *tp) {
    if (tp->ibuf) cudaFree(tp->ibuf);
    if (tp->dbuf) cudaFree(tp->dbuf);
    memset(tp, 0, sizeof *tp);
}
#endif  // SIH_TL_ONLY_NT

}  // namespace sih
