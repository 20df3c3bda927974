// tile_kernel.cuh -- fast path v4 for MID-SIZE controllers (N*nu + 1 <= 64, any nx / R / m):
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
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <functional>
#include <string.h>
#include <stdlib.h>
#include "devplan.h"
#include "step_kernel.cuh"
#include "mma_kernel.cuh"

namespace sih {

constexpr int TL_WARPS = 8;          // == scenarios per CTA == M of the MMA
constexpr int TL_NTMAX = 8;          // tile rows of the augmented matrix: nv + 1 <= 64

struct TileCfg {
    const int *ptr;                  // row pointers of every tile matrix (absolute tile indices)
    const int *kt;                   // k-tile of every stored tile
    const int *hmap;                 // Hessian product: destination (tile*64 + row*8) of output tile T
    const double *B;                 // stored tiles, 32 doubles each, B-fragment order
    int me, mw, mdu, mf, mh, mr, my, myp, mvn;       // offsets of the row-pointer arrays in ptr[]
    int n_r, n_v, n_f, n_h, n_y, n_yp, n_m;          // output tile counts
    int NT, NP, Ib, rb, HS;
    int S_sv, S_r, S_v, S_p, S_q, S_m, S_sc;         // per-scenario strides (doubles; == 4 mod 16 where read as A)
    int o_H, o_sv, o_e, o_om, o_t, o_omdu, o_pes, o_sq, o_vn, o_z, o_y, o_sc;
    int total;                       // doubles of shared memory per CTA
};

__host__ __device__ __forceinline__ int tl_tid(int I, int J) { return (I * (I + 1)) / 2 + J; }

// accumulators of output tile T of one tile-CSR matrix; in_q = in + q*S + g (A-fragment base of this lane)
__device__ __forceinline__ void tile_row(const TileCfg &C, const int poff, const int T, const double *in_q,
                                         const int lane, double &r0, double &r1) {
    const int i0 = C.ptr[poff + T], i1 = C.ptr[poff + T + 1];
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    for (int base = i0; base < i1; base += 32) {
        const int n = (i1 - base < 32) ? (i1 - base) : 32;
        const int ktv = (lane < n) ? __ldg(C.kt + base + lane) : 0;
        const double *Bp = C.B + (size_t)base * 32 + lane;
        int j = 0;
        for (; j + 4 <= n; j += 4) {
            const int k0 = __shfl_sync(SIH_FULL, ktv, j), k1 = __shfl_sync(SIH_FULL, ktv, j + 1);
            const int k2 = __shfl_sync(SIH_FULL, ktv, j + 2), k3 = __shfl_sync(SIH_FULL, ktv, j + 3);
            const double b0 = __ldg(Bp + (j) * 32), b1 = __ldg(Bp + (j + 1) * 32);
            const double b2 = __ldg(Bp + (j + 2) * 32), b3 = __ldg(Bp + (j + 3) * 32);
            const double a0 = in_q[4 * k0], a1 = in_q[4 * k1], a2 = in_q[4 * k2], a3 = in_q[4 * k3];
            dmma884(c0, c1, a0, b0);
            dmma884(d0, d1, a1, b1);
            dmma884(c0, c1, a2, b2);
            dmma884(d0, d1, a3, b3);
        }
        for (; j < n; ++j) {
            const int k0 = __shfl_sync(SIH_FULL, ktv, j);
            dmma884(c0, c1, in_q[4 * k0], __ldg(Bp + j * 32));
        }
    }
    r0 = c0 + d0; r1 = c1 + d1;
}

__device__ __forceinline__ double tl_rsqrt(const double d) { return fast_rsqrt(d); }

// Right-looking tiled Cholesky of the augmented matrix [H ; b^T] of ONE scenario by ONE warp.
// A: packed lower 8x8 tiles (row-major 64 doubles each) in shared memory.  On exit the strictly
// lower part holds L, every 4x4 pivot block holds W = L11^-1 (lower), yv[0..nv) = L^-1 b.
template <int NTMAX>
__device__ int tile_factor(double *A, const int NT, const int NP, const int nv, const int Ib, const int rb,
                           const int lane, double *yv) {
    const int q = lane >> 2, g = lane & 3;
    int notpd = 0;
    for (int p = 0; p < NP; ++p) {
        const int J = p >> 1, c0 = (p & 1) * 4;
        double *D = A + tl_tid(J, J) * 64 + c0 * 8 + c0;          // pivot block, element (r,c) at D[r*8+c]
        const int nreal = (nv - 4 * p < 4) ? (nv - 4 * p) : 4;    // columns >= nreal of the block: identity
        double a00 = D[0], a10 = D[8], a11 = D[9], a20 = D[16], a21 = D[17], a22 = D[18];
        double a30 = D[24], a31 = D[25], a32 = D[26], a33 = D[27];
        if (!(a00 > 0.0)) { notpd = 1; a00 = 1.0; }
        const double i00 = tl_rsqrt(a00);
        const double l10 = a10 * i00, l20 = a20 * i00, l30 = a30 * i00;
        double d1 = fma(-l10, l10, a11);
        if (nreal < 2) d1 = 1.0; else if (!(d1 > 0.0)) { notpd = 1; d1 = 1.0; }
        const double i11 = tl_rsqrt(d1);
        const double l21 = (nreal < 2) ? 0.0 : fma(-l20, l10, a21) * i11;
        const double l31 = (nreal < 2) ? 0.0 : fma(-l30, l10, a31) * i11;
        double d2 = fma(-l21, l21, fma(-l20, l20, a22));
        if (nreal < 3) d2 = 1.0; else if (!(d2 > 0.0)) { notpd = 1; d2 = 1.0; }
        const double i22 = tl_rsqrt(d2);
        const double l32 = (nreal < 3) ? 0.0 : fma(-l31, l21, fma(-l30, l20, a32)) * i22;
        double d3 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
        if (nreal < 4) d3 = 1.0; else if (!(d3 > 0.0)) { notpd = 1; d3 = 1.0; }
        const double i33 = tl_rsqrt(d3);
        // W = L11^-1
        const double w10 = -l10 * i00 * i11;
        const double w21 = -l21 * i11 * i22;
        const double w32 = -l32 * i22 * i33;
        const double w20 = -fma(l21, w10, l20 * i00) * i22;
        const double w31 = -fma(l32, w21, l31 * i11) * i33;
        const double w30 = -fma(l32, w20, fma(l31, w10, l30 * i00)) * i33;
        // element (r,c) of W for this lane's (q,g) resp. (lane>>2, lane&3)
        double wqg;
        {
            const double r0 = (g == 0) ? i00 : 0.0;
            const double r1 = (g == 0) ? w10 : ((g == 1) ? i11 : 0.0);
            const double r2 = (g == 0) ? w20 : ((g == 1) ? w21 : ((g == 2) ? i22 : 0.0));
            const double r3 = (g == 0) ? w30 : ((g == 1) ? w31 : ((g == 2) ? w32 : i33));
            wqg = (q == 0) ? r0 : ((q == 1) ? r1 : ((q == 2) ? r2 : ((q == 3) ? r3 : 0.0)));
            if (q >= nreal) wqg = 0.0;
        }
        // panel solve  L21 = A21 W^T  for every row tile (rows in or above the pivot block are skipped)
        for (int I = J; I < NT; ++I) {
            double *tile = A + tl_tid(I, J) * 64;
            double a = tile[q * 8 + c0 + g];
            const bool skip = (I == J) && (q < c0 + 4);
            if (skip) a = 0.0;
            double o0 = 0.0, o1 = 0.0;
            dmma884(o0, o1, a, wqg);
            if (g < 2 && !skip) *reinterpret_cast<double2 *>(tile + q * 8 + c0 + 2 * g) = make_double2(o0, o1);
        }
        // y of these columns when the right-hand-side row sits inside the pivot block
        if (nreal < 4) {
            if (lane < nreal) {
                const double y0 = (nreal == 1) ? l10 : ((nreal == 2) ? l20 : l30);
                const double y1 = (nreal == 2) ? l21 : l31;
                yv[4 * p + lane] = (lane == 0) ? y0 : ((lane == 1) ? y1 : l32);
            }
        }
        if (lane < 16 && g <= q && q < nreal) D[q * 8 + g] = wqg;     // (q,g) = (lane>>2, lane&3) for lane < 16
        __syncwarp();
        if (nreal == 4 && lane < 4) yv[4 * p + lane] = A[tl_tid(Ib, J) * 64 + rb * 8 + c0 + lane];
        // trailing update: tile(I,K) -= L21_I L21_K^T  (A fragment of I, B fragment of K: the same layout)
        double pa[NTMAX];
#pragma unroll
        for (int I = 0; I < NTMAX; ++I) {
            double v = 0.0;
            if (I >= J && I < NT) {
                v = A[tl_tid(I, J) * 64 + q * 8 + c0 + g];
                if (I == J && q < c0 + 4) v = 0.0;
            }
            pa[I] = v;
        }
        const int K0 = J + (c0 ? 1 : 0);
#pragma unroll
        for (int K = 0; K < NTMAX; ++K) {
            if (K < K0 || K >= NT) continue;
            const double bK = pa[K];
#pragma unroll
            for (int I = K; I < NTMAX; ++I) {
                if (I < NT) {
                    double2 *cp = reinterpret_cast<double2 *>(A + tl_tid(I, K) * 64 + q * 8 + 2 * g);
                    double2 c = *cp;
                    dmma884(c.x, c.y, -pa[I], bK);
                    *cp = c;
                }
            }
        }
        __syncwarp();
    }
    return notpd;
}

// L^T z = y with the stored W blocks: zv[0..nv)
__device__ void tile_backsolve(const double *A, const int NP, const int nv, const int lane, const double *yv,
                               double *zv) {
    const int rq = lane >> 2, cg = lane & 3;
    for (int p = NP - 1; p >= 0; --p) {
        const int J = p >> 1, c0 = (p & 1) * 4;
        const int nreal = (nv - 4 * p < 4) ? (nv - 4 * p) : 4;
        double acc = 0.0;
        for (int i = 4 * p + 4 + rq; i < nv; i += 8)
            acc = fma(A[tl_tid(i >> 3, J) * 64 + (i & 7) * 8 + c0 + cg], zv[i], acc);
        acc += __shfl_xor_sync(SIH_FULL, acc, 4);
        acc += __shfl_xor_sync(SIH_FULL, acc, 8);
        acc += __shfl_xor_sync(SIH_FULL, acc, 16);
        const double t = (cg < nreal) ? yv[4 * p + cg] - acc : 0.0;
        const double *D = A + tl_tid(J, J) * 64 + c0 * 8 + c0;
        double z = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double tr = __shfl_sync(SIH_FULL, t, r, 4);
            if (r >= cg && r < nreal) z = fma(D[r * 8 + cg], tr, z);
        }
        if (rq == 0 && cg < nreal) zv[4 * p + cg] = z;
        __syncwarp();
    }
}

template <int NTMAX>
__global__ void __launch_bounds__(TL_WARPS * 32, 1)
tile_step_kernel(const DevPlan P, const TileCfg C, const sihmpc_io io, const long long batch, const int n_steps,
                 int *__restrict__ start_step, int *__restrict__ n_handover) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, q = lane >> 2, g = lane & 3;
    const int nx = P.nx, nu = P.nu, ny = P.ny, nv = P.nv, m = P.m, R = P.R, Ny = P.Ny;
    const int nstat = ny + SIHMPC_NSTATS_EXTRA;
    const bool advance = (io.flags & SIHMPC_F_ADVANCE) != 0;
    for (int i = threadIdx.x; i < C.total; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    double *Hs = sm + C.o_H, *sv = sm + C.o_sv, *eb = sm + C.o_e, *omb = sm + C.o_om, *tb = sm + C.o_t;
    double *omdu = sm + C.o_omdu, *pes = sm + C.o_pes, *sq = sm + C.o_sq, *vn = sm + C.o_vn;
    double *scb = sm + C.o_sc;
    // views of the scenario this warp owns in the per-scenario phases
    double *Hw = Hs + (size_t)w * C.HS, *svw = sv + w * C.S_sv, *ew = eb + w * C.S_r, *omw = omb + w * C.S_r;
    double *omduw = omdu + w * C.S_v, *pw = pes + w * C.S_p, *sqw = sq + w * C.S_q, *vnw = vn + w * C.S_m;
    double *zw = sm + C.o_z + w * 64, *yw = sm + C.o_y + w * 64, *scw = scb + w * C.S_sc;
    double *uw = scw + 2, *stw = scw + 2 + nu;
    double *xtw = omw;                                // plant scratch: omega is dead after the cost
    const int o_ysp = nx + nv, o_one = nx + nv + ny;
    const int nsq = Ny + 2 * ny;

    const long long n_oct = (batch + TL_WARPS - 1) / TL_WARPS;
    for (long long oct = blockIdx.x; oct < n_oct; oct += gridDim.x) {
        const long long s = oct * TL_WARPS + w;
        const bool valid = s < batch;
        const long long sl = valid ? s : batch - 1;
        bool alive = valid;
        // ---------------- load the scenario
        for (int i = lane; i < nx; i += 32) svw[i] = io.x[sl * nx + i];
        for (int i = lane; i < nv; i += 32) svw[nx + i] = 0.0;
        for (int i = lane; i < m; i += 32) pw[i] = io.pesos[sl * m + i];
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
        if (lane == 0) { svw[o_one] = 1.0; pw[m] = 1.0; scw[1] = alive ? 1.0 : 0.0; }
        const double *pv = io.plant_val ? io.plant_val + (size_t)sl * P.pl_nnz : P.pl_val;

        for (int step = 0; step < n_steps; ++step) {
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
            for (int T = w; T < C.n_h; T += TL_WARPS) {
                double h0, h1;
                tile_row(C, C.mh, T, pes + q * C.S_p + g, lane, h0, h1);
                *reinterpret_cast<double2 *>(Hs + (size_t)q * C.HS + __ldg(C.hmap + T) + 2 * g) = make_double2(h0, h1);
            }
            __syncthreads();                              // [B]
            // ---------------- P2: right-hand side b = -2 E^T t into row nv of the augmented matrix
            for (int T = w; T < C.n_f; T += TL_WARPS) {
                double f0, f1;
                tile_row(C, C.mf, T, tb + q * C.S_r + g, lane, f0, f1);
                *reinterpret_cast<double2 *>(Hs + (size_t)q * C.HS + tl_tid(C.Ib, T) * 64 + C.rb * 8 + 2 * g) = make_double2(f0, f1);
            }
            __syncthreads();                              // [C]
            // ---------------- P3 (warp per scenario): factorise, solve, check the boxes
            const int notpd = tile_factor<NTMAX>(Hw, C.NT, C.NP, nv, C.Ib, C.rb, lane, yw);
            tile_backsolve(Hw, C.NP, nv, lane, yw, zw);
            for (int i = lane; i < nv; i += 32) svw[nx + i] = zw[i];
            double viol = 0.0, viol_abs = 0.0;
            if (P.has_box)
                for (int i = lane; i < nv; i += 32) {
                    const int j = i % nu;
                    const double lo = P.dulb[j], hi = P.duub[j], val = zw[i];
                    if (lo > -SIH_INF) { const double a = lo - val; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(lo))); }
                    if (hi < SIH_INF) { const double a = val - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(hi))); }
                }
            if (P.has_cum)
                for (int i = lane; i < nv; i += 32) {
                    const int j = i % nu;
                    double val = 0.0;
                    for (int c = j; c <= i; c += nu) val += zw[c];
                    const double lo = P.ulb[j] - uw[j], hi = P.uub[j] - uw[j];
                    if (lo > -SIH_INF) { const double a = lo - val; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(lo))); }
                    if (hi < SIH_INF) { const double a = val - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(hi))); }
                }
            viol = warp_max(viol);
            viol_abs = warp_max(viol_abs);
            if (alive && (viol > 1e-10 || notpd)) {
                alive = false;                            // hand over to the general kernel from this step on
                if (lane == 0) { start_step[s] = step; atomicAdd(n_handover, 1); scw[1] = 0.0; }
                for (int i = lane; i < nx; i += 32) io.x[s * nx + i] = svw[i];
                for (int i = lane; i < nu; i += 32) io.u[s * nu + i] = uw[i];
                for (int i = lane; i < m; i += 32) io.pesos[s * m + i] = pw[i];
                if (io.stats) for (int i = lane; i < nstat; i += 32) io.stats[s * nstat + i] = stw[i];
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
                        if (io.n_active) io.n_active[s] = 0;
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
            for (int i = lane; i < nx; i += 32) {
                double a = 0.0;
                const int p1 = P.pl_rowptr[i + 1];
                for (int pp = P.pl_rowptr[i]; pp < p1; ++pp) a = fma(pv[pp], svw[P.pl_col[pp]], a);
                xtw[i] = a;
            }
            __syncwarp();
            for (int i = lane; i < nx; i += 32) svw[i] = xtw[i];
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
                        const long long sq_ = oct * TL_WARPS + q;
                        if (io.traj_y && scb[q * C.S_sc + 1] != 0.0) io.traj_y[((long long)step * batch + sq_) * ny + i] = yv_;
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
                            const int i = r % ny, k = r / ny;
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
                        pw[i] = io.alfa * pw[i] + (1.0 - io.alfa) * (1.0 / (g1 - vnw[i]));
                    }
                } else {
                    for (int i = lane; i < m; i += 32) {
                        const double g0 = P.gamma[i];
                        pw[i] = 1.0 / (g0 - fmin(fmax(vnw[i], 0.0), 0.99 * g0));
                    }
                }
            }
        }  // steps

        // ---------------- write back
        if (alive && advance) {
            for (int i = lane; i < nx; i += 32) io.x[s * nx + i] = svw[i];
            for (int i = lane; i < nu; i += 32) io.u[s * nu + i] = uw[i];
            for (int i = lane; i < m; i += 32) io.pesos[s * m + i] = pw[i];
            if (io.stats) for (int i = lane; i < nstat; i += 32) io.stats[s * nstat + i] = stw[i];
        }
        __syncthreads();
    }
}

// ---- host side ----------------------------------------------------------------------------------
struct TilePlan {
    int ok;                  // 1: the tile path can run this controller
    TileCfg cfg;
    int *ibuf;               // device: ptr | kt | hmap
    double *dbuf;            // device: stored tiles
    int *start_step;
    int *n_handover;
    long long capacity;
    size_t smem_bytes;
    int n_tiles;
};

struct TileBuilder {
    std::vector<int> ptr, kt, hmap;
    std::vector<double> B;
    // adds a matrix with nT output tiles; M(r,c) over rows [0, 8 nT) and k-space [0, nc); returns the ptr offset
    int add(int nT, int nc, const std::function<double(int, int)> &M) {
        const int off = (int)ptr.size();
        const int KT = (nc + 3) / 4;
        for (int T = 0; T < nT; ++T) {
            ptr.push_back((int)kt.size());
            for (int t = 0; t < KT; ++t) {
                double tile[32];
                bool nz = false;
                for (int lane = 0; lane < 32; ++lane) {
                    const int n = lane >> 2, k = lane & 3;
                    const int c = 4 * t + k;
                    const double v = (c < nc) ? M(8 * T + n, c) : 0.0;
                    tile[lane] = v;
                    if (v != 0.0) nz = true;
                }
                if (nz) { kt.push_back(t); B.insert(B.end(), tile, tile + 32); }
            }
        }
        ptr.push_back((int)kt.size());
        return off;
    }
};

inline int tl_stride(int n) { int s = (n + 3) & ~3; while (s % 16 != 4) s += 4; return s; }

inline cudaError_t tile_configure(const sihmpc_plan_desc *d, const DevPlan &P, TilePlan *tp) {
    memset(tp, 0, sizeof *tp);
    const int nx = d->nx, nu = d->nu, ny = d->ny, nv = P.nv, m = d->m, R = d->R, Ny = P.Ny, N = d->N;
    if (nv + 1 > 8 * TL_NTMAX || d->ncd > 0 || P.eval_term) return cudaSuccess;
    TileCfg &C = tp->cfg;
    C.NT = (nv + 1 + 7) / 8; C.NP = (nv + 3) / 4; C.Ib = nv / 8; C.rb = nv % 8;
    const int ntiles = C.NT * (C.NT + 1) / 2;
    C.HS = ntiles * 64; while (C.HS % 16 != 8) C.HS += 2;
    const int RE = R + ny, K_sv = nx + nv + ny + 1, nsq = Ny + 2 * ny;
    C.n_r = (RE + 7) / 8; C.n_v = (nv + 7) / 8; C.n_f = C.NT; C.n_y = (ny + 7) / 8; C.n_yp = (Ny + 7) / 8; C.n_m = (m + 7) / 8;
    int SR = 8 * C.n_r; if (SR < nx) SR = nx;
    C.S_sv = tl_stride(K_sv); C.S_r = tl_stride(SR); C.S_v = tl_stride(8 * C.n_v); C.S_p = tl_stride(m + 1);
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
    // e = [Px | Py | e0][x ; ysp ; 1], rows R.. = Mi1x x
    C.me = tb.add(C.n_r, K_sv, [&](int r, int c) -> double {
        if (r >= RE) return 0.0;
        if (c < nx) return r < R ? d->Px[(size_t)r * nx + c] : d->Mi1x[(size_t)(r - R) * nx + c];
        if (c >= o_ysp && c < o_one) return r < R ? d->Py[(size_t)r * ny + (c - o_ysp)] : 0.0;
        if (c == o_one) return r < R ? d->e0[r] : 0.0;
        return 0.0; });
    // omega = [Wrow | w0row][pesos ; 1]
    C.mw = tb.add(C.n_r, m + 1, [&](int r, int c) -> double {
        if (r >= R) return 0.0;
        return c < m ? d->Wrow[(size_t)r * m + c] : d->w0row[r]; });
    // omega_du = Wdu pesos
    C.mdu = tb.add(C.n_v, m + 1, [&](int r, int c) -> double {
        return (r < nv && c < m) ? d->Wdu[(size_t)r * m + c] : 0.0; });
    // b = -2 E^T t
    C.mf = tb.add(C.n_f, R, [&](int r, int c) -> double {
        return r < nv ? -2.0 * d->E[(size_t)c * nv + r] : 0.0; });
    // Hessian tiles (lower), k-space [pesos ; 1]
    {
        C.mh = (int)tb.ptr.size();
        const int KT = (m + 1 + 3) / 4;
        for (int I = 0; I < C.NT; ++I)
            for (int J = 0; J <= I; ++J)
                for (int r = 0; r < 8; ++r) {
                    const int row = 8 * I + r;
                    if (row >= nv) continue;
                    tb.hmap.push_back(tl_tid(I, J) * 64 + r * 8);
                    tb.ptr.push_back((int)tb.kt.size());
                    for (int t = 0; t < KT; ++t) {
                        double tile[32];
                        bool nz = false;
                        for (int lane = 0; lane < 32; ++lane) {
                            const int n = lane >> 2, k = lane & 3;
                            const int col = 8 * J + n, kk = 4 * t + k;
                            double v = 0.0;
                            if (col <= row && col < nv && kk <= m) {
                                if (kk == m) v = d->Ht[(size_t)row * nv + col];
                                else {
                                    for (int a = 0; a < d->mh; ++a)
                                        if (d->hj_index[a] == kk) v += d->HJ[((size_t)a * nv + row) * nv + col];
                                    if (col == row) v += 2.0 * d->Wdu[(size_t)row * m + kk];
                                }
                            }
                            tile[lane] = v;
                            if (v != 0.0) nz = true;
                        }
                        if (nz) { tb.kt.push_back(t); tb.B.insert(tb.B.end(), tile, tile + 32); }
                    }
                }
        tb.ptr.push_back((int)tb.kt.size());
        C.n_h = (int)tb.hmap.size();
    }
    // r += [E ; Mi1u] dU
    C.mr = tb.add(C.n_r, K_sv, [&](int r, int c) -> double {
        if (r >= RE || c < o_v || c >= o_v + nv) return 0.0;
        return r < R ? d->E[(size_t)r * nv + (c - o_v)] : d->Mi1u[(size_t)(r - R) * nv + (c - o_v)]; });
    // y = [C | D][x+ ; du]
    C.my = tb.add(C.n_y, K_sv, [&](int r, int c) -> double {
        if (r >= ny) return 0.0;
        if (c < nx) return d->C[(size_t)r * nx + c];
        if (c < nx + nu) return d->D[(size_t)r * nu + (c - nx)];
        return 0.0; });
    // Y' = Myx x+ + Myu shift(dU)
    C.myp = tb.add(C.n_yp, K_sv, [&](int r, int c) -> double {
        if (r >= Ny) return 0.0;
        if (c < nx) return d->Myx[(size_t)r * nx + c];
        const int v = c - o_v;
        return (v >= nu && v < nv) ? d->Myu[(size_t)r * nv + (v - nu)] : 0.0; });
    // V_i = sum_r Wrow[r][i] sq_r + sum_v Wdu[v-nu][i] dU_v^2
    C.mvn = tb.add(C.n_m, nsq + nv, [&](int i, int c) -> double {
        if (i >= m) return 0.0;
        if (c < nsq) return d->Wrow[(size_t)c * m + i];
        const int v = c - nsq;
        return v >= nu ? d->Wdu[(size_t)(v - nu) * m + i] : 0.0; });
    (void)N;

    const size_t np = tb.ptr.size(), nk = tb.kt.size(), nh = tb.hmap.size();
    std::vector<int> ib(np + nk + nh);
    memcpy(ib.data(), tb.ptr.data(), np * sizeof(int));
    memcpy(ib.data() + np, tb.kt.data(), nk * sizeof(int));
    memcpy(ib.data() + np + nk, tb.hmap.data(), nh * sizeof(int));
    cudaError_t e = cudaMalloc(&tp->ibuf, ib.size() * sizeof(int));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(tp->ibuf, ib.data(), ib.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tp->dbuf, (tb.B.size() + 32) * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(tp->dbuf, tb.B.data(), tb.B.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    C.ptr = tp->ibuf; C.kt = tp->ibuf + np; C.hmap = tp->ibuf + np + nk; C.B = tp->dbuf;
    tp->n_tiles = (int)nk;
    auto kern = tile_step_kernel<TL_NTMAX>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp->smem_bytes);
    if (e != cudaSuccess) return e;
    tp->ok = 1;
    return cudaSuccess;
}

inline cudaError_t tile_ensure_capacity(TilePlan *tp, long long batch) {
    if (tp->capacity >= batch) return cudaSuccess;
    if (tp->n_handover) cudaFree(tp->n_handover);
    tp->n_handover = nullptr; tp->start_step = nullptr; tp->capacity = 0;
    cudaError_t e = cudaMalloc(&tp->n_handover, (size_t)(batch + 1) * sizeof(int));
    if (e == cudaSuccess) { tp->start_step = tp->n_handover + 1; tp->capacity = batch; }
    return e;
}

inline cudaError_t tile_launch(const TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch,
                               int n_steps, int sm_count, cudaStream_t st) {
    const long long n_oct = (batch + TL_WARPS - 1) / TL_WARPS;
    int per_sm = (int)((227 * 1024) / (tp.smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const long long cap = (long long)sm_count * per_sm;
    const int grid = (int)(n_oct < cap ? n_oct : cap);
    tile_step_kernel<TL_NTMAX><<<grid, TL_WARPS * 32, tp.smem_bytes, st>>>(P, tp.cfg, io, batch, n_steps,
                                                                            tp.start_step, tp.n_handover);
    return cudaGetLastError();
}

inline void tile_destroy(TilePlan *tp) {
    if (tp->ibuf) cudaFree(tp->ibuf);
    if (tp->dbuf) cudaFree(tp->dbuf);
    if (tp->n_handover) cudaFree(tp->n_handover);
    memset(tp, 0, sizeof *tp);
}

}  // namespace sih
