// dist_kernel.cuh -- register-tiled fast path ("quad kernel") for small controllers.
//
// FOUR LANES PER SCENARIO, eight scenarios per warp.  Row i of every vector/matrix that belongs
// to one scenario lives in lane (i & 3) of the scenario's quad, register slot (i >> 2):
//   * the Hessian / Cholesky factor (lower triangle) is REGISTER resident, rows cyclic over the
//     quad; the factorisation is right-looking, the column being eliminated is exchanged inside
//     the quad with warp shuffles, the forward substitution is fused into the factorisation and
//     the backward substitution uses quad-wide shuffle reductions;
//   * every shared (per-controller) matrix is staged ONCE per CTA into shared memory in
//     "quad order" [in-pair][slot][lane][2] so that one LDS.128 hands each lane two
//     coefficients of its own row and the eight quads of a warp hit the same 64 bytes
//     (broadcast);
//   * per-scenario input vectors (x, omega o e, dU) are all-gathered through a small per-warp
//     shared buffer and read back with broadcast LDS.128.
// All dimensions are compile-time (template), all loops fully unrolled, no integer address
// arithmetic inside the inner products.
//
// The kernel computes the UNCONSTRAINED optimum and checks the move / input boxes.  A scenario
// whose optimum violates a bound is handed, from that step on, to the general kernel
// (step_kernel.cuh), which runs the exact dual active-set; the hand-over is per scenario
// (start_step[]), so the closed loop stays exact.
//
// Replaces the same reference code as step_kernel.cuh (ihmpccontroller.py:567-604,534-564,
// 619-636,93-113; example_distillation.py:81-113).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <string.h>
#include <stdlib.h>
#include "devplan.h"

namespace sih {

struct QuadConst {          // offsets (in doubles) into the packed constant block
    int px, py, e0, w0, wrow, et, ee, myx, myu, ht, hj, wdu, wdus, gamma, tau, bounds, cmat, vinq,
        ell_val, total_d;
    int ell_col, ell_csr, hj_idx, vin_row, total_i;   // offsets in ints after the doubles
    int ell_w;
};

template <int NX_, int NU_, int NY_, int ND_, int N_, int M_>
struct QuadDims {
    static constexpr int NX = NX_, NU = NU_, NY = NY_, ND = ND_, N = N_, M = M_;
    static constexpr int NV = N * NU;
    static constexpr int NYP = N * NY;                 // Y_pred rows
    static constexpr int R = NYP + 2 * NY + ND;        // residual rows
    static constexpr int RE = R + NY;                  // + xi(N+1) rows (warm start)
    static constexpr int SV = (NV + 3) / 4;            // register slots per lane
    static constexpr int SR = (RE + 3) / 4;
    static constexpr int SY = (NYP + 3) / 4;
    static constexpr int SX = (NX + 3) / 4;
    static constexpr int SC = (NY + 3) / 4;
    static constexpr int SQ = (NYP + 2 * NY + 3) / 4;  // rows entering the sub-objective values
    static constexpr int XUP = (NX + NU + 1) / 2;      // input pairs: [x ; du]
    static constexpr int VP = (NV + 1) / 2;
    static constexpr int RP = (R + 1) / 2;
    static constexpr int HPAIRS = SV * (SV + 1);       // sum_t (4t+4)/2
    static constexpr int BX = 2 * XUP, BT = 2 * RP, BV = 2 * VP;
    static constexpr int BRAW = BX + BT + 2 * BV;
    static constexpr int BS = BRAW + ((2 - BRAW % 4 + 4) % 4);     // BS % 4 == 2: conflict-free quads
    static constexpr int WARP_DOUBLES = 8 * BS + SR * 32;          // replicated buffers + e stash
};

// host: pack  out(rows) = Mat(rows x cols) . in  into quad order [ip][ns][4][2]
inline void quad_pack(std::vector<double> &buf, int off, const double *mat, int rows, int cols, int ld,
                      int ns, int ip, double scale) {
    for (int jp = 0; jp < ip; ++jp)
        for (int t = 0; t < ns; ++t)
            for (int g = 0; g < 4; ++g)
                for (int h = 0; h < 2; ++h) {
                    const int r = 4 * t + g, c = 2 * jp + h;
                    const double v = (r < rows && c < cols) ? mat[(size_t)r * ld + c] : 0.0;
                    buf[off + (((size_t)jp * ns + t) * 4 + g) * 2 + h] = scale * v;
                }
}

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ double quad_max(double v) {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}
__device__ __forceinline__ double quad_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }

// acc[t] += sum_j C[4t+g][j] in[j]   (constants in quad order, `in` replicated per scenario)
template <int NS, int IP>
__device__ __forceinline__ void quad_matvec(const double *__restrict__ cq, const double *__restrict__ in,
                                            double (&acc)[NS], int g) {
    const double2 *c2 = reinterpret_cast<const double2 *>(cq) + g;
    const double2 *i2 = reinterpret_cast<const double2 *>(in);
#pragma unroll
    for (int jp = 0; jp < IP; ++jp) {
        const double2 xi = i2[jp];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const double2 c = c2[(jp * NS + t) * 4];
            acc[t] = fma(c.x, xi.x, acc[t]);
            acc[t] = fma(c.y, xi.y, acc[t]);
        }
    }
}

template <int K> __device__ __forceinline__ double sel(const double (&a)[K], int i) {
    double v = a[0];
#pragma unroll
    for (int k = 1; k < K; ++k) v = (i == k) ? a[k] : v;
    return v;
}

// 1/sqrt(d) for a positive normal d: MUFU.RSQ64H seed (about 2^-22) + two Newton steps
// (error ~2^-43 then rounding level).  The library rsqrt() carries a longer dependent chain plus
// special-case branches; the factorisation's critical path is one of these per column.
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

// column K of the right-looking factorisation (rows cyclic over the quad), forward substitution fused
template <typename D, int K>
__device__ __forceinline__ void chol_columns(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g, int &notpd) {
    if constexpr (K < D::NV) {
        constexpr int SV = D::SV, NV = D::NV;
        constexpr int tk = K >> 2, gk = K & 3;
        double d = quad_bcast(L[tk][K], gk);
        const double fk = quad_bcast(fr[tk], gk);
        if (!(d > 0.0)) { notpd = 1; d = 1e-300; }
        const double inv = fast_rsqrt(d);
        const double yk = fk * inv;
        double lik[SV];
#pragma unroll
        for (int t = 0; t < SV; ++t) lik[t] = 0.0;
#pragma unroll
        for (int t = tk; t < SV; ++t)
            if (4 * t + 3 > K) lik[t] = (4 * t + g > K) ? L[t][K] * inv : 0.0;
#pragma unroll
        for (int j = K + 1; j < NV; ++j) {
            const double ljk = quad_bcast(lik[j >> 2], j & 3);
#pragma unroll
            for (int t = (j >> 2); t < SV; ++t) L[t][j] = fma(-lik[t], ljk, L[t][j]);
        }
#pragma unroll
        for (int t = tk; t < SV; ++t)
            if (4 * t + 3 > K) { fr[t] = fma(-lik[t], yk, fr[t]); L[t][K] = lik[t]; }
        if (g == gk) { fr[tk] = yk; L[tk][K] = inv; }     // diagonal slot keeps 1/L_kk
        chol_columns<D, K + 1>(L, fr, g, notpd);
    }
}

// row J of the backward substitution L^T z = y (dot products over the quad)
template <typename D, int J>
__device__ __forceinline__ void back_rows(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g) {
    if constexpr (J >= 0) {
        constexpr int SV = D::SV;
        constexpr int tj = J >> 2, gj = J & 3;
        double part = 0.0;
#pragma unroll
        for (int t = tj; t < SV; ++t)
            if (4 * t + 3 > J) part = fma((4 * t + g > J) ? L[t][J] : 0.0, fr[t], part);
        const double sum = quad_sum(part);
        if (g == gj) fr[tj] = (fr[tj] - sum) * L[tj][J];
        back_rows<D, J - 1>(L, fr, g);
    }
}

// ---- block LDL^T with 2x2 pivots (default solver of the fast paths) ------------------------------
// H = L D L^T, D block diagonal (2x2 blocks), L unit block-lower.  Ten sequential rounds instead of
// twenty, no square roots, and the exchange of the two pivot columns does not depend on the
// reciprocal of the block determinant, so it overlaps with that chain.  Positive definiteness is
// checked through the leading minors of every block (a > 0, det > 0).
// Storage after block B (columns k = 2B, k+1; slot s = B/2, lanes g0 = 2(B%2), g0+1):
//   rows below the block:  L[t][k] = w1, L[t][k+1] = w2     (W = [p q] D^-1)
//   pivot rows:            lane g0: L[s][k] = D^-1_11 ; lane g0+1: L[s][k] = D^-1_21, L[s][k+1] = D^-1_22
// p ? a : b as a SELECT whatever the compiler's if-conversion heuristics think: ptxas turns a conditional 64-bit move
// on a per-lane predicate into a divergent branch (BSSY / BRA / BSYNC), which costs ~20 cycles on a dependent chain
__device__ __forceinline__ double sel_f64(const bool p, const double a, const double b) {
    double r;
    asm("{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n selp.f64 %0, %1, %2, q;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
    return r;
}
__device__ __forceinline__ double fast_rcp(const double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    return y;
}

template <typename D, int B>
__device__ __forceinline__ void ldl_blocks(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g, int &notpd) {
    if constexpr (2 * B < D::NV) {
        constexpr int SV = D::SV, NV = D::NV;
        constexpr int k = 2 * B, s = B / 2, g0 = 2 * (B % 2);
        double a = quad_bcast(L[s][k], g0);
        const double bb = quad_bcast(L[s][k], g0 + 1);
        const double c = quad_bcast(L[s][k + 1], g0 + 1);
        const double yk = quad_bcast(fr[s], g0), yk1 = quad_bcast(fr[s], g0 + 1);
        double det = fma(a, c, -bb * bb);
        {   // selects, not a branch: the rounds are one dependent chain and a divergent branch (BSSY/BSYNC) sits on it
            const bool bad = !(a > 0.0) || !(det > 0.0);
            notpd |= bad ? 1 : 0;
            det = sel_f64(bad, 1e300, det);
            a = sel_f64(bad, 1.0, a);
        }
        const double idet = fast_rcp(det);
        const double d11 = c * idet, d21 = -bb * idet, d22 = a * idet;
        double pq[SV][2], w[SV][2];
#pragma unroll
        for (int t = 0; t < SV; ++t) { pq[t][0] = pq[t][1] = 0.0; w[t][0] = w[t][1] = 0.0; }
#pragma unroll
        for (int t = s; t < SV; ++t)
            if (4 * t + 3 > k + 1) {
                const bool below = (4 * t + g > k + 1);
                pq[t][0] = below ? L[t][k] : 0.0;
                pq[t][1] = below ? L[t][k + 1] : 0.0;
                w[t][0] = fma(pq[t][0], d11, pq[t][1] * d21);
                w[t][1] = fma(pq[t][0], d21, pq[t][1] * d22);
            }
#pragma unroll
        for (int j = k + 2; j < NV; ++j) {
            const double pj = quad_bcast(pq[j >> 2][0], j & 3), qj = quad_bcast(pq[j >> 2][1], j & 3);
#pragma unroll
            for (int t = (j >> 2); t < SV; ++t) L[t][j] = fma(-w[t][1], qj, fma(-w[t][0], pj, L[t][j]));
        }
#pragma unroll
        for (int t = s; t < SV; ++t)
            if (4 * t + 3 > k + 1) {
                fr[t] = fma(-w[t][1], yk1, fma(-w[t][0], yk, fr[t]));
                const bool below = (4 * t + g > k + 1);
                L[t][k] = sel_f64(below, w[t][0], L[t][k]);
                L[t][k + 1] = sel_f64(below, w[t][1], L[t][k + 1]);
            }
        L[s][k] = sel_f64(g == g0, d11, sel_f64(g == g0 + 1, d21, L[s][k]));
        L[s][k + 1] = sel_f64(g == g0 + 1, d22, L[s][k + 1]);
        ldl_blocks<D, B + 1>(L, fr, g, notpd);
    }
}

// z' = D^-1 y for both blocks of every slot at once (partners exchange with one xor-shuffle)
template <typename D>
__device__ __forceinline__ void ldl_dsolve(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g) {
#pragma unroll
    for (int s = 0; s < D::SV; ++s) {
        const bool hi = (g & 2) != 0, odd = (g & 1) != 0;
        const double Lk = hi ? L[s][4 * s + 2] : L[s][4 * s];
        const double Lk1 = hi ? L[s][4 * s + 3] : L[s][4 * s + 1];
        const double py = __shfl_xor_sync(0xffffffffu, fr[s], 1);
        const double pL = __shfl_xor_sync(0xffffffffu, Lk, 1);
        // even lane: z' = D11 y + D21 y_partner ; odd lane: z' = D21 y_partner + D22 y
        fr[s] = odd ? fma(Lk, py, Lk1 * fr[s]) : fma(Lk, fr[s], pL * py);
    }
}

// L^T z = z', blocks from last to first
template <typename D, int B>
__device__ __forceinline__ void ldl_back(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g) {
    if constexpr (B >= 0) {
        constexpr int SV = D::SV;
        constexpr int k = 2 * B, s = B / 2, g0 = 2 * (B % 2);
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int t = SV - 1; t >= s; --t)
            if (4 * t + 3 > k + 1) {
                const bool below = (4 * t + g > k + 1);
                s1 = fma(below ? L[t][k] : 0.0, fr[t], s1);
                s2 = fma(below ? L[t][k + 1] : 0.0, fr[t], s2);
            }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        fr[s] -= sel_f64(g == g0, s1, sel_f64(g == g0 + 1, s2, 0.0));
        ldl_back<D, B - 1>(L, fr, g);
    }
}

// fr <- H^{-1} fr   (H in L, lower triangle, rows cyclic over the quad)
template <typename D>
__device__ __forceinline__ void quad_solve_spd(double (&L)[D::SV][D::NV], double (&fr)[D::SV], const int g, int &notpd) {
#ifdef SIH_SOLVER_CHOLESKY
    chol_columns<D, 0>(L, fr, g, notpd);
    back_rows<D, D::NV - 1>(L, fr, g);
#else
    ldl_blocks<D, 0>(L, fr, g, notpd);
    ldl_dsolve<D>(L, fr, g);
    ldl_back<D, D::NV / 2 - 1>(L, fr, g);
#endif
}

template <typename D, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
quad_step_kernel(const DevPlan P, const QuadConst Q, const double *__restrict__ qbuf, const sihmpc_io io,
                 const long long batch, const int n_steps, int *__restrict__ start_step,
                 int *__restrict__ n_handover) {
    constexpr int NX = D::NX, NU = D::NU, NY = D::NY, M = D::M;
    constexpr int NV = D::NV, NYP = D::NYP, R = D::R;
    constexpr int SV = D::SV, SR = D::SR, SY = D::SY, SX = D::SX, SC = D::SC, SQ = D::SQ;
    constexpr int XUP = D::XUP, VP = D::VP, RP = D::RP;
    constexpr int BS = D::BS;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane & 3, q = lane >> 2;
    // ---- stage the controller's constants (once per CTA)
    double *cst = smem;
    for (int i = threadIdx.x; i < Q.total_d; i += blockDim.x) cst[i] = qbuf[i];
    int *cint = reinterpret_cast<int *>(cst + Q.total_d);
    {
        const int *src = reinterpret_cast<const int *>(qbuf + Q.total_d);
        for (int i = threadIdx.x; i < Q.total_i; i += blockDim.x) cint[i] = src[i];
    }
    __syncthreads();
    double *wbase = cst + Q.total_d + (((Q.total_i + 3) >> 2) << 1) + (size_t)wib * D::WARP_DOUBLES;
    double *rep = wbase + q * BS;                          // this scenario's replicated buffers
    double *bx = rep, *bt = rep + D::BX, *bv = bt + D::BT, *bvs = bv + D::BV;
    double *stash = wbase + 8 * BS + lane;                 // [slot][32 lanes]

    const long long gq = ((long long)blockIdx.x * WARPS + wib) * 8 + q;    // global quad index
    const long long nq = (long long)gridDim.x * WARPS * 8;
    const long long rounds = (batch + nq - 1) / nq;
    const int nstat = NY + SIHMPC_NSTATS_EXTRA;
    const bool advance = (io.flags & SIHMPC_F_ADVANCE) != 0;
    const int W = Q.ell_w;
    const int mh = P.mh;

    for (long long rd = 0; rd < rounds; ++rd) {
        const long long s = gq + rd * nq;
        const bool valid = s < batch;
        const long long sc = valid ? s : 0;
        bool alive = valid;                       // false once handed over to the general kernel
        // ---- load state: x replicated in smem, small vectors replicated in registers
        __syncwarp();
        for (int i = g; i < D::BX; i += 4) bx[i] = (i < NX) ? io.x[sc * NX + i] : 0.0;
        double uu[NU], pes[M], ise[SC];
        double st_du2 = 0.0, st_viol = 0.0, st_J = 0.0;
        int st_or = 0;
#pragma unroll
        for (int j = 0; j < NU; ++j) uu[j] = io.u[sc * NU + j];
#pragma unroll
        for (int i = 0; i < M; ++i) pes[i] = io.pesos[sc * M + i];
#pragma unroll
        for (int t = 0; t < SC; ++t) ise[t] = 0.0;
        if (io.stats) {
#pragma unroll
            for (int t = 0; t < SC; ++t) ise[t] = (4 * t + g < NY) ? io.stats[sc * nstat + 4 * t + g] : 0.0;
            st_du2 = io.stats[sc * nstat + NY]; st_viol = io.stats[sc * nstat + NY + 1];
            st_J = io.stats[sc * nstat + NY + 2]; st_or = (int)io.stats[sc * nstat + NY + 3];
        }
        __syncwarp();

        for (int step = 0; step < n_steps; ++step) {
            const int kabs = io.k0 + step;
            double ysp[NY];
            if (io.ysp_seq) {
                int seg = 0;
                for (int a = 1; a < io.n_seg; ++a)
                    if (io.seg_start[sc * io.n_seg + a] <= kabs) seg = a;
#pragma unroll
                for (int i = 0; i < NY; ++i) ysp[i] = io.ysp_seq[(sc * io.n_seg + seg) * NY + i];
            } else {
#pragma unroll
                for (int i = 0; i < NY; ++i) ysp[i] = io.ysp[sc * NY + i];
            }
            // ---------------- K1: e = Px x + Py ysp + e0 (rows R.. hold xi(N+1)), t = omega o e
            {
                double er[SR];
#pragma unroll
                for (int t = 0; t < SR; ++t) er[t] = cst[Q.e0 + t * 4 + g];
                quad_matvec<SR, XUP>(cst + Q.px, bx, er, g);
#pragma unroll
                for (int t = 0; t < SR; ++t) {
                    double o = cst[Q.w0 + t * 4 + g];
#pragma unroll
                    for (int i = 0; i < M; ++i) o = fma(cst[Q.wrow + (t * 4 + g) * M + i], pes[i], o);
#pragma unroll
                    for (int c = 0; c < NY; ++c) er[t] = fma(cst[Q.py + (t * 4 + g) * NY + c], ysp[c], er[t]);
                    const int r = 4 * t + g;
                    if (r < D::BT) bt[r] = (r < R) ? o * er[t] : 0.0;
                    stash[t * 32] = er[t];
                }
            }
            __syncwarp();
            // ---------------- gradient rows: fr = -(2 E^T t)   (factor 2 folded into the constants)
            double fr[SV];
#pragma unroll
            for (int t = 0; t < SV; ++t) fr[t] = 0.0;
            quad_matvec<SV, RP>(cst + Q.et, bt, fr, g);
#pragma unroll
            for (int t = 0; t < SV; ++t) fr[t] = -fr[t];
            // ---------------- Hessian rows (lower triangle) into registers
            double L[SV][NV];
            {
                double pw[M];
#pragma unroll
                for (int a = 0; a < M; ++a) pw[a] = (a < mh) ? sel<M>(pes, cint[Q.hj_idx + a]) : 0.0;
                const double2 *h0 = reinterpret_cast<const double2 *>(cst + Q.ht) + g;
                const double2 *hj = reinterpret_cast<const double2 *>(cst + Q.hj) + g;
                int pp = 0;
#pragma unroll
                for (int t = 0; t < SV; ++t) {
                    double o = 0.0;                        // 2 * omega_du goes on the diagonal
#pragma unroll
                    for (int i = 0; i < M; ++i) o = fma(cst[Q.wdu + (t * 4 + g) * M + i], pes[i], o);
                    o *= 2.0;
                    // the diagonal of row 4t+g sits in pair 2t + (g>>1), half (g&1): arithmetic
                    // selects only, so that L stays in registers (no per-lane array indexing)
                    const double dlo = (g & 1) ? 0.0 : o, dhi = (g & 1) ? o : 0.0;
#pragma unroll
                    for (int jp = 0; jp < (4 * t + 4) / 2; ++jp) {
                        double2 h = h0[pp * 4];
#pragma unroll
                        for (int a = 0; a < M; ++a)
                            if (a < mh) {
                                const double2 ha = hj[(a * D::HPAIRS + pp) * 4];
                                h.x = fma(pw[a], ha.x, h.x);
                                h.y = fma(pw[a], ha.y, h.y);
                            }
                        if (jp == 2 * t || jp == 2 * t + 1) {
                            const bool mine = (jp - 2 * t) == (g >> 1);
                            h.x += mine ? dlo : 0.0;
                            h.y += mine ? dhi : 0.0;
                        }
                        if (2 * jp < NV) L[t][2 * jp] = h.x;
                        if (2 * jp + 1 < NV) L[t][2 * jp + 1] = h.y;
                        ++pp;
                    }
                }
            }
            // ---------------- K2: right-looking Cholesky with fused forward substitution, then the
            // backward substitution (compile-time column index: template recursion, so that L
            // never leaves the register file)
            int notpd = 0;
            quad_solve_spd<D>(L, fr, g, notpd);
            // all-gather dU
#pragma unroll
            for (int t = 0; t < SV; ++t)
                if (4 * t + g < D::BV) bv[4 * t + g] = (4 * t + g < NV) ? fr[t] : 0.0;
            __syncwarp();
            // ---------------- feasibility of the unconstrained optimum (move box, input box)
            double viol = 0.0, viol_abs = 0.0;
            if (P.has_box) {
#pragma unroll
                for (int t = 0; t < SV; ++t) {
                    const int i = 4 * t + g;
                    if (i < NV) {
                        const int j = i % NU;
                        const double lo = cst[Q.bounds + j], hi = cst[Q.bounds + NU + j];
                        if (lo > -SIH_INF) { const double a = lo - fr[t]; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(lo))); }
                        if (hi < SIH_INF) { const double a = fr[t] - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(hi))); }
                    }
                }
            }
            if (P.has_cum) {
                for (int j = g; j < NU; j += 4) {
                    const double uj = sel<NU>(uu, j);
                    const double lo = cst[Q.bounds + 2 * NU + j] - uj, hi = cst[Q.bounds + 3 * NU + j] - uj;
                    double cs = 0.0, cmin = SIH_INF, cmax = -SIH_INF;
#pragma unroll
                    for (int k = 0; k < D::N; ++k) { cs += bv[k * NU + j]; cmin = fmin(cmin, cs); cmax = fmax(cmax, cs); }
                    if (lo > -SIH_INF) { const double a = lo - cmin; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(lo))); }
                    if (hi < SIH_INF) { const double a = cmax - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, a / (1.0 + fabs(hi))); }
                }
            }
            viol = quad_max(viol);
            viol_abs = quad_max(viol_abs);
            if (alive && (viol > 1e-10 || notpd)) {
                // hand this scenario to the general kernel from this step on; its state is still
                // the state at the start of the step (nothing has been written yet)
                alive = false;
                if (g == 0) { start_step[s] = step; atomicAdd(n_handover, 1); }
                for (int i = g; i < NX; i += 4) io.x[s * NX + i] = bx[i];
                if (g == 0) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) io.u[s * NU + j] = uu[j];
#pragma unroll
                    for (int i = 0; i < M; ++i) io.pesos[s * M + i] = pes[i];
                }
                if (io.stats) {
#pragma unroll
                    for (int t = 0; t < SC; ++t) if (4 * t + g < NY) io.stats[s * nstat + 4 * t + g] = ise[t];
                    if (g == 0) {
                        io.stats[s * nstat + NY] = st_du2; io.stats[s * nstat + NY + 1] = st_viol;
                        io.stats[s * nstat + NY + 2] = st_J; io.stats[s * nstat + NY + 3] = (double)st_or;
                    }
                }
            }
            // ---------------- K3: residual r = e + E dU, cost, slacks
            double syv[3 * NY];                    // [syN | siN | siN_next]
            double Jval;
            {
                double rr[SR];
#pragma unroll
                for (int t = 0; t < SR; ++t) rr[t] = stash[t * 32];
                quad_matvec<SR, VP>(cst + Q.ee, bv, rr, g);
                double ja = 0.0;
#pragma unroll
                for (int t = 0; t < SR; ++t) {
                    double o = cst[Q.w0 + t * 4 + g];
#pragma unroll
                    for (int i = 0; i < M; ++i) o = fma(cst[Q.wrow + (t * 4 + g) * M + i], pes[i], o);
                    ja = fma(o, rr[t] * rr[t], ja);          // rows >= R carry zero weight
                }
#pragma unroll
                for (int t = 0; t < SV; ++t) {
                    double o = 0.0;
#pragma unroll
                    for (int i = 0; i < M; ++i) o = fma(cst[Q.wdu + (t * 4 + g) * M + i], pes[i], o);
                    ja = fma(o, fr[t] * fr[t], ja);          // fr = dU rows (zero beyond NV)
                }
                Jval = quad_sum(ja);
#pragma unroll
                for (int i = 0; i < NY; ++i) {
                    syv[i] = quad_bcast(rr[(NYP + i) >> 2], (NYP + i) & 3);
                    syv[NY + i] = quad_bcast(rr[(NYP + NY + i) >> 2], (NYP + NY + i) & 3);
                    syv[2 * NY + i] = quad_bcast(rr[(R + i) >> 2], (R + i) & 3);
                }
            }
            const bool last_step = (step == n_steps - 1);
            if (alive) {
                if (last_step) {
                    if (io.du) for (int j = g; j < NU; j += 4) io.du[s * NU + j] = bv[j];
                    if (io.w_opt) {
                        constexpr int nw = NV + 2 * NY;
#pragma unroll
                        for (int t = 0; t < SV; ++t) if (4 * t + g < NV) io.w_opt[s * nw + 4 * t + g] = fr[t];
                        if (g == 0) {
#pragma unroll
                            for (int i = 0; i < 2 * NY; ++i) io.w_opt[s * nw + NV + i] = syv[i];
                        }
                    }
                    if (g == 0) {
                        if (io.J) io.J[s] = Jval;
                        if (io.status) io.status[s] = 0;
                        if (io.n_active) io.n_active[s] = 0;
                    }
                }
                if (io.traj_du) for (int j = g; j < NU; j += 4) io.traj_du[((long long)step * batch + s) * NU + j] = bv[j];
                if (io.traj_J && g == 0) io.traj_J[(long long)step * batch + s] = Jval;
                if ((io.flags & SIHMPC_F_UPDATE_VIN) && io.vinant && (last_step || !advance)) {
                    for (int sl = g; sl < P.n_vin; sl += 4) {
                        const double sn = sel<3 * NY>(syv, 2 * NY + cint[Q.vin_row + sl]);
                        io.vinant[s * P.n_vin + sl] = cst[Q.vinq + sl] * sn * sn;
                    }
                }
            }
            if (!advance) continue;
            // ---------------- advance: plant x+ = [A|B][x;du], u+ = u + du
            for (int j = g; j < NU; j += 4) bx[NX + j] = bv[j];
            __syncwarp();
            double xn[SX];
            {
                const double *pv = io.plant_val ? io.plant_val + (size_t)sc * P.pl_nnz : nullptr;
#pragma unroll
                for (int t = 0; t < SX; ++t) {
                    double a = 0.0;
                    for (int w = 0; w < W; ++w) {
                        const int e = (t * W + w) * 4 + g;
                        const int ci = cint[Q.ell_csr + e];
                        const double val = (pv && ci >= 0) ? pv[ci] : cst[Q.ell_val + e];
                        a = fma(val, bx[cint[Q.ell_col + e]], a);
                    }
                    xn[t] = a;
                }
            }
            double du2 = 0.0;
#pragma unroll
            for (int j = 0; j < NU; ++j) { const double d = bv[j]; uu[j] += d; du2 = fma(d, d, du2); }
            if (io.dist_step && io.dist_step[sc] == kabs) {
#pragma unroll
                for (int t = 0; t < SC; ++t) if (4 * t + g < NY) xn[t] += io.dist[sc * NY + 4 * t + g];
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < SX; ++t) if (4 * t + g < NX) bx[4 * t + g] = xn[t];
            // shifted plan (ihmpccontroller.py:538-541)
#pragma unroll
            for (int t = 0; t < SV; ++t) {
                const int i = 4 * t + g;
                if (i < D::BV) bvs[i] = (i + NU < NV) ? bv[i + NU] : 0.0;
            }
            __syncwarp();
            // measured output y = C x+ + D du and closed-loop statistics
            {
                double yo[SC];
#pragma unroll
                for (int t = 0; t < SC; ++t) yo[t] = 0.0;
                quad_matvec<SC, XUP>(cst + Q.cmat, bx, yo, g);
#pragma unroll
                for (int t = 0; t < SC; ++t) {
                    const int i = 4 * t + g;
                    if (i < NY) {
                        const double err = yo[t] - sel<NY>(ysp, i);
                        ise[t] = fma(err, err, ise[t]);
                        if (alive && io.traj_y) io.traj_y[((long long)step * batch + s) * NY + i] = yo[t];
                    }
                }
                st_du2 += du2; st_viol = fmax(st_viol, viol_abs); st_J = Jval;
            }
            // ---------------- sub-objective values at the new state and weight update
            if (io.flags & (SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1)) {
                double yp[SY];
#pragma unroll
                for (int t = 0; t < SY; ++t) yp[t] = 0.0;
                quad_matvec<SY, XUP>(cst + Q.myx, bx, yp, g);
                quad_matvec<SY, VP>(cst + Q.myu, bvs, yp, g);
                double vn[M];
#pragma unroll
                for (int i = 0; i < M; ++i) vn[i] = 0.0;
#pragma unroll
                for (int t = 0; t < SQ; ++t) {
                    const int r = 4 * t + g;
                    double sq;
                    if (r < NYP) {
                        const int i = r % NY;
                        const double res = (t < SY ? yp[t < SY ? t : 0] : 0.0) - sel<NY>(ysp, i) - sel<3 * NY>(syv, i)
                                           - cst[Q.tau + (t < SY ? t : 0) * 4 + g] * sel<3 * NY>(syv, 2 * NY + i);
                        sq = res * res;
                    } else if (r < NYP + NY) {
                        const double a = sel<3 * NY>(syv, r - NYP); sq = a * a;
                    } else if (r < NYP + 2 * NY) {
                        const double a = sel<3 * NY>(syv, 2 * NY + (r - NYP - NY)); sq = a * a;
                    } else sq = 0.0;
#pragma unroll
                    for (int i = 0; i < M; ++i) vn[i] = fma(cst[Q.wrow + (t * 4 + g) * M + i], sq, vn[i]);
                }
#pragma unroll
                for (int t = 0; t < SV; ++t) {
                    const double a = fr[t] * fr[t];
#pragma unroll
                    for (int i = 0; i < M; ++i) vn[i] = fma(cst[Q.wdus + (t * 4 + g) * M + i], a, vn[i]);
                }
#pragma unroll
                for (int i = 0; i < M; ++i) vn[i] = quad_sum(vn[i]);
                if (io.flags & SIHMPC_F_WEIGHTS_SAT2) {
                    double smax = -SIH_INF;
#pragma unroll
                    for (int i = 0; i < M; ++i) smax = fmax(smax, vn[i] / cst[Q.gamma + i]);
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        const double g0 = cst[Q.gamma + i];
                        const double g1 = fmax(g0, g0 * smax + 1e-6);
                        pes[i] = io.alfa * pes[i] + (1.0 - io.alfa) * (1.0 / (g1 - vn[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        const double g0 = cst[Q.gamma + i];
                        pes[i] = 1.0 / (g0 - fmin(fmax(vn[i], 0.0), 0.99 * g0));
                    }
                }
            }
        }  // steps

        // ---------------- write back
        __syncwarp();            // warp-wide, OUTSIDE the divergent branch (a handed-over quad skips the block)
        if (alive && advance) {
            for (int i = g; i < NX; i += 4) io.x[s * NX + i] = bx[i];
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < NU; ++j) io.u[s * NU + j] = uu[j];
#pragma unroll
                for (int i = 0; i < M; ++i) io.pesos[s * M + i] = pes[i];
            }
            if (io.stats) {
#pragma unroll
                for (int t = 0; t < SC; ++t) if (4 * t + g < NY) io.stats[s * nstat + 4 * t + g] = ise[t];
                if (g == 0) {
                    io.stats[s * nstat + NY] = st_du2; io.stats[s * nstat + NY + 1] = st_viol;
                    io.stats[s * nstat + NY + 2] = st_J; io.stats[s * nstat + NY + 3] = (double)st_or;
                }
            }
        }
    }
}

#ifndef SIH_NO_QUAD_HOST   // (the tensor fast path includes this file for the factorisation helpers only)
// ---- host side: shape table, packing, launch ---------------------------------------------------
struct QuadPlan {
    int shape;                 // index into the instantiation table, -1 = none
    QuadConst qc;
    double *qbuf;              // device
    int *start_step;           // device [capacity]
    int *n_handover;           // device
    long long capacity;
    int grid_cap;              // debug: upper limit of the grid (0: none)
    int blocks_per_sm;
    size_t smem_bytes;
};

using DistillationDims = QuadDims<22, 2, 2, 4, 10, 4>;
constexpr int QUAD_WARPS = 4;
#ifndef SIH_QUAD_MINB
#define SIH_QUAD_MINB 2
#endif
constexpr int QUAD_MINB = SIH_QUAD_MINB;

inline int quad_match_shape(const DevPlan &P) {
    if (P.nx == 22 && P.nu == 2 && P.ny == 2 && P.nd == 4 && P.N == 10 && P.m == 4 && P.ncd == 0 && P.nv == P.nvm) return 0;
    return -1;
}

template <typename D>
inline void quad_build_constants(const sihmpc_plan_desc *d, QuadConst &qc, std::vector<double> &buf,
                                 std::vector<int> &ibuf) {
    constexpr int NX = D::NX, NU = D::NU, NY = D::NY, M = D::M, NV = D::NV, R = D::R, RE = D::RE;
    constexpr int NYP = D::NYP;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    qc.px = take(D::XUP * D::SR * 8);
    qc.py = take(D::SR * 4 * NY);
    qc.e0 = take(D::SR * 4);
    qc.w0 = take(D::SR * 4);
    qc.wrow = take(D::SR * 4 * M);
    qc.et = take(D::RP * D::SV * 8);
    qc.ee = take(D::VP * D::SR * 8);
    qc.myx = take(D::XUP * D::SY * 8);
    qc.myu = take(D::VP * D::SY * 8);
    qc.ht = take(D::HPAIRS * 8);
    qc.hj = take(M * D::HPAIRS * 8);
    qc.wdu = take(D::SV * 4 * M);
    qc.wdus = take(D::SV * 4 * M);
    qc.gamma = take(M);
    qc.tau = take(D::SY * 4);
    qc.bounds = take(4 * NU);
    qc.cmat = take(D::XUP * D::SC * 8);
    qc.vinq = take(d->n_vin > 0 ? d->n_vin : 1);
    // ELL plant
    int W = 1;
    for (int i = 0; i < NX; ++i) { int n = d->pl_rowptr[i + 1] - d->pl_rowptr[i]; if (n > W) W = n; }
    qc.ell_w = W;
    qc.ell_val = take(D::SX * W * 4);
    qc.total_d = o;
    buf.assign(o, 0.0);
    // extended matrices: rows R.. = xi(N+1) maps
    std::vector<double> PxE((size_t)RE * NX, 0.0), EE((size_t)RE * NV, 0.0), ET2((size_t)NV * R, 0.0);
    for (int r = 0; r < R; ++r) {
        for (int j = 0; j < NX; ++j) PxE[(size_t)r * NX + j] = d->Px[(size_t)r * NX + j];
        for (int c = 0; c < NV; ++c) { EE[(size_t)r * NV + c] = d->E[(size_t)r * NV + c]; ET2[(size_t)c * R + r] = d->E[(size_t)r * NV + c]; }
    }
    for (int i = 0; i < NY; ++i) {
        for (int j = 0; j < NX; ++j) PxE[(size_t)(R + i) * NX + j] = d->Mi1x[(size_t)i * NX + j];
        for (int c = 0; c < NV; ++c) EE[(size_t)(R + i) * NV + c] = d->Mi1u[(size_t)i * NV + c];
    }
    quad_pack(buf, qc.px, PxE.data(), RE, NX, NX, D::SR, D::XUP, 1.0);
    quad_pack(buf, qc.et, ET2.data(), NV, R, R, D::SV, D::RP, 2.0);
    quad_pack(buf, qc.ee, EE.data(), RE, NV, NV, D::SR, D::VP, 1.0);
    quad_pack(buf, qc.myx, d->Myx, NYP, NX, NX, D::SY, D::XUP, 1.0);
    quad_pack(buf, qc.myu, d->Myu, NYP, NV, NV, D::SY, D::VP, 1.0);
    {   // [C | D] over [x ; du]
        std::vector<double> CD((size_t)NY * (NX + NU), 0.0);
        for (int i = 0; i < NY; ++i) {
            for (int j = 0; j < NX; ++j) CD[(size_t)i * (NX + NU) + j] = d->C[(size_t)i * NX + j];
            for (int j = 0; j < NU; ++j) CD[(size_t)i * (NX + NU) + NX + j] = d->D[(size_t)i * NU + j];
        }
        quad_pack(buf, qc.cmat, CD.data(), NY, NX + NU, NX + NU, D::SC, D::XUP, 1.0);
    }
    for (int t = 0; t < D::SR; ++t)
        for (int g = 0; g < 4; ++g) {
            const int r = 4 * t + g;
            if (r < R) {
                buf[qc.e0 + t * 4 + g] = d->e0[r];
                buf[qc.w0 + t * 4 + g] = d->w0row[r];
                for (int i = 0; i < M; ++i) buf[qc.wrow + (t * 4 + g) * M + i] = d->Wrow[(size_t)r * M + i];
                for (int c = 0; c < NY; ++c) buf[qc.py + (t * 4 + g) * NY + c] = d->Py[(size_t)r * NY + c];
            }
        }
    for (int t = 0; t < D::SV; ++t)
        for (int g = 0; g < 4; ++g) {
            const int j = 4 * t + g;
            if (j < NV) {
                for (int i = 0; i < M; ++i) {
                    buf[qc.wdu + (t * 4 + g) * M + i] = d->Wdu[(size_t)j * M + i];
                    buf[qc.wdus + (t * 4 + g) * M + i] = (j >= NU) ? d->Wdu[(size_t)(j - NU) * M + i] : 0.0;
                }
            }
        }
    // Hessian pairs
    {
        int pp = 0;
        for (int t = 0; t < D::SV; ++t)
            for (int jp = 0; jp < (4 * t + 4) / 2; ++jp, ++pp)
                for (int g = 0; g < 4; ++g)
                    for (int h = 0; h < 2; ++h) {
                        const int i = 4 * t + g, j = 2 * jp + h;
                        const bool in = (i < NV && j <= i);
                        buf[qc.ht + (pp * 4 + g) * 2 + h] = in ? d->Ht[(size_t)i * NV + j] : 0.0;
                        for (int a = 0; a < d->mh; ++a)
                            buf[qc.hj + (((size_t)a * D::HPAIRS + pp) * 4 + g) * 2 + h] =
                                in ? d->HJ[((size_t)a * NV + i) * NV + j] : 0.0;
                    }
    }
    for (int i = 0; i < M; ++i) buf[qc.gamma + i] = d->gamma[i];
    for (int t = 0; t < D::SY; ++t)
        for (int g = 0; g < 4; ++g) { const int r = 4 * t + g; if (r < NYP) buf[qc.tau + t * 4 + g] = d->tau[r / NY]; }
    for (int j = 0; j < NU; ++j) {
        buf[qc.bounds + j] = d->dulb[j]; buf[qc.bounds + NU + j] = d->duub[j];
        buf[qc.bounds + 2 * NU + j] = d->ulb[j]; buf[qc.bounds + 3 * NU + j] = d->uub[j];
    }
    for (int sIdx = 0; sIdx < d->n_vin; ++sIdx) buf[qc.vinq + sIdx] = d->vin_q[sIdx];
    // int tables
    int io_ = 0;
    auto takei = [&](int n) { int r = io_; io_ += n; return r; };
    qc.ell_col = takei(D::SX * W * 4);
    qc.ell_csr = takei(D::SX * W * 4);
    qc.hj_idx = takei(M);
    qc.vin_row = takei(d->n_vin > 0 ? d->n_vin : 1);
    qc.total_i = io_;
    ibuf.assign(io_, 0);
    for (int t = 0; t < D::SX; ++t)
        for (int g = 0; g < 4; ++g) {
            const int i = 4 * t + g;
            for (int w = 0; w < W; ++w) {
                const int e = (t * W + w) * 4 + g;
                ibuf[qc.ell_csr + e] = -1; ibuf[qc.ell_col + e] = 0;
                if (i < NX) {
                    const int p = d->pl_rowptr[i] + w;
                    if (p < d->pl_rowptr[i + 1]) {
                        ibuf[qc.ell_csr + e] = p; ibuf[qc.ell_col + e] = d->pl_col[p];
                        buf[qc.ell_val + e] = d->pl_val[p];
                    }
                }
            }
        }
    for (int a = 0; a < d->mh; ++a) ibuf[qc.hj_idx + a] = d->hj_index[a];
    for (int sIdx = 0; sIdx < d->n_vin; ++sIdx) ibuf[qc.vin_row + sIdx] = d->vin_row[sIdx] - (NYP + NY);
}

template <typename D>
inline size_t quad_smem_bytes(const QuadConst &qc) {
    return ((size_t)qc.total_d + (((qc.total_i + 3) >> 2) << 1) + (size_t)QUAD_WARPS * D::WARP_DOUBLES) * sizeof(double);
}

// returns cudaSuccess and fills qp (qp->shape = -1 when the controller has no fast path)
inline cudaError_t quad_configure(const sihmpc_plan_desc *d, const DevPlan &P, QuadPlan *qp) {
    memset(qp, 0, sizeof *qp);
    qp->shape = quad_match_shape(P);
    if (qp->shape < 0) return cudaSuccess;
    std::vector<double> buf; std::vector<int> ibuf;
    using D = DistillationDims;
    quad_build_constants<D>(d, qp->qc, buf, ibuf);
    const size_t nd = buf.size(), ni = ibuf.size();
    std::vector<double> all(nd + ((ni + 1) >> 1), 0.0);
    memcpy(all.data(), buf.data(), nd * sizeof(double));
    memcpy(all.data() + nd, ibuf.data(), ni * sizeof(int));
    cudaError_t e = cudaMalloc(&qp->qbuf, all.size() * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(qp->qbuf, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    qp->smem_bytes = quad_smem_bytes<D>(qp->qc);
    auto kern = quad_step_kernel<D, QUAD_WARPS, QUAD_MINB>;
    e = sih_raise_smem_limit(reinterpret_cast<const void *>(kern), qp->smem_bytes);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, QUAD_WARPS * 32, qp->smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) { qp->shape = -1; return cudaSuccess; }
    qp->blocks_per_sm = occ;
    return cudaSuccess;
}

// [counter | start_step[batch]] in the caller's workspace (capi.cu: bind_handover)
inline void quad_bind(QuadPlan *qp, int *ctr, long long batch) {
    qp->n_handover = ctr; qp->start_step = ctr + 1; qp->capacity = batch;
}

inline bool quad_supported(const QuadPlan &qp, const sihmpc_io &) { return qp.shape >= 0; }

inline cudaError_t quad_launch(const QuadPlan &qp, const DevPlan &P, const sihmpc_io &io, long long batch,
                               int n_steps, int sm_count, cudaStream_t st) {
    using D = DistillationDims;
    const long long per_block = QUAD_WARPS * 8;
    long long need = (batch + per_block - 1) / per_block;
    const long long cap = (long long)sm_count * qp.blocks_per_sm;
    int grid = (int)(need < cap ? need : cap);
    if (qp.grid_cap > 0 && qp.grid_cap < grid) grid = qp.grid_cap;      // debug: force many rounds
    // start_step[] = n_steps (nothing to hand over), counter = 0
    quad_step_kernel<D, QUAD_WARPS, QUAD_MINB><<<grid, QUAD_WARPS * 32, qp.smem_bytes, st>>>(
        P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_handover);
    return cudaGetLastError();
}

inline void quad_destroy(QuadPlan *qp) {
    if (qp->qbuf) cudaFree(qp->qbuf);
    memset(qp, 0, sizeof *qp);
    qp->shape = -1;
}

#endif  // SIH_NO_QUAD_HOST

}  // namespace sih
