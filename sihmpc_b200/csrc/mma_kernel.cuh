// mma_kernel.cuh -- fast path v3: quad layout + FP64 TENSOR-CORE products (DMMA.8x8x4).
//
// Same decomposition as dist_kernel.cuh (four lanes per scenario, eight scenarios per warp, the
// Cholesky factor register resident and rows cyclic over the quad), but every product with a
// SHARED controller matrix runs on the FP64 tensor pipe:
//
//      out^T[scenario][row] += vec^T[scenario][k] * M^T[k][row]        mma.sync.m8n8k4.f64
//
//   A fragment (8x4)  : lane (q,g) holds element g of k-tile t of scenario q's vector -- exactly
//                       the quad distribution, so per-scenario vectors are used where they live;
//   B fragment (4x8)  : lane (q,g) holds M[row 8T+q][k-tile t, element g] -- one LDS.64 per lane
//                       per MMA from the constant block staged once per CTA (256 contiguous bytes
//                       per warp: conflict free);
//   C/D fragment (8x8): lane (q,g) receives rows 8T+2g and 8T+2g+1 of scenario q.
// An output fragment pair (T,h) is used directly as k-tile 2T+h of the next product: the constant
// tiles are permuted on the host so that chained products need no data movement at all, and the
// QP variables are eliminated in the permuted order the gradient product delivers them in
// (slot s = 2T+h, lane g <-> variable 8T+2g+h), so the factorisation consumes the fragments as is.
//
// Per step and per warp (8 scenarios) this is ~160 DMMA + ~160 LDS.64 instead of ~2000 DFMA +
// ~1300 LDS.128 for the same products (profiles/r01_v2_quad_kernel_ncu_summary.md: the v2 kernel
// was bound by shared-memory delivery).  bench_micro/dmma_micro.cu: DMMA reaches 37 TFLOP/s at one
// warp per SM sub-partition.
//
// Reference semantics replaced: the same as step_kernel.cuh.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <functional>
#include <string.h>
#include <stdlib.h>
#include "devplan.h"
#include "dist_kernel.cuh"
#include "mma_api.h"

namespace sih {

// NVE_: number of QP variables that are actually eliminated.  Moves whose dead time pushes their whole
// effect past the horizon have a zero column in E, Ht, HJ: their optimum is exactly 0 (presolve on the
// host, mma_dead_variables), so the factorisation runs on the NVE_ live ones only; 0 = all of them.
template <int NX_, int NU_, int NY_, int ND_, int N_, int M_, int NVE_ = 0>
struct MmaDims {
    static constexpr int NX = NX_, NU = NU_, NY = NY_, ND = ND_, N = N_, M = M_;
    static constexpr int NVL = N * NU;                 // logical QP variables (w_opt layout)
    static constexpr int NV = NVE_ > 0 ? NVE_ : NVL;   // live variables = dimension of the factorisation
    static constexpr int NYP = N * NY;
    static constexpr int R = NYP + 2 * NY + ND;
    static constexpr int RE = R + NY;
    static constexpr int SV = NV / 4;                  // factorisation slots (NV % 4 == 0 required)
    static constexpr int TX = (NX + 7) / 8;            // C-layout tiles of x
    static constexpr int TR = (RE + 7) / 8;            // tiles of the residual stack
    static constexpr int TV = (NV + 7) / 8;            // tiles of the gradient / dU
    static constexpr int TY = (NYP + 7) / 8;           // tiles of Y_pred
    static constexpr int TQ = (NYP + 2 * NY + 7) / 8;  // tiles of the squared-residual stack
    static constexpr int TC = (NY + 7) / 8;            // tiles of the measured output
    static constexpr int TM = (M + 7) / 8;             // tiles of the sub-objective values
    static constexpr int KX = 2 * TX;                  // k-tiles of x
    static constexpr int KA = (NY + 1 + 3) / 4;        // k-tiles of [ysp ; 1]
    static constexpr int KP = (M + 3) / 4;             // k-tiles of pesos
    static constexpr int KH = 1;                       // k-tile [1 ; pesos of <= 3 dense Hessian blocks]
    static constexpr int KD = (NU + 3) / 4;            // k-tiles of du
    static constexpr int KE = (3 * NY + 3) / 4;        // k-tiles of [ysp ; syN ; siN] (shifted-plan residuals)
    static constexpr int KHC = 2;                      // k-tiles [1 ; pesos of dense blocks ; pesos of move weights] (bound space)
    static constexpr int HPAIRS = SV * (SV + 1);
    static constexpr int BX = 8 * TX, BV = NVL;
    static constexpr int BRAW = BX + NU + (NU & 1) + BV;
    // stride of the replicated per-scenario buffers == 4 mod 16 doubles: the quads of a half warp then sit on four distinct
    // 8-word bank groups (measured on B200 against == 2 mod 4, 8 and 12 mod 16: +1.8 % on the whole step)
    static constexpr int BS = BRAW + ((4 - BRAW % 16 + 16) % 16);
    static constexpr int WELL = 4;                     // ELL width of the plant map (rows with more non-zeros: no fast path)
    static constexpr int PV = TX * 2 * WELL * 4;       // staged plant values per scenario
    static constexpr int PVS = PV + 4;                 // stride of the staged plant values: == 4 mod 16 doubles, so the eight scenarios of a
                                                       // warp read their 4 consecutive values from 8 distinct bank groups (PV itself is a
                                                       // multiple of 16: every scenario on the same banks, 8-way conflicts)
    static constexpr int WARP_DOUBLES = 8 * BS + 8 * PVS;
    static_assert(NV % 4 == 0, "N*nu must be a multiple of 4 for the MMA fast path");
    static_assert(M <= 8 && NY <= 4 && NU <= 4, "shape outside the MMA fast path");
    // elimination position held by (slot s, lane g) of the factorisation (live[position] = logical variable)
    __host__ __device__ static constexpr int var_of(int s, int g) {
        return (NV % 8 == 4 && s == SV - 1) ? (NV - 4 + g) : (8 * (s / 2) + 2 * g + (s % 2));
    }
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// c[T][.] += sum_t a[t] (x) B[T][t]     B tile element of this lane: cq[(T*KT + t)*32 + lane]
// The k-tiles are dealt round-robin to NS independent accumulator sets (summed at the end) so that
// NT*NS MMA chains are in flight: a single chain is bound by the DMMA latency, not by the pipe.
template <int NT, int KT, int NS_ = 0>
__device__ __forceinline__ void mma_block(const double *__restrict__ cq, const double (&a)[KT],
                                          double (&c)[NT][2], const int lane) {
    constexpr int WANT = (NS_ > 0) ? NS_ : ((NT >= 8) ? 1 : (8 + NT - 1) / NT);
    constexpr int NS = (WANT < KT) ? WANT : KT;
    double p[NS][NT][2];
#pragma unroll
    for (int s = 1; s < NS; ++s)
#pragma unroll
        for (int T = 0; T < NT; ++T) p[s][T][0] = p[s][T][1] = 0.0;
#pragma unroll
    for (int T = 0; T < NT; ++T) { p[0][T][0] = c[T][0]; p[0][T][1] = c[T][1]; }
    double bq[KT][NT];
#pragma unroll
    for (int t = 0; t < KT; ++t)
#pragma unroll
        for (int T = 0; T < NT; ++T) bq[t][T] = cq[(T * KT + t) * 32 + lane];
#pragma unroll
    for (int t = 0; t < KT; ++t)
#pragma unroll
        for (int T = 0; T < NT; ++T) dmma884(p[t % NS][T][0], p[t % NS][T][1], a[t], bq[t][T]);
#pragma unroll
    for (int T = 0; T < NT; ++T) {
        double s0 = p[0][T][0], s1 = p[0][T][1];
#pragma unroll
        for (int s = 1; s < NS; ++s) { s0 += p[s][T][0]; s1 += p[s][T][1]; }
        c[T][0] = s0; c[T][1] = s1;
    }
}

// host: B tiles of  out(T,n) = sum M[orow(T,n)][icol(t,k)] * in(t,k)   (negative index = padding)
inline void mma_pack(std::vector<double> &buf, int off, int NT, int KT,
                     const std::function<int(int, int)> &orow, const std::function<int(int, int)> &icol,
                     const std::function<double(int, int)> &M) {
    for (int T = 0; T < NT; ++T)
        for (int t = 0; t < KT; ++t)
            for (int lane = 0; lane < 32; ++lane) {
                const int n = lane >> 2, k = lane & 3;
                const int r = orow(T, n), c = icol(t, k);
                buf[off + ((size_t)T * KT + t) * 32 + lane] = (r >= 0 && c >= 0) ? M(r, c) : 0.0;
            }
}

// value of factorisation slot sl in a C-layout tile set [TV][2]
template <typename D>
__device__ __forceinline__ double mma_slot(const double (&c)[D::TV][2], const int sl) {
    return (D::NV % 8 == 4 && sl == D::SV - 1) ? c[D::TV - 1][0] : c[sl / 2][sl % 2];
}

// gradient of the step cost in the bound space at the point v (slots):  g = S^T (2 E^T (omega o (e + E S v)) + 2 omega_du o S v),
// S = the map from bound-space variables to moves (difference operator for cumulative moves, identity for the move box)
template <typename D>
__device__ __forceinline__ void mma_bound_gradient(const double *__restrict__ cst, const MmaConst &Q,
                                                   const double (&e)[D::TR][2], const double (&om)[D::TR][2],
                                                   const double (&od)[D::SV], const double (&v)[D::SV],
                                                   double (&gout)[D::SV], double (&dv)[D::SV], const int lane) {
    constexpr int SV = D::SV, TR = D::TR, TV = D::TV;
    double t0[TV][2];
#pragma unroll
    for (int T = 0; T < TV; ++T) t0[T][0] = t0[T][1] = 0.0;
    mma_block<TV, SV>(cst + Q.bdp, v, t0, lane);
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) dv[sl] = mma_slot<D>(t0, sl);
    double r[TR][2];
#pragma unroll
    for (int T = 0; T < TR; ++T) { r[T][0] = e[T][0]; r[T][1] = e[T][1]; }
    mma_block<TR, SV>(cst + Q.br, dv, r, lane);
    double at[2 * TR];
#pragma unroll
    for (int T = 0; T < TR; ++T) { at[2 * T] = om[T][0] * r[T][0]; at[2 * T + 1] = om[T][1] * r[T][1]; }
#pragma unroll
    for (int T = 0; T < TV; ++T) t0[T][0] = t0[T][1] = 0.0;
    mma_block<TV, 2 * TR>(cst + Q.bfc, at, t0, lane);
    double wd[SV];
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) wd[sl] = 2.0 * od[sl] * dv[sl];
    mma_block<TV, SV>(cst + Q.bdpt, wd, t0, lane);
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) gout[sl] = mma_slot<D>(t0, sl);
}

// In-kernel active set for the scenarios of this warp whose unconstrained optimum violates a bound (`need`).
// The QP is restated in the space where the violated family of bounds is SIMPLE (cumulative moves for the input
// box ihmpccontroller.py:420-424, the moves themselves for the move box :42-43), so an active bound is a fixed
// variable: its row and column of the Hessian are replaced by the identity (per-scenario bit masks, the register
// resident factorisation stays static and the eight scenarios of the warp stay in lockstep).  Violated free
// variables are fixed, fixed variables whose multiplier has the wrong sign are released, until the KKT conditions
// hold -- which certifies the optimum of this strictly convex problem whatever path led there; a scenario that
// does not get there in MAXIT rounds returns false and is handed over to the general kernel as before.
// mlo/mhi: bit p = elimination position p fixed at its lower/upper bound (replicated over the quad); kept by the
// caller as the warm start of the next step.  On success fr = the optimal moves in factorisation slots.
template <typename D>
__device__ __forceinline__ bool mma_active_set(const double *__restrict__ cst, const int *__restrict__ cint,
                                               const MmaConst &Q, const double (&e)[D::TR][2],
                                               const double (&om)[D::TR][2], const double (&apes)[D::KP],
                                               const double (&pes)[D::M], const double (&uu)[D::NU],
                                               const int (&vlog)[D::SV], unsigned &mlo, unsigned &mhi, const bool need,
                                               double (&fr)[D::SV], const int lane, const int g, const int mh) {
    constexpr int SV = D::SV, TV = D::TV, NU = D::NU, M = D::M, KP = D::KP, KHC = D::KHC;
    constexpr int MAXIT = 24;
    double blo[SV], bhi[SV], tlo[SV], thi[SV], od[SV];
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) {
        const int j = vlog[sl] % NU;
        const double off = (Q.cspace == 1) ? sel<NU>(uu, j) : 0.0;
        const int bo = (Q.cspace == 1) ? 2 * NU : 0;
        const double lo = cst[Q.bounds + bo + j], hi = cst[Q.bounds + bo + NU + j];
        blo[sl] = lo - off; bhi[sl] = hi - off;
        tlo[sl] = 1e-10 * (1.0 + fabs(lo)); thi[sl] = 1e-10 * (1.0 + fabs(hi));
    }
    {
        double omdu[TV][2];
#pragma unroll
        for (int T = 0; T < TV; ++T) omdu[T][0] = omdu[T][1] = 0.0;
        mma_block<TV, KP>(cst + Q.bomdu, apes, omdu, lane);
#pragma unroll
        for (int sl = 0; sl < SV; ++sl) od[sl] = mma_slot<D>(omdu, sl);
    }
    double ahc[KHC];
#pragma unroll
    for (int t = 0; t < KHC; ++t) {
        const int el = 4 * t + g;                          // 0 -> 1 ; 1+a -> dense block a ; 1+mh+b -> move weight b
        double v = (el == 0) ? 1.0 : 0.0;
        if (el >= 1 && el - 1 < mh) v = sel<M>(pes, cint[Q.hj_idx + el - 1]);
        else if (el >= 1 + mh && el - 1 - mh < Q.n_du) v = sel<M>(pes, cint[Q.du_idx + el - 1 - mh]);
        ahc[t] = v;
    }
    bool done = !need, ok = false;
    double cres[SV];
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) cres[sl] = 0.0;
    if (!need) mlo = mhi = 0u;
    for (int it = 0; it < MAXIT; ++it) {
        if (!__any_sync(0xffffffffu, !done)) break;
        const unsigned mf = mlo | mhi;
        double z[SV], rhs[SV], dv[SV];
#pragma unroll
        for (int sl = 0; sl < SV; ++sl) {
            const int p = 4 * sl + g;
            z[sl] = ((mlo >> p) & 1u) ? blo[sl] : (((mhi >> p) & 1u) ? bhi[sl] : 0.0);
        }
        mma_bound_gradient<D>(cst, Q, e, om, od, z, rhs, dv, lane);
#pragma unroll
        for (int sl = 0; sl < SV; ++sl) rhs[sl] = ((mf >> (4 * sl + g)) & 1u) ? 0.0 : -rhs[sl];
        // Hessian in the bound space, fixed rows/columns replaced by the identity
        double L[SV][D::NV];
        {
            int pp = 0;
#pragma unroll
            for (int t = 0; t < SV; ++t) {
                const bool rowfix = ((mf >> (4 * t + g)) & 1u) != 0u;
#pragma unroll
                for (int jp = 0; jp < (4 * t + 4) / 2; ++jp) {
                    double h[1][2];
                    h[0][0] = h[0][1] = 0.0;
                    mma_block<1, KHC>(cst + Q.bhc + (size_t)pp * KHC * 32, ahc, h, lane);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int j = 2 * jp + hh;
                        const bool fix = rowfix || (((mf >> j) & 1u) != 0u);
                        L[t][j] = fix ? ((j == 4 * t + g) ? 1.0 : 0.0) : h[0][hh];
                    }
                    ++pp;
                }
            }
        }
        int npd = 0;
        quad_solve_spd<D>(L, rhs, g, npd);
        double c[SV];
        unsigned vlo = 0u, vhi = 0u, wrel = 0u;
#pragma unroll
        for (int sl = 0; sl < SV; ++sl) {
            const int p = 4 * sl + g;
            const bool fixed = ((mf >> p) & 1u) != 0u;
            c[sl] = fixed ? z[sl] : rhs[sl];
            if (!fixed && c[sl] < blo[sl] - tlo[sl]) vlo |= 1u << p;
            if (!fixed && c[sl] > bhi[sl] + thi[sl]) vhi |= 1u << p;
        }
        vlo |= __shfl_xor_sync(0xffffffffu, vlo, 1); vlo |= __shfl_xor_sync(0xffffffffu, vlo, 2);
        vhi |= __shfl_xor_sync(0xffffffffu, vhi, 1); vhi |= __shfl_xor_sync(0xffffffffu, vhi, 2);
        // multipliers of the fixed variables = gradient there (only looked at once the point is feasible: bounds
        // are added all at once, released one at a time -- the variable with the worst multiplier)
        if (__any_sync(0xffffffffu, mf != 0u && !done && (vlo | vhi) == 0u)) {
            double gc[SV];
            mma_bound_gradient<D>(cst, Q, e, om, od, c, gc, dv, lane);
            double gs = 0.0;
#pragma unroll
            for (int sl = 0; sl < SV; ++sl) gs = fmax(gs, fabs(gc[sl]));
            gs = 1e-12 * (1.0 + quad_max(gs));
            double worst = 0.0; int wp = -1;
#pragma unroll
            for (int sl = 0; sl < SV; ++sl) {
                const int p = 4 * sl + g;
                const double bad = ((mlo >> p) & 1u) ? -gc[sl] : (((mhi >> p) & 1u) ? gc[sl] : 0.0);
                if (bad > gs && bad > worst) { worst = bad; wp = p; }
            }
            const double wmax = quad_max(worst);
            int cand = (wp >= 0 && worst == wmax) ? wp : 0x7fffffff;     // lowest position among equals
            cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, 1));
            cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, 2));
            if (cand != 0x7fffffff) wrel = 1u << cand;
            if (it < MAXIT / 3) {                  // early rounds: all of them at once (few rounds when it works)
#pragma unroll
                for (int sl = 0; sl < SV; ++sl) {
                    const int p = 4 * sl + g;
                    if (((mlo >> p) & 1u) && gc[sl] < -gs) wrel |= 1u << p;
                    if (((mhi >> p) & 1u) && gc[sl] > gs) wrel |= 1u << p;
                }
                wrel |= __shfl_xor_sync(0xffffffffu, wrel, 1); wrel |= __shfl_xor_sync(0xffffffffu, wrel, 2);
            }
        }
        npd |= __shfl_xor_sync(0xffffffffu, npd, 1); npd |= __shfl_xor_sync(0xffffffffu, npd, 2);
        if (!done) {
            if ((vlo | vhi | wrel) == 0u && !npd) {
                done = true; ok = true;
#pragma unroll
                for (int sl = 0; sl < SV; ++sl) cres[sl] = c[sl];
            } else if ((vlo | vhi) != 0u) {
                mlo |= vlo; mhi |= vhi;
            } else {
                mlo &= ~wrel; mhi &= ~wrel;
            }
        }
    }
    {   // moves of the accepted point
        double t0[TV][2];
#pragma unroll
        for (int T = 0; T < TV; ++T) t0[T][0] = t0[T][1] = 0.0;
        mma_block<TV, SV>(cst + Q.bdp, cres, t0, lane);
        if (need && ok) {
#pragma unroll
            for (int sl = 0; sl < SV; ++sl) fr[sl] = mma_slot<D>(t0, sl);
        }
    }
    return ok;
}

// Sub-objective values of the shifted plan at the state xr and the satisficing weights from them
// (satWeights2 / satWeights, ihmpccontroller.py:619-636,607-617).  fr: the moves of the last solve in factorisation
// slots (UNshifted: the shift is folded into the constants), syv = [syN ; siN ; siN_next], ysp the set-point that
// solve was made for.
template <typename D>
__device__ __forceinline__ void mma_weight_update(const double *__restrict__ cst, const MmaConst &Q,
                                                  const double (&xr)[D::TX][2], const double (&fr)[D::SV],
                                                  const double (&ysp)[D::NY], const double (&syv)[3 * D::NY],
                                                  double (&pes)[D::M], const int flags, const double alfa,
                                                  const int lane, const int g) {
    constexpr int NY = D::NY, M = D::M, SV = D::SV, TX = D::TX, TQ = D::TQ, TM = D::TM, KX = D::KX;
    // residual stack of the shifted plan in ONE product: rows k*ny+i = Y'_k,i - ysp_i - syN_i - tau_k siN_i, then
    // syN and siN themselves -- the set-point and the slacks ride along as two more k-tiles (their
    // coefficients -1, -tau_k, 1 sit in the constant block) instead of a dozen selects per row
    double yq[TQ][2];
#pragma unroll
    for (int T = 0; T < TQ; ++T) yq[T][0] = yq[T][1] = 0.0;
    {
        double ax[KX + SV + D::KE];
#pragma unroll
        for (int T = 0; T < TX; ++T) { ax[2 * T] = xr[T][0]; ax[2 * T + 1] = xr[T][1]; }
#pragma unroll
        for (int sl = 0; sl < SV; ++sl) ax[KX + sl] = fr[sl];     // shift folded into the constants
#pragma unroll
        for (int t = 0; t < D::KE; ++t) {                         // [ysp ; syN ; siN(next)], element 4t+g
            const int e = 4 * t + g;
            ax[KX + SV + t] = (e < NY) ? sel<NY>(ysp, e) : ((e < 3 * NY) ? sel<3 * NY>(syv, e < 2 * NY ? e - NY : e) : 0.0);
        }
        mma_block<TQ, KX + SV + D::KE>(cst + Q.byp, ax, yq, lane);
    }
    double asq[2 * TQ + SV];
#pragma unroll
    for (int T = 0; T < TQ; ++T) { asq[2 * T] = yq[T][0] * yq[T][0]; asq[2 * T + 1] = yq[T][1] * yq[T][1]; }
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) asq[2 * TQ + sl] = fr[sl] * fr[sl];
    double vc[TM][2];
#pragma unroll
    for (int T = 0; T < TM; ++T) vc[T][0] = vc[T][1] = 0.0;
    mma_block<TM, 2 * TQ + SV>(cst + Q.bvn, asq, vc, lane);
    double vn[M];
#pragma unroll
    for (int i = 0; i < M; ++i) vn[i] = quad_bcast(vc[i / 8][i % 2], (i % 8) / 2);
    if (flags & SIHMPC_F_WEIGHTS_SAT2) {
        double smax = -SIH_INF;
#pragma unroll
        for (int i = 0; i < M; ++i) smax = fmax(smax, vn[i] * cst[Q.inv_gamma + i]);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const double g0 = cst[Q.gamma + i];
            const double g1 = fmax(g0, g0 * smax + 1e-6);
            pes[i] = alfa * pes[i] + (1.0 - alfa) * fast_rcp(g1 - vn[i]);     // g1 - vn >= 1e-6 > 0
        }
    } else {
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const double g0 = cst[Q.gamma + i];
            pes[i] = fast_rcp(g0 - fmin(fmax(vn[i], 0.0), 0.99 * g0));
        }
    }
}

// loads of per-scenario state that ANOTHER kernel may have written while this one runs (the streaming resume workers
// read what the first pass hands over): around L1, which is not coherent across SMs
template <bool CG>
__device__ __forceinline__ double mma_gload(const double *p) { return CG ? __ldcg(p) : *p; }
__device__ __forceinline__ int mma_vload(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

// MODE 0: first pass.  Unconstrained optimum + bound check; a scenario that violates a bound is appended to `list`
//         and leaves the kernel (the active set inlined here cost the hot loop 20 %: register pressure).
// MODE 1: resume pass over the scenarios of `list` (n_list = their counter), each from its own start_step, with the
//         in-kernel active set (mma_active_set); eight listed scenarios per warp, every quad on its own step index.
//         What it cannot resolve is flagged in start_step / n_handover for the general kernel (third pass).
// MODE 2: first pass with the active set inline (workloads where most scenarios sit on bounds).
// MODE 4: STREAMING resume workers.  A handful of CTAs launched WITH the first pass (second stream): every warp claims
//         the next entry of the hand-over list (atomic), waits until the first pass has published it (the entry is
//         written after the scenario's state and a fence) or has finished, and steps that scenario like MODE 1 -- the
//         resume pass is a latency chain of a few dozen scenarios, this overlaps all of it but the last entries.
// MODE 3: first pass of CONTROLLER-RESIDENT stepping (sihmpc_resident_step, one step per launch, no plant): the
//         weights are brought up to date FIRST, at the measured state with the plan of the previous solve
//         (res_wstart / io.w_opt / res_ysp_prev), then the step is MODE 0.
template <typename D, int WARPS, int MINB, int MODE>
__global__ void __launch_bounds__(WARPS * 32, MINB)
mma_step_kernel(const DevPlan P, const MmaConst Q, const double *__restrict__ qbuf, const sihmpc_io io,
                const long long batch, const int n_steps, int *__restrict__ start_step,
                int *__restrict__ n_handover, int *__restrict__ n_inkernel, int *__restrict__ list,
                const int *__restrict__ n_list, const double *__restrict__ res_wstart = nullptr,
                const double *__restrict__ res_ysp_prev = nullptr, const int res_flags = 0,
                int *__restrict__ sync_ctr = nullptr, const int n_first_ctas = 0) {
    // sync_ctr[0]: CTAs of the first pass that have finished, sync_ctr[1]: list entries claimed by the workers
    constexpr bool STREAM = (MODE == 4);
    constexpr bool AS = (MODE == 1 || MODE == 2 || MODE == 4), RESUME = (MODE == 1 || MODE == 4), RESIDENT = (MODE == 3);
    long long n_work = batch;
    // resume pass over a SHORT list: one listed scenario per warp instead of eight -- the quads of a warp iterate in
    // lockstep, so a warp of eight constrained scenarios pays the rounds of the slowest one at every step; alone in
    // its warp a scenario pays its own (the pass is a latency tail of a few dozen scenarios, not throughput)
    bool spread = false;
    if constexpr (RESUME && !STREAM) {
        n_work = (long long)(*n_list) - 0x7f7f7f7f;
        spread = n_work <= (long long)gridDim.x * WARPS;
        if ((long long)blockIdx.x * WARPS * (spread ? 1 : 8) >= n_work) return;
    }
    if constexpr (STREAM) spread = true;
    constexpr int NU = D::NU, NY = D::NY, M = D::M, NV = D::NV, NYP = D::NYP, R = D::R;
    const int nxr = P.nx;                     // states of THIS model (<= D::NX = 8 TX: the compiled capacity)
    constexpr int SV = D::SV, TX = D::TX, TR = D::TR, TV = D::TV, TY = D::TY, TQ = D::TQ, TC = D::TC, TM = D::TM;
    constexpr int KX = D::KX, KA = D::KA, KP = D::KP, KH = D::KH, KD = D::KD;
    constexpr int BS = D::BS;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane & 3, q = lane >> 2;
    double *cst = smem;
    {   // constant block -> shared memory, 16 bytes per load and eight loads in flight per thread: with one 8-byte load per
        // iteration this staging was ~25 us per CTA -- two thirds of a ONE-step launch (batch-1 latency, the chunks of
        // sihmpc_resident_step); it is amortised over the steps of a longer launch
        const double2 *src2 = reinterpret_cast<const double2 *>(qbuf);
        double2 *dst2 = reinterpret_cast<double2 *>(cst);
        const int n2 = Q.total_d >> 1;
        for (int i0 = threadIdx.x; i0 < n2; i0 += 8 * blockDim.x) {
            double2 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = i0 + k * blockDim.x; if (i < n2) v[k] = src2[i]; }
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = i0 + k * blockDim.x; if (i < n2) dst2[i] = v[k]; }
        }
        if ((Q.total_d & 1) && threadIdx.x == 0) cst[Q.total_d - 1] = qbuf[Q.total_d - 1];
    }
    int *cint = reinterpret_cast<int *>(cst + Q.total_d);
    {
        const int *src = reinterpret_cast<const int *>(qbuf + Q.total_d);
        for (int i = threadIdx.x; i < Q.total_i; i += blockDim.x) cint[i] = src[i];
    }
    __syncthreads();
    double *wbase = cst + Q.total_d + (((Q.total_i + 3) >> 2) << 1) + (size_t)wib * D::WARP_DOUBLES;
    double *bx = wbase + q * BS;              // replicated [x (8*TX) ; du (NU)] of this scenario
    double *bv = bx + D::BX + NU + (NU & 1);  // replicated dU (logical order)
    double *pval = wbase + 8 * BS + q * D::PVS;   // this scenario's plant values in ELL order

    int vlog[SV];                             // logical variable of (slot, this lane)
#pragma unroll
    for (int sl = 0; sl < SV; ++sl) vlog[sl] = cint[Q.vlive + D::var_of(sl, g)];
    for (int e = g; e < D::NVL; e += 4) bv[e] = 0.0;      // presolved moves stay exactly zero
    unsigned jpack = 0u;                      // input index of the live move behind every C-layout element of this lane (2 bits each)
#pragma unroll
    for (int T = 0; T < TV; ++T)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int a = (NV % 8 == 4 && T == TV - 1) ? (h == 0 ? NV - 4 + g : -1) : 8 * T + 2 * g + h;
            if (a >= 0 && a < NV) jpack |= (unsigned)(cint[Q.vlive + a] % NU) << (2 * (2 * T + h));
        }
    __syncwarp();
    const long long gq = spread ? ((q == 0) ? (long long)blockIdx.x * WARPS + wib : n_work)
                                : ((long long)blockIdx.x * WARPS + wib) * 8 + q;
    const long long nq = (long long)gridDim.x * WARPS * 8;
    const long long rounds = STREAM ? (1LL << 40) : (spread ? 1 : (n_work + nq - 1) / nq);
    const int nstat = NY + SIHMPC_NSTATS_EXTRA;
    const bool advance = (io.flags & SIHMPC_F_ADVANCE) != 0;
    const int mh = P.mh;

    for (long long rd = 0; rd < rounds; ++rd) {
        long long s = gq + rd * nq;
        bool valid = s < n_work;
        int st0 = 0, n_iter = n_steps;
        if constexpr (STREAM) {
            // claim the next list entry; wait (lane 0 polls, around L1) until it is published or the first pass is over
            int got = -1;
            if (lane == 0) {
                const int idx = atomicAdd(sync_ctr + 1, 1) - 0x7f7f7f7f;
                unsigned long long t_start, t_now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
                for (;;) {                                                           // bounded (0.5 s): a lost worker only delays
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                    if (t_now - t_start > 500000000ull) break;
                    const int npub = mma_vload(n_list) - 0x7f7f7f7f;
                    if (idx < npub) {
                        const int e = mma_vload(list + idx);
                        if (e != 0x7f7f7f7f) { got = e; break; }
                    } else if (mma_vload(sync_ctr) - 0x7f7f7f7f >= n_first_ctas) {
                        if (idx >= mma_vload(n_list) - 0x7f7f7f7f) break;             // nothing more will come
                        continue;
                    }
                    __nanosleep(256);
                }
            }
            got = __shfl_sync(0xffffffffu, got, 0);
            if (got < 0) break;                                                        // uniform
            __threadfence();
            s = got;
            valid = (q == 0);
            st0 = valid ? mma_vload(start_step + s) : n_steps;
            if (st0 < 0 || st0 > n_steps) st0 = n_steps;
            n_iter = __reduce_max_sync(0xffffffffu, n_steps - st0);
            __syncwarp();
            if (valid && g == 0) start_step[s] = 0x7f7f7f7f;    // resolved here unless flagged again below
        } else if constexpr (RESUME) {
            s = valid ? list[s] : 0;
            st0 = valid ? start_step[s] : n_steps;
            if (st0 < 0 || st0 > n_steps) st0 = n_steps;
            n_iter = __reduce_max_sync(0xffffffffu, n_steps - st0);
            __syncwarp();
            if (valid && g == 0) start_step[s] = 0x7f7f7f7f;    // resolved here unless flagged again below
        }
        const long long sc = valid ? s : 0;
        bool alive = valid;
        // ---- plant values of this scenario (nominal or per-scenario), staged once per round
        __syncwarp();
        {
            const double *pv = io.plant_val ? io.plant_val + (size_t)sc * P.pl_nnz : nullptr;
            for (int e = g; e < D::PV; e += 4) {
                const int ci = cint[Q.ell_csr + e];
                pval[e] = (pv && ci >= 0) ? pv[ci] : cst[Q.ell_val + e];
            }
        }
        // ---- set-point schedule and disturbance step in registers (no global loads per step)
        int seg = 0, next_start = 0x7fffffff, dstep = -1;
        if (io.ysp_seq) {
            for (int a = 1; a < io.n_seg; ++a)
                if (io.seg_start[sc * io.n_seg + a] <= io.k0) seg = a;
            if (seg + 1 < io.n_seg) next_start = io.seg_start[sc * io.n_seg + seg + 1];
        }
        if (io.dist_step) dstep = io.dist_step[sc];
        double ysp[NY];
#pragma unroll
        for (int i = 0; i < NY; ++i) ysp[i] = io.ysp_seq ? io.ysp_seq[(sc * io.n_seg + seg) * NY + i] : io.ysp[sc * NY + i];
        // ---- state: x in C-layout registers (rows 8T+2g+h), replicated copy in smem for the plant
        double xr[TX][2];
#pragma unroll
        for (int T = 0; T < TX; ++T)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 8 * T + 2 * g + h;
                xr[T][h] = (r < nxr) ? mma_gload<STREAM>(io.x + sc * nxr + r) : 0.0;
            }
        double uu[NU], pes[M], ise[2];
        double st_du2 = 0.0, st_viol = 0.0, st_J = 0.0;
        int st_or = 0;
        unsigned mlo = 0u, mhi = 0u;         // warm start of the in-kernel active set (bound masks of the last step)
#pragma unroll
        for (int j = 0; j < NU; ++j) uu[j] = mma_gload<STREAM>(io.u + sc * NU + j);
#pragma unroll
        for (int i = 0; i < M; ++i) pes[i] = mma_gload<STREAM>(io.pesos + sc * M + i);
        ise[0] = ise[1] = 0.0;
        if (io.stats) {
#pragma unroll
            for (int h = 0; h < 2; ++h) ise[h] = (2 * g + h < NY) ? mma_gload<STREAM>(io.stats + sc * nstat + 2 * g + h) : 0.0;
            st_du2 = mma_gload<STREAM>(io.stats + sc * nstat + NY); st_viol = mma_gload<STREAM>(io.stats + sc * nstat + NY + 1);
            st_J = mma_gload<STREAM>(io.stats + sc * nstat + NY + 2); st_or = (int)mma_gload<STREAM>(io.stats + sc * nstat + NY + 3);
        }

        // state of an alive scenario back to global memory (end of the launch, or hand-over)
        auto write_state = [&]() {
#pragma unroll
            for (int T = 0; T < TX; ++T)
#pragma unroll
                for (int h = 0; h < 2; ++h) { const int r = 8 * T + 2 * g + h; if (r < nxr) io.x[s * nxr + r] = xr[T][h]; }
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < NU; ++j) io.u[s * NU + j] = uu[j];
#pragma unroll
                for (int i = 0; i < M; ++i) io.pesos[s * M + i] = pes[i];
            }
            if (io.stats) {
#pragma unroll
                for (int h = 0; h < 2; ++h) if (2 * g + h < NY) io.stats[s * nstat + 2 * g + h] = ise[h];
                if (g == 0) {
                    io.stats[s * nstat + NY] = st_du2; io.stats[s * nstat + NY + 1] = st_viol;
                    io.stats[s * nstat + NY + 2] = st_J; io.stats[s * nstat + NY + 3] = (double)st_or;
                }
            }
        };

        if constexpr (RESIDENT) {
            if (res_flags & (SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1)) {
                // previous solve of this scenario: moves (unshifted, io.w_opt of the last call), syN, siN_next, set-point
                constexpr int nwl = D::NVL + 2 * NY;
                double frp[SV], syp[3 * NY], yspp[NY];
#pragma unroll
                for (int sl = 0; sl < SV; ++sl) frp[sl] = io.w_opt[sc * nwl + vlog[sl]];
#pragma unroll
                for (int i = 0; i < NY; ++i) {
                    syp[i] = io.w_opt[sc * nwl + D::NVL + i];
                    syp[NY + i] = 0.0;
                    syp[2 * NY + i] = res_wstart[sc * nwl + D::NVL + NY + i];
                    yspp[i] = res_ysp_prev[sc * NY + i];
                }
                mma_weight_update<D>(cst, Q, xr, frp, yspp, syp, pes, res_flags, io.alfa, lane, g);
            }
        }
        for (int it = 0; it < n_iter; ++it) {
            const int step = st0 + it;
            if constexpr (RESUME) { if (step >= n_steps) alive = false; }
            const int kabs = io.k0 + step;
            if (kabs >= next_start) {                    // rare: the schedule switches segment
                while (seg + 1 < io.n_seg && io.seg_start[sc * io.n_seg + seg + 1] <= kabs) ++seg;
                next_start = (seg + 1 < io.n_seg) ? io.seg_start[sc * io.n_seg + seg + 1] : 0x7fffffff;
#pragma unroll
                for (int i = 0; i < NY; ++i) ysp[i] = io.ysp_seq[(sc * io.n_seg + seg) * NY + i];
            }
            // A-operand views of the small replicated vectors: element g of [ysp;1], [pesos;1], [1;pw]
            double aux[KA], apes[KP], ahp[KH];
#pragma unroll
            for (int t = 0; t < KA; ++t) {
                const int e = 4 * t + g;
                aux[t] = (e < NY) ? sel<NY>(ysp, e) : (e == NY ? 1.0 : 0.0);
            }
#pragma unroll
            for (int t = 0; t < KP; ++t) {
                const int e = 4 * t + g;
                apes[t] = (e < M) ? sel<M>(pes, e) : 0.0;
            }
#pragma unroll
            for (int t = 0; t < KH; ++t) {
                const int e = 4 * t + g;                       // 0 -> 1.0 ; 1+a -> pesos[hj_index[a]]
                const int a = (e >= 1 && e - 1 < M) ? e - 1 : 0;
                const double pa = sel<M>(pes, cint[Q.hj_idx + a]);       // unconditional load + selects: no branch in the step loop
                ahp[t] = sel_f64(e == 0, 1.0, sel_f64(e - 1 < mh, pa, 0.0));
            }
            // ---------------- K1: e = Px x + Py ysp + e0 ; omega = Wrow pesos + w0row
            double er[TR][2], om[TR][2];
#pragma unroll
            for (int T = 0; T < TR; ++T) {
                er[T][0] = er[T][1] = 0.0;
                const double2 w0 = reinterpret_cast<const double2 *>(cst + Q.w0c)[T * 4 + g];
                om[T][0] = w0.x; om[T][1] = w0.y;
            }
            {
                double ax[KX + KA];
#pragma unroll
                for (int T = 0; T < TX; ++T) { ax[2 * T] = xr[T][0]; ax[2 * T + 1] = xr[T][1]; }
#pragma unroll
                for (int t = 0; t < KA; ++t) ax[KX + t] = aux[t];
                mma_block<TR, KX + KA>(cst + Q.be, ax, er, lane);
            }
            mma_block<TR, KP>(cst + Q.bw, apes, om, lane);
            // ---------------- gradient: fr = -(2 E^T (omega o e)), delivered in factorisation slots
            double fr[SV];
            {
                double at[2 * TR];
#pragma unroll
                for (int T = 0; T < TR; ++T) { at[2 * T] = om[T][0] * er[T][0]; at[2 * T + 1] = om[T][1] * er[T][1]; }
                double fc[TV][2];
#pragma unroll
                for (int T = 0; T < TV; ++T) fc[T][0] = fc[T][1] = 0.0;
                mma_block<TV, 2 * TR>(cst + Q.bf, at, fc, lane);
#pragma unroll
                for (int sl = 0; sl < SV; ++sl)
                    fr[sl] = -((NV % 8 == 4 && sl == SV - 1) ? fc[TV - 1][0] : fc[sl / 2][sl % 2]);
            }
            // ---------------- Hessian rows: H = Ht + sum pw_a HJ_a (tensor pipe), + 2 omega_du on the diagonal
            double L[SV][NV];
            {
                double omdu[TV][2];
#pragma unroll
                for (int T = 0; T < TV; ++T) omdu[T][0] = omdu[T][1] = 0.0;
                mma_block<TV, KP>(cst + Q.bomdu, apes, omdu, lane);
                int pp = 0;
#pragma unroll
                for (int t = 0; t < SV; ++t) {
                    const double o = 2.0 * ((NV % 8 == 4 && t == SV - 1) ? omdu[TV - 1][0] : omdu[t / 2][t % 2]);
                    const double dlo = (g & 1) ? 0.0 : o, dhi = (g & 1) ? o : 0.0;
#pragma unroll
                    for (int jp = 0; jp < (4 * t + 4) / 2; ++jp) {
                        double h[1][2];
                        h[0][0] = h[0][1] = 0.0;
                        mma_block<1, KH>(cst + Q.bh + (size_t)pp * KH * 32, ahp, h, lane);
                        if (jp == 2 * t || jp == 2 * t + 1) {
                            const bool mine = (jp - 2 * t) == (g >> 1);
                            h[0][0] += mine ? dlo : 0.0;
                            h[0][1] += mine ? dhi : 0.0;
                        }
                        L[t][2 * jp] = h[0][0];
                        L[t][2 * jp + 1] = h[0][1];
                        ++pp;
                    }
                }
            }
            // ---------------- K2: factorisation + substitutions (registers, quad shuffles)
            int notpd = 0;
            quad_solve_spd<D>(L, fr, g, notpd);
            // replicated dU (logical order) for the bound check, the outputs and the plant
#pragma unroll
            for (int sl = 0; sl < SV; ++sl) bv[vlog[sl]] = fr[sl];
            __syncwarp();
            // ---------------- feasibility of the unconstrained optimum: viol = max over rows of (violation - 1e-10 (1 + |bound|)),
            // a bound is violated when viol > 0 (the relative test of the general kernel without its divisions)
            double viol = 0.0, viol_abs = 0.0;
            if (P.has_box) {
#pragma unroll
                for (int sl = 0; sl < SV; ++sl) {
                    const int j = vlog[sl] % NU;
                    const double lo = cst[Q.bounds + j], hi = cst[Q.bounds + NU + j];
                    if (lo > -SIH_INF) { const double a = lo - fr[sl]; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(lo), a)); }
                    if (hi < SIH_INF) { const double a = fr[sl] - hi; viol_abs = fmax(viol_abs, a); viol = fmax(viol, fma(-1e-10, 1.0 + fabs(hi), a)); }
                }
            }
            if (P.has_cum) {
                // cumulative moves at the live positions as ONE product on the tensor pipe (every lane gets 2 TV of
                // them), an infinite bound drops out by itself (lo - c = -inf); round 1-2: a serial prefix-sum loop over
                // the horizon on NU of the four lanes (8 % of the step's instructions, divergent)
                double cm[TV][2];
#pragma unroll
                for (int T = 0; T < TV; ++T) cm[T][0] = cm[T][1] = 0.0;
                mma_block<TV, SV>(cst + Q.bcum, fr, cm, lane);
                double lo_[NU], hi_[NU];
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    lo_[j] = cst[Q.bounds + 2 * NU + j] - uu[j]; hi_[j] = cst[Q.bounds + 3 * NU + j] - uu[j];
                }
#pragma unroll
                for (int T = 0; T < TV; ++T)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (NV % 8 == 4 && T == TV - 1 && h == 1) continue;          // padding row of the half tile
                        const int j = (jpack >> (2 * (2 * T + h))) & 3;
                        const double lo = sel<NU>(lo_, j), hi = sel<NU>(hi_, j);
                        const double a = lo - cm[T][h], b = cm[T][h] - hi;
                        viol_abs = fmax(viol_abs, fmax(a, b));
                        viol = fmax(viol, fmax(fma(-1e-10, 1.0 + fabs(lo), a), fma(-1e-10, 1.0 + fabs(hi), b)));
                    }
                if (Q.cum0) {                  // uniform, rare: the first move of an input is presolved -> c = 0 is in the set
#pragma unroll
                    for (int j = 0; j < NU; ++j)
                        if ((Q.cum0 >> j) & 1) {
                            const double a = lo_[j], b = -hi_[j];
                            viol_abs = fmax(viol_abs, fmax(a, b));
                            viol = fmax(viol, fmax(fma(-1e-10, 1.0 + fabs(lo_[j]), a), fma(-1e-10, 1.0 + fabs(hi_[j]), b)));
                        }
                }
            }
            viol = quad_max(viol);
            viol_abs = quad_max(viol_abs);
            int nact = 0;
            if constexpr (AS) {
                if (Q.cspace >= 0) {           // bound violated: active set on the register-resident problem
                    const bool need = alive && !notpd && viol > 0.0;
                    if (__any_sync(0xffffffffu, need)) {
                        const bool ok = mma_active_set<D>(cst, cint, Q, er, om, apes, pes, uu, vlog, mlo, mhi, need, fr,
                                                          lane, g, mh);
                        __syncwarp();
                        if (need && ok) {
#pragma unroll
                            for (int sl = 0; sl < SV; ++sl) bv[vlog[sl]] = fr[sl];
                            viol = 0.0; viol_abs = 0.0;
                            nact = __popc(mlo | mhi);
                            if (g == 0) atomicAdd(n_inkernel, 1);
                        }
                        __syncwarp();
                    } else {
                        mlo = mhi = 0u;
                    }
                }
            }
            if (alive && (viol > 0.0 || notpd)) {
                alive = false;             // hand over from this step on (MODE 0: resume pass; else: general kernel)
                mlo = mhi = 0u;
                // the state first, then (fence) the step it stopped at, then (fence) the list entry: a streaming worker
                // that sees the entry sees everything behind it
                write_state();
                __threadfence();
                __syncwarp(0xFu << (lane & ~3));            // THIS quad only: the branch is per quad (never a full-warp barrier here)
                if (g == 0) {
                    start_step[s] = step;
                    __threadfence();
                    const int slot = atomicAdd(n_handover, 1) - 0x7f7f7f7f;
                    if constexpr (MODE == 0 || MODE == 3) atomicExch(list + slot, (int)s);
                }
            }
            // ---------------- K3: r = e + E dU, cost, slacks
            double syv[3 * NY];
            double Jval;
            {
                mma_block<TR, SV>(cst + Q.br, fr, er, lane);        // er now holds r (rows >= R: xi(N+1))
                double ja = 0.0;
#pragma unroll
                for (int T = 0; T < TR; ++T) { ja = fma(om[T][0], er[T][0] * er[T][0], ja); ja = fma(om[T][1], er[T][1] * er[T][1], ja); }
                double omdu[TV][2];
#pragma unroll
                for (int T = 0; T < TV; ++T) omdu[T][0] = omdu[T][1] = 0.0;
                mma_block<TV, KP>(cst + Q.bomdu, apes, omdu, lane);
#pragma unroll
                for (int sl = 0; sl < SV; ++sl) {
                    const double o = (NV % 8 == 4 && sl == SV - 1) ? omdu[TV - 1][0] : omdu[sl / 2][sl % 2];
                    ja = fma(o, fr[sl] * fr[sl], ja);
                }
                Jval = quad_sum(ja);
#pragma unroll
                for (int i = 0; i < NY; ++i) {
                    syv[i] = quad_bcast(er[(NYP + i) / 8][(NYP + i) % 2], ((NYP + i) % 8) / 2);
                    syv[NY + i] = quad_bcast(er[(NYP + NY + i) / 8][(NYP + NY + i) % 2], ((NYP + NY + i) % 8) / 2);
                    syv[2 * NY + i] = quad_bcast(er[(R + i) / 8][(R + i) % 2], ((R + i) % 8) / 2);
                }
            }
            const bool last_step = (step == n_steps - 1);
            // uniform guard first: a step in the middle of a launch without trajectories writes nothing
            const bool out_step = last_step || io.traj_du || io.traj_J ||
                                  ((io.flags & SIHMPC_F_UPDATE_VIN) && io.vinant && !advance);
            if (out_step && alive) {
                if (last_step) {
                    if (io.du) for (int j = g; j < NU; j += 4) io.du[s * NU + j] = bv[j];
                    if (io.w_opt) {
                        constexpr int nw = D::NVL + 2 * NY;
#pragma unroll
                        for (int sl = 0; sl < SV; ++sl) io.w_opt[s * nw + vlog[sl]] = fr[sl];
                        for (int i = g; i < D::NVL - NV; i += 4) io.w_opt[s * nw + cint[Q.vdead + i]] = 0.0;
                        if (g == 0) {
#pragma unroll
                            for (int i = 0; i < 2 * NY; ++i) io.w_opt[s * nw + D::NVL + i] = syv[i];
                        }
                    }
                    if (g == 0) {
                        if (io.J) io.J[s] = Jval;
                        if (io.status) io.status[s] = 0;
                        if (io.n_active) io.n_active[s] = nact;
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
            // ---------------- advance: plant x+ = [A|B][x;du] (ELL, optional per-scenario values)
#pragma unroll
            for (int T = 0; T < TX; ++T) { bx[8 * T + 2 * g] = xr[T][0]; bx[8 * T + 2 * g + 1] = xr[T][1]; }
            for (int j = g; j < NU; j += 4) bx[D::BX + j] = bv[j];
            __syncwarp();
            {
                const int *ecol = cint + Q.ell_col + g;
                const double *eval = pval + g;
#pragma unroll
                for (int T = 0; T < TX; ++T)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int w = 0; w < D::WELL; w += 2) {
                            const int e = ((T * 2 + h) * D::WELL + w) * 4;
                            a0 = fma(eval[e], bx[ecol[e]], a0);
                            a1 = fma(eval[e + 4], bx[ecol[e + 4]], a1);
                        }
                        xr[T][h] = a0 + a1;
                    }
            }
            double du2 = 0.0;
            double adu[KD];
#pragma unroll
            for (int j = 0; j < NU; ++j) { const double d = bv[j]; uu[j] += d; du2 = fma(d, d, du2); }
#pragma unroll
            for (int t = 0; t < KD; ++t) { const int e = 4 * t + g; adu[t] = (e < NU) ? bv[e < NU ? e : 0] : 0.0; }
            if (kabs == dstep) {
#pragma unroll
                for (int h = 0; h < 2; ++h) if (2 * g + h < NY) xr[0][h] += io.dist[sc * NY + 2 * g + h];
            }
            __syncwarp();          // all reads of bx/bv of this step are done
            // measured output y = C x+ + D du, closed-loop statistics
            {
                double ax[KX + KD];
#pragma unroll
                for (int T = 0; T < TX; ++T) { ax[2 * T] = xr[T][0]; ax[2 * T + 1] = xr[T][1]; }
#pragma unroll
                for (int t = 0; t < KD; ++t) ax[KX + t] = adu[t];
                double yo[TC][2];
#pragma unroll
                for (int T = 0; T < TC; ++T) yo[T][0] = yo[T][1] = 0.0;
                mma_block<TC, KX + KD>(cst + Q.by, ax, yo, lane);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = 2 * g + h;
                    if (i < NY) {
                        const double err = yo[0][h] - sel<NY>(ysp, i);
                        ise[h] = fma(err, err, ise[h]);
                        if (alive && io.traj_y) io.traj_y[((long long)step * batch + s) * NY + i] = yo[0][h];
                    }
                }
                st_du2 += du2; st_viol = fmax(st_viol, viol_abs); st_J = Jval;
            }
            // ---------------- sub-objective values at the new state, weight update
            if (io.flags & (SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1))
                mma_weight_update<D>(cst, Q, xr, fr, ysp, syv, pes, io.flags, io.alfa, lane, g);
            if constexpr (RESUME) {
                if (alive && last_step) { write_state(); alive = false; }
            }
        }  // steps

        __syncwarp();
        if (!RESUME && alive && advance) {
            write_state();
        }
        if constexpr (RESIDENT) {          // no advance in this mode: the weights brought up to date above still go back
            if (alive && !advance && g == 0) {
#pragma unroll
                for (int i = 0; i < M; ++i) io.pesos[s * M + i] = pes[i];
            }
        }
    }
    if constexpr (MODE == 0) {
        if (sync_ctr) {                    // this CTA of the first pass is done: the workers stop waiting once all are
            __syncthreads();
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(sync_ctr, 1); }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------
#ifndef SIH_MMA_WARPS
#define SIH_MMA_WARPS 4
#endif
constexpr int MMA_WARPS = SIH_MMA_WARPS;
#ifndef SIH_MMA_MINB
#define SIH_MMA_MINB 2
#endif
constexpr int MMA_MINB = SIH_MMA_MINB;

// presolve: variable v is dead when nothing but its own move weight depends on it (zero column in E, Mi1u,
// Ht, HJ), that weight is positive and 0 lies inside its move box: dU_v = 0 exactly
inline std::vector<int> mma_dead_variables(const DevPlan &P, const sihmpc_plan_desc *d) {
    std::vector<int> dead;
    const int nv = P.nv, R = d->R, m = d->m, ny = d->ny, nu = d->nu;
    for (int v = 0; v < nv; ++v) {
        bool z = true;
        for (int r = 0; r < R && z; ++r) z = (d->E[(size_t)r * nv + v] == 0.0);
        for (int r = 0; r < ny && z; ++r) z = (d->Mi1u[(size_t)r * nv + v] == 0.0);
        for (int c = 0; c < nv && z; ++c) z = (d->Ht[(size_t)v * nv + c] == 0.0 && d->Ht[(size_t)c * nv + v] == 0.0);
        for (int a = 0; a < d->mh && z; ++a)
            for (int c = 0; c < nv && z; ++c)
                z = (d->HJ[((size_t)a * nv + v) * nv + c] == 0.0 && d->HJ[((size_t)a * nv + c) * nv + v] == 0.0);
        bool w = false;
        for (int i = 0; i < m; ++i) { if (d->Wdu[(size_t)v * m + i] > 0.0) w = true; if (d->Wdu[(size_t)v * m + i] < 0.0) z = false; }
        const int j = v % nu;
        if (z && w && d->dulb[j] <= 0.0 && d->duub[j] >= 0.0) dead.push_back(v);
    }
    return dead;
}

template <typename D>
inline void mma_build_constants(const sihmpc_plan_desc *d, const DevPlan &P, MmaConst &qc, std::vector<double> &buf,
                                std::vector<int> &ibuf) {
    constexpr int NU = D::NU, NY = D::NY, M = D::M, NV = D::NV, NVL = D::NVL, R = D::R, RE = D::RE, NYP = D::NYP;
    constexpr int SV = D::SV;
    const int NX = d->nx;                     // states of this model (<= D::NX), stride of the descriptor's blocks
    // live[position] = logical variable; the last NVL-NV dead variables are presolved
    std::vector<int> live, deadv;
    {
        std::vector<int> dd = mma_dead_variables(P, d);
        std::vector<char> drop(NVL, 0);
        for (int i = 0; i < NVL - NV; ++i) drop[dd[dd.size() - 1 - i]] = 1;
        for (int v = 0; v < NVL; ++v) (drop[v] ? deadv : live).push_back(v);
    }
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    qc.be = take(D::TR * (D::KX + D::KA) * 32);
    qc.bw = take(D::TR * D::KP * 32);
    qc.w0c = take(D::TR * 8);
    qc.bf = take(D::TV * 2 * D::TR * 32);
    qc.br = take(D::TR * SV * 32);
    qc.bh = take(D::HPAIRS * D::KH * 32);
    qc.bomdu = take(D::TV * D::KP * 32);
    qc.by = take(D::TC * (D::KX + D::KD) * 32);
    qc.byp = take(D::TQ * (D::KX + SV + D::KE) * 32);
    qc.bvn = take(D::TM * (2 * D::TQ + SV) * 32);
    qc.tau = take(D::TY * 8);
    qc.gamma = take(M);
    qc.inv_gamma = take(M);
    qc.bounds = take(4 * NU);
    qc.vinq = take(d->n_vin > 0 ? d->n_vin : 1);
    const int W = D::WELL;
    qc.ell_w = W;
    qc.ell_val = take(D::TX * 2 * W * 4);
    qc.bfc = take(D::TV * 2 * D::TR * 32);
    qc.bhc = take(D::HPAIRS * D::KHC * 32);
    qc.bdp = take(D::TV * SV * 32);
    qc.bdpt = take(D::TV * SV * 32);
    qc.bcum = take(D::TV * SV * 32);
    qc.total_d = o;
    buf.assign(o, 0.0);

    // index maps
    auto crow = [](int lim) { return [lim](int T, int n) { int r = 8 * T + n; return r < lim ? r : -1; }; };
    auto ccol = [](int lim) { return [lim](int t, int k) { int r = 8 * (t / 2) + 2 * k + (t % 2); return r < lim ? r : -1; }; };
    auto vcol = [&](int t, int k) { return (t < SV) ? live[D::var_of(t, k)] : -1; };       // dU in factorisation slots
    auto vrow = [&](int T, int n) {                                                            // gradient tiles -> slots
        if (NV % 8 == 4 && T == D::TV - 1) return (n % 2 == 0) ? live[NV - 4 + n / 2] : -1;
        int r = 8 * T + n; return r < NV ? live[r] : -1;
    };
    // e = [Px | Py | e0] [x ; ysp ; 1], rows R.. = xi(N+1) map (Mi1x)
    mma_pack(buf, qc.be, D::TR, D::KX + D::KA, crow(RE),
             [&](int t, int k) { if (t < D::KX) { int c = 8 * (t / 2) + 2 * k + (t % 2); return c < NX ? c : -1; }
                                 int e = 4 * (t - D::KX) + k; return e <= NY ? NX + e : -1; },
             [&](int r, int c) {
                 if (c < NX) return r < R ? d->Px[(size_t)r * NX + c] : d->Mi1x[(size_t)(r - R) * NX + c];
                 if (c < NX + NY) return r < R ? d->Py[(size_t)r * NY + (c - NX)] : 0.0;
                 return r < R ? d->e0[r] : 0.0; });
    // omega = [Wrow | w0row] [pesos ; 1]
    mma_pack(buf, qc.bw, D::TR, D::KP, crow(R),
             [&](int t, int k) { int e = 4 * t + k; return e < M ? e : -1; },
             [&](int r, int c) { return d->Wrow[(size_t)r * M + c]; });
    for (int T = 0; T < D::TR; ++T)
        for (int g = 0; g < 4; ++g)
            for (int h = 0; h < 2; ++h) { int r = 8 * T + 2 * g + h; if (r < R) buf[qc.w0c + (T * 4 + g) * 2 + h] = d->w0row[r]; }
    // f = 2 E^T t : out = variables (slot order), in = residual rows (C-layout)
    mma_pack(buf, qc.bf, D::TV, 2 * D::TR, vrow, ccol(R),
             [&](int v, int r) { return 2.0 * d->E[(size_t)r * NVL + v]; });
    // r += E_ext dU
    mma_pack(buf, qc.br, D::TR, SV, crow(RE), vcol,
             [&](int r, int v) { return r < R ? d->E[(size_t)r * NVL + v] : d->Mi1u[(size_t)(r - R) * NVL + v]; });
    // omega_du = Wdu pesos  (delivered in factorisation slots)
    mma_pack(buf, qc.bomdu, D::TV, D::KP, vrow,
             [&](int t, int k) { int e = 4 * t + k; return e < M ? e : -1; },
             [&](int v, int c) { return d->Wdu[(size_t)v * M + c]; });
    // Hessian pairs: tile pp row n=2g'+h <-> L[t][2jp+h] of lane g' ; in = [1 ; pw_a]
    {
        int pp = 0;
        for (int t = 0; t < SV; ++t)
            for (int jp = 0; jp < (4 * t + 4) / 2; ++jp, ++pp) {
                mma_pack(buf, qc.bh + pp * D::KH * 32, 1, D::KH,
                         [&](int, int n) { return n; },
                         [&](int tt, int k) { int e = 4 * tt + k; return e <= d->mh ? e : -1; },
                         [&](int n, int e) {
                             const int gq = n / 2, h = n % 2;
                             const int prow = 4 * t + gq, pcol = 2 * jp + h;      // positions in elimination order
                             if (pcol > prow) return 0.0;
                             const int vi = live[D::var_of(prow / 4, prow % 4)], vj = live[D::var_of(pcol / 4, pcol % 4)];
                             return e == 0 ? d->Ht[(size_t)vi * NVL + vj] : d->HJ[((size_t)(e - 1) * NVL + vi) * NVL + vj]; });
            }
    }
    // y = [C | D] [x+ ; du]
    mma_pack(buf, qc.by, D::TC, D::KX + D::KD, crow(NY),
             [&](int t, int k) { if (t < D::KX) { int c = 8 * (t / 2) + 2 * k + (t % 2); return c < NX ? c : -1; }
                                 int e = 4 * (t - D::KX) + k; return e < NU ? NX + e : -1; },
             [&](int r, int c) { return c < NX ? d->C[(size_t)r * NX + c] : d->D[(size_t)r * NU + (c - NX)]; });
    // residual stack of the shifted plan: rows r < NYP: Y' - ysp - syN - tau siN with Y' = Myx x+ + Myu shift(dU)
    // (column v of the shifted map is Myu[:, v - NU]); rows NYP..: syN, siN.  Inputs: x+, dU (slots), [ysp ; syN ; siN]
    mma_pack(buf, qc.byp, D::TQ, D::KX + SV + D::KE, crow(NYP + 2 * NY),
             [&](int t, int k) { if (t < D::KX) { int c = 8 * (t / 2) + 2 * k + (t % 2); return c < NX ? c : -1; }
                                 if (t < D::KX + SV) return NX + live[D::var_of(t - D::KX, k)];
                                 int e = 4 * (t - D::KX - SV) + k; return e < 3 * NY ? 5000 + e : -1; },
             [&](int r, int c) {
                 if (c >= 5000) {
                     const int e = c - 5000;
                     if (r < NYP) {
                         const int i = r % NY, kk = r / NY;
                         if (e < 2 * NY) return (e % NY == i) ? -1.0 : 0.0;          // - ysp_i , - syN_i
                         return (e - 2 * NY == i) ? -d->tau[kk] : 0.0;             // - tau_k siN_i
                     }
                     return (e == NY + (r - NYP)) ? 1.0 : 0.0;                      // rows syN_i, siN_i
                 }
                 if (r >= NYP) return 0.0;
                 if (c < NX) return d->Myx[(size_t)r * NX + c];
                 int v = c - NX; return v >= NU ? d->Myu[(size_t)r * NVL + (v - NU)] : 0.0; });
    // V_i = sum_r Wrow[r][i] sq_r + sum_v Wdu[v-NU][i] dU_v^2
    mma_pack(buf, qc.bvn, D::TM, 2 * D::TQ + SV, crow(M),
             [&](int t, int k) { if (t < 2 * D::TQ) { int r = 8 * (t / 2) + 2 * k + (t % 2); return r < NYP + 2 * NY ? r : -1; }
                                 return 1000 + live[D::var_of(t - 2 * D::TQ, k)]; },
             [&](int i, int c) { if (c < 1000) return d->Wrow[(size_t)c * M + i];
                                 int v = c - 1000; return v >= NU ? d->Wdu[(size_t)(v - NU) * M + i] : 0.0; });
    // ---- bound space of the in-kernel active set (mma_active_set): variables c = one per live move, dU_live = S c
    // with S = difference operator along the horizon per input (c = cumulative move: the input box becomes a simple
    // bound; presolved moves are zero, so the predecessor of a live move is the nearest earlier live move of the same
    // input) or S = I (move box).  Indices below are LIVE indices (position in `live`).
    std::vector<int> du_w;                                   // weights that act on the moves
    for (int i = 0; i < M; ++i) {
        bool any = false;
        for (int v = 0; v < NVL && !any; ++v) any = (d->Wdu[(size_t)v * M + i] != 0.0);
        if (any) du_w.push_back(i);
    }
    qc.n_du = (int)du_w.size();
    qc.cspace = (P.has_cum && !P.has_box) ? 1 : ((P.has_box && !P.has_cum) ? 0 : -1);
    if (1 + d->mh + qc.n_du > 4 * D::KHC || getenv("SIHMPC_NO_INKERNEL_AS")) qc.cspace = -1;
    {
        std::vector<double> S((size_t)NV * NV, 0.0);        // dU_live[a] = sum_b S[a][b] c[b]
        for (int a = 0; a < NV; ++a) {
            S[(size_t)a * NV + a] = 1.0;
            if (qc.cspace == 1) {
                int prev = -1;
                for (int b = 0; b < NV; ++b)
                    if (live[b] % NU == live[a] % NU && live[b] < live[a] && (prev < 0 || live[b] > live[prev])) prev = b;
                if (prev >= 0) S[(size_t)a * NV + prev] = -1.0;
            }
        }
        auto lcol = [&](int t, int k) { return (t < SV) ? D::var_of(t, k) : -1; };
        auto lrow = [&](int T, int n) {
            if (NV % 8 == 4 && T == D::TV - 1) return (n % 2 == 0) ? NV - 4 + n / 2 : -1;
            int r = 8 * T + n; return r < NV ? r : -1;
        };
        mma_pack(buf, qc.bdp, D::TV, SV, lrow, lcol, [&](int a, int b) { return S[(size_t)a * NV + b]; });
        mma_pack(buf, qc.bdpt, D::TV, SV, lrow, lcol, [&](int a, int b) { return S[(size_t)b * NV + a]; });
        // cumulative move of input j after the live move a: c_a = sum of the live moves of the same input up to a
        // (presolved moves are zero) -- the rows of the input box ihmpccontroller.py:420-424 at the unconstrained optimum
        mma_pack(buf, qc.bcum, D::TV, SV, lrow, lcol,
                 [&](int a, int b) { return (live[b] % NU == live[a] % NU && live[b] <= live[a]) ? 1.0 : 0.0; });
        qc.cum0 = 0;
        for (int j = 0; j < NU; ++j) {
            bool first_live = false;
            for (int a = 0; a < NV; ++a) if (live[a] == j) first_live = true;
            if (!first_live) qc.cum0 |= 1 << j;
        }
        // gradient map: 2 (E_live S)^T
        std::vector<double> Ec((size_t)R * NV, 0.0);
        for (int r = 0; r < R; ++r)
            for (int b = 0; b < NV; ++b) {
                double acc = 0.0;
                for (int a = 0; a < NV; ++a) acc += d->E[(size_t)r * NVL + live[a]] * S[(size_t)a * NV + b];
                Ec[(size_t)r * NV + b] = acc;
            }
        mma_pack(buf, qc.bfc, D::TV, 2 * D::TR, lrow, ccol(R), [&](int b, int r) { return 2.0 * Ec[(size_t)r * NV + b]; });
        // Hessian blocks S^T X S for X = Ht, HJ_a, 2 diag(Wdu[:, i])
        const int nblk = 1 + d->mh + qc.n_du;
        std::vector<double> Hc((size_t)nblk * NV * NV, 0.0), X((size_t)NV * NV), XS((size_t)NV * NV);
        for (int e = 0; e < nblk; ++e) {
            for (int a = 0; a < NV; ++a)
                for (int b = 0; b < NV; ++b) {
                    const int va = live[a], vb = live[b];
                    double x;
                    if (e == 0) x = d->Ht[(size_t)va * NVL + vb];
                    else if (e <= d->mh) x = d->HJ[((size_t)(e - 1) * NVL + va) * NVL + vb];
                    else x = (a == b) ? 2.0 * d->Wdu[(size_t)va * M + du_w[e - 1 - d->mh]] : 0.0;
                    X[(size_t)a * NV + b] = x;
                }
            for (int a = 0; a < NV; ++a)
                for (int b = 0; b < NV; ++b) {
                    double acc = 0.0;
                    for (int k = 0; k < NV; ++k) acc += X[(size_t)a * NV + k] * S[(size_t)k * NV + b];
                    XS[(size_t)a * NV + b] = acc;
                }
            for (int a = 0; a < NV; ++a)
                for (int b = 0; b < NV; ++b) {
                    double acc = 0.0;
                    for (int k = 0; k < NV; ++k) acc += S[(size_t)k * NV + a] * XS[(size_t)k * NV + b];
                    Hc[((size_t)e * NV + a) * NV + b] = acc;
                }
        }
        int pp = 0;
        for (int t = 0; t < SV; ++t)
            for (int jp = 0; jp < (4 * t + 4) / 2; ++jp, ++pp)
                mma_pack(buf, qc.bhc + pp * D::KHC * 32, 1, D::KHC,
                         [&](int, int n) { return n; },
                         [&](int tt, int k) { int e = 4 * tt + k; return e < nblk && e < 4 * D::KHC ? e : -1; },
                         [&](int n, int e) {
                             const int gq = n / 2, h = n % 2;
                             const int prow = 4 * t + gq, pcol = 2 * jp + h;
                             if (pcol > prow) return 0.0;
                             const int a = D::var_of(prow / 4, prow % 4), b = D::var_of(pcol / 4, pcol % 4);
                             return Hc[((size_t)e * NV + a) * NV + b]; });
    }
    for (int T = 0; T < D::TY; ++T)
        for (int g = 0; g < 4; ++g)
            for (int h = 0; h < 2; ++h) { int r = 8 * T + 2 * g + h; if (r < NYP) buf[qc.tau + (T * 4 + g) * 2 + h] = d->tau[r / NY]; }
    for (int i = 0; i < M; ++i) { buf[qc.gamma + i] = d->gamma[i]; buf[qc.inv_gamma + i] = 1.0 / d->gamma[i]; }
    for (int j = 0; j < NU; ++j) {
        buf[qc.bounds + j] = d->dulb[j]; buf[qc.bounds + NU + j] = d->duub[j];
        buf[qc.bounds + 2 * NU + j] = d->ulb[j]; buf[qc.bounds + 3 * NU + j] = d->uub[j];
    }
    for (int sIdx = 0; sIdx < d->n_vin; ++sIdx) buf[qc.vinq + sIdx] = d->vin_q[sIdx];
    int io_ = 0;
    auto takei = [&](int n) { int r = io_; io_ += n; return r; };
    qc.ell_col = takei(D::TX * 2 * W * 4);
    qc.ell_csr = takei(D::TX * 2 * W * 4);
    qc.hj_idx = takei(M);
    qc.vin_row = takei(d->n_vin > 0 ? d->n_vin : 1);
    qc.vlive = takei(NV);
    qc.vdead = takei(NVL - NV > 0 ? NVL - NV : 1);
    qc.du_idx = takei(M);
    qc.total_i = io_;
    ibuf.assign(io_, 0);
    for (int i = 0; i < qc.n_du; ++i) ibuf[qc.du_idx + i] = du_w[i];
    for (int i = 0; i < NV; ++i) ibuf[qc.vlive + i] = live[i];
    for (int i = 0; i < NVL - NV; ++i) ibuf[qc.vdead + i] = deadv[i];
    for (int T = 0; T < D::TX; ++T)
        for (int h = 0; h < 2; ++h)
            for (int g = 0; g < 4; ++g) {
                const int i = 8 * T + 2 * g + h;
                for (int w = 0; w < W; ++w) {
                    const int e = ((T * 2 + h) * W + w) * 4 + g;
                    ibuf[qc.ell_csr + e] = -1; ibuf[qc.ell_col + e] = 0;
                    if (i < NX) {
                        const int p = d->pl_rowptr[i] + w;
                        if (p < d->pl_rowptr[i + 1]) {
                            const int c = d->pl_col[p];
                            ibuf[qc.ell_csr + e] = p;
                            ibuf[qc.ell_col + e] = (c < NX) ? c : D::BX + (c - NX);     // du sits after x in the buffer
                            buf[qc.ell_val + e] = d->pl_val[p];
                        }
                    }
                }
            }
    for (int a = 0; a < d->mh; ++a) ibuf[qc.hj_idx + a] = d->hj_index[a];
    for (int sIdx = 0; sIdx < d->n_vin; ++sIdx) ibuf[qc.vin_row + sIdx] = d->vin_row[sIdx] - (NYP + NY);
}

template <typename D>
inline size_t mma_smem_bytes(const MmaConst &qc) {
    return ((size_t)qc.total_d + (((qc.total_i + 3) >> 2) << 1) + (size_t)MMA_WARPS * D::WARP_DOUBLES) * sizeof(double);
}

template <typename D>
inline cudaError_t mma_configure_shape(const sihmpc_plan_desc *d, const DevPlan &P, MmaPlan *qp) {
    std::vector<double> buf; std::vector<int> ibuf;
    mma_build_constants<D>(d, P, qp->qc, buf, ibuf);
    const size_t nd = buf.size(), ni = ibuf.size();
    std::vector<double> all(nd + ((ni + 1) >> 1), 0.0);
    memcpy(all.data(), buf.data(), nd * sizeof(double));
    memcpy(all.data() + nd, ibuf.data(), ni * sizeof(int));
    cudaError_t e = cudaMalloc(&qp->qbuf, all.size() * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(qp->qbuf, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    qp->smem_bytes = mma_smem_bytes<D>(qp->qc);
    if (qp->smem_bytes > 227 * 1024) { qp->shape = -1; return cudaSuccess; }
    auto kern = mma_step_kernel<D, MMA_WARPS, MMA_MINB, 0>;
    auto kern_resume = mma_step_kernel<D, MMA_WARPS, MMA_MINB, 1>;
    auto kern_resident = mma_step_kernel<D, MMA_WARPS, MMA_MINB, 3>;
    auto kern_stream = mma_step_kernel<D, MMA_WARPS, MMA_MINB, 4>;
    e = sih_raise_smem_limit(reinterpret_cast<const void *>(kern), qp->smem_bytes);
    if (e == cudaSuccess) e = sih_raise_smem_limit(reinterpret_cast<const void *>(kern_resume), qp->smem_bytes);
    if (e == cudaSuccess) e = sih_raise_smem_limit(reinterpret_cast<const void *>(kern_resident), qp->smem_bytes);
    if (e == cudaSuccess) e = sih_raise_smem_limit(reinterpret_cast<const void *>(kern_stream), qp->smem_bytes);
    if (e != cudaSuccess) return e;
    int occ = 0, occ_r = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MMA_WARPS * 32, qp->smem_bytes);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, kern_resume, MMA_WARPS * 32, qp->smem_bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) { qp->shape = -1; return cudaSuccess; }
    if (occ_r < 1) qp->qc.cspace = -1;
    qp->blocks_per_sm = occ;
    return cudaSuccess;
}

template <typename D>
inline cudaError_t mma_launch_shape(MmaPlan &qp, const DevPlan &P, const sihmpc_io &io, long long batch,
                                    int n_steps, int grid, cudaStream_t st, bool allow_resume) {
    const bool stream_resume = allow_resume && !qp.res_active && qp.aux && mma_resumes(qp, batch, n_steps) && qp.stream_workers > 0 &&
                               batch >= qp.stream_min_batch;
    if (stream_resume) {
        // workers on the auxiliary stream, concurrent with the first pass (which leaves them room when it fills the GPU)
        const int nw = qp.stream_workers;
        const int cap = qp.blocks_per_sm * qp.sm_count;
        const int g1 = (grid + nw > cap) ? ((cap - nw > 1) ? cap - nw : 1) : grid;
        cudaError_t e = cudaEventRecord(qp.ev_fork, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(qp.aux, qp.ev_fork, 0);
        if (e != cudaSuccess) return e;
        mma_step_kernel<D, MMA_WARPS, MMA_MINB, 4><<<nw, MMA_WARPS * 32, qp.smem_bytes, qp.aux>>>(
            P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_fail, qp.n_inkernel, qp.list, qp.n_handover,
            nullptr, nullptr, 0, qp.ctr + 3, g1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        mma_step_kernel<D, MMA_WARPS, MMA_MINB, 0><<<g1, MMA_WARPS * 32, qp.smem_bytes, st>>>(
            P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_handover, qp.n_inkernel, qp.list, nullptr,
            nullptr, nullptr, 0, qp.ctr + 3, g1);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(qp.ev_join, qp.aux);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, qp.ev_join, 0);
        qp.last_resumed = 1;
        if (e != cudaSuccess) return e;
        // sweep: the sequential resume pass over whatever the workers did not take (nothing, normally: resolved entries
        // carry start_step = "done" and cost a state load).  It is what keeps the result right when the two kernels do
        // NOT overlap -- a profiler or CUDA_LAUNCH_BLOCKING serialises them, the workers then give up after their bounded wait
        mma_step_kernel<D, MMA_WARPS, MMA_MINB, 1><<<grid, MMA_WARPS * 32, qp.smem_bytes, st>>>(
            P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_fail, qp.n_inkernel, qp.list, qp.n_handover);
        return cudaGetLastError();
    }
    if (qp.res_active) {
        mma_step_kernel<D, MMA_WARPS, MMA_MINB, 3><<<grid, MMA_WARPS * 32, qp.smem_bytes, st>>>(
            P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_handover, qp.n_inkernel, qp.list, nullptr,
            qp.res_wstart, qp.res_ysp_prev, qp.res_flags);
    } else {
        mma_step_kernel<D, MMA_WARPS, MMA_MINB, 0><<<grid, MMA_WARPS * 32, qp.smem_bytes, st>>>(
            P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_handover, qp.n_inkernel, qp.list, nullptr);
    }
    cudaError_t e = cudaGetLastError();
    qp.last_resumed = 0;
    if (e != cudaSuccess || !allow_resume || !mma_resumes(qp, batch, n_steps)) return e;
    qp.last_resumed = 1;
    // resume pass over the listed scenarios (their number is only known on the device: blocks beyond it leave at once)
    mma_step_kernel<D, MMA_WARPS, MMA_MINB, 1><<<grid, MMA_WARPS * 32, qp.smem_bytes, st>>>(
        P, qp.qc, qp.qbuf, io, batch, n_steps, qp.start_step, qp.n_fail, qp.n_inkernel, qp.list, qp.n_handover);
    return cudaGetLastError();
}

}  // namespace sih
