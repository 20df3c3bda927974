// devplan.h -- device-side view of a compiled controller (all pointers are device memory).
// Layout note: every matrix that is consumed with "lanes over rows" is stored TRANSPOSED
// ([col][row]) so that the 32 lanes of a warp read consecutive doubles (coalesced, and
// bank-conflict free once staged); the suffix T marks those copies.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include "../../include/sihmpc_b200.h"

#ifndef SIH_FULL
#define SIH_FULL 0xffffffffu
#define SIH_INF (__longlong_as_double(0x7ff0000000000000LL))
#endif

struct DevPlan {
    int nx, nu, ny, nd, N, nv, m, mh, R, ncd, n_vin, pl_nnz;
    int nvm;             // N*nu: the moves among the nv decision variables (nv - nvm = ranged terminal residuals)
    int Ny;              // N*ny
    int o_sy, o_si, o_t; // residual row offsets
    int has_box, has_cum;
    int eval_term;       // some weighted sub-objective includes the terminal rows (evaluation plans)
    int ld;              // leading dimension of the in-smem Hessian / factor (odd)
    double Ts;
    const double *E, *ET, *PxT, *PyT, *e0;
    const double *Wrow, *WrowT, *w0row, *Wdu, *WduT;
    const double *HJ, *Ht;
    const int32_t *hj_index;
    const double *MyxT, *MyuT, *Mi1xT, *Mi1uT, *tau, *gamma;
    const double *dulb, *duub, *ulb, *uub;
    const double *Gc, *GcT, *GcxT, *GcyT, *gc0, *Lc, *Uc, *gc_q;
    const int32_t *gc_vin;
    const int32_t *vin_row;
    const double *vin_q;
    const int32_t *pl_rowptr, *pl_col;
    const double *pl_val;
    const double *CT, *DT;
    const double *PhiT, *GamT;   // [nx][N*nx], [nv][N*nx] for sihmpc_predict
};

// per-warp shared-memory layout (offsets in doubles)
struct WarpLayout {
    int L, dinv, xin, xnew, uu, ysp, pes, vin, e, om, omdu, f, v, nrm, ybuf, dbuf, rbuf, cbuf,
        sy, yout, glo, ghi, vn, st, te, pwh, total;
};

__host__ __device__ inline WarpLayout make_layout(const DevPlan &P) {
    WarpLayout w;
    int o = 0;
#define SIH_TAKE(field, n) do { w.field = o; o += ((n) + 1) & ~1; } while (0)   /* 16-byte aligned */
    SIH_TAKE(L, P.nv * P.ld);
    SIH_TAKE(dinv, P.nv);
    SIH_TAKE(xin, P.nx + P.nu);
    SIH_TAKE(xnew, P.nx);
    SIH_TAKE(uu, P.nu);
    SIH_TAKE(ysp, P.ny);
    SIH_TAKE(pes, P.m > 0 ? P.m : 1);
    SIH_TAKE(vin, P.n_vin > 0 ? P.n_vin : 1);
    SIH_TAKE(e, P.R);
    SIH_TAKE(om, P.R);
    SIH_TAKE(omdu, P.nv);
    SIH_TAKE(f, P.nv);
    SIH_TAKE(v, P.nv);
    SIH_TAKE(nrm, P.nv);
    SIH_TAKE(ybuf, P.nv);
    SIH_TAKE(dbuf, P.nv);
    SIH_TAKE(rbuf, P.nv);
    SIH_TAKE(cbuf, P.nv);
    SIH_TAKE(sy, 2 * P.ny);
    SIH_TAKE(yout, P.ny);
    SIH_TAKE(glo, P.ncd > 0 ? P.ncd : 1);
    SIH_TAKE(ghi, P.ncd > 0 ? P.ncd : 1);
    SIH_TAKE(vn, P.m > 0 ? P.m : 1);
    SIH_TAKE(st, P.ny + SIHMPC_NSTATS_EXTRA);
    SIH_TAKE(te, P.R);
    SIH_TAKE(pwh, P.mh > 0 ? P.mh : 1);
#undef SIH_TAKE
    w.total = o;
    return w;
}

// per-warp global scratch for the active set (doubles): Y[nv][nv], S[nv][lds], Rf[nv][lds],
// lam[nv], rdinv[nv], act[nv] (stored as int32 in a double-sized slot each)
__host__ __device__ inline size_t scratch_doubles_per_warp(int nv) {
    size_t lds = (size_t)(nv | 1);
    return (size_t)nv * nv + 2 * (size_t)nv * lds + 3 * (size_t)nv + 8;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the KERNEL, not to a plan: two live plans served by the
// same instantiation (the controller's hot plan and its evaluation plan, two shapes with the same tile-row count)
// must not lower each other's limit -- it is only ever raised.
inline cudaError_t sih_raise_smem_limit(const void *func, size_t bytes) {
    static std::mutex mu;
    static std::map<const void *, size_t> limit;
    std::lock_guard<std::mutex> lock(mu);
    auto it = limit.find(func);
    if (it != limit.end() && it->second >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) limit[func] = bytes;
    return e;
}
