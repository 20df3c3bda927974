// mma_units.cu -- instantiations of the tensor fast path v3 (mma_kernel.cuh), one translation unit per
// (horizon N, state capacity NXP) so that build.py compiles them in parallel:
//     -DSIH_MMA_N=<6|8|10|12> -DSIH_MMA_NXP=<16|24|32>   ->  mma_unit_<N>_<NXP>(...)   (both presolve variants)
// Without the defines: the host side (shape matching, dispatch).
//
// Family: 2x2 models with one pole per channel (nd = 4) and any dead times (nx = 8 + 2 theta_max <= 32), control
// horizons N = 6, 8, 10, 12 (N*nu = 12 ... 24), the four weighted sub-objectives of example_distillation.py:49-53.
// Moves whose dead time pushes their whole effect past the horizon are presolved in groups of four (mma_dead_variables).
#define SIH_NO_QUAD_HOST
#include "mma_kernel.cuh"

namespace sih {

#define SIH_MMA_CAT3(a, b, c) a##b##_##c
#define SIH_MMA_NAME(N_, X_) SIH_MMA_CAT3(mma_unit_, N_, X_)

#if defined(SIH_MMA_N) && defined(SIH_MMA_NXP)
// op 0: configure, op 1: launch
cudaError_t SIH_MMA_NAME(SIH_MMA_N, SIH_MMA_NXP)(int op, int nve, const sihmpc_plan_desc *d, const DevPlan &P, MmaPlan &qp,
                                                 const sihmpc_io *io, long long batch, int n_steps, int grid, cudaStream_t st,
                                                 bool allow_resume) {
    using Full = MmaDims<SIH_MMA_NXP, 2, 2, 4, SIH_MMA_N, 4>;
    using Pre = MmaDims<SIH_MMA_NXP, 2, 2, 4, SIH_MMA_N, 4, 2 * SIH_MMA_N - 4>;
    if (nve == 2 * SIH_MMA_N)
        return op == 0 ? mma_configure_shape<Full>(d, P, &qp) : mma_launch_shape<Full>(qp, P, *io, batch, n_steps, grid, st, allow_resume);
    return op == 0 ? mma_configure_shape<Pre>(d, P, &qp) : mma_launch_shape<Pre>(qp, P, *io, batch, n_steps, grid, st, allow_resume);
}
#else
#define SIH_MMA_DECL(N_, X_) \
    cudaError_t SIH_MMA_NAME(N_, X_)(int, int, const sihmpc_plan_desc *, const DevPlan &, MmaPlan &, const sihmpc_io *, long long, \
                                     int, int, cudaStream_t, bool);
#define SIH_MMA_FOR_ALL(F) F(6, 16) F(6, 24) F(6, 32) F(8, 16) F(8, 24) F(8, 32) F(10, 16) F(10, 24) F(10, 32) F(12, 16) F(12, 24) F(12, 32)
SIH_MMA_FOR_ALL(SIH_MMA_DECL)

static cudaError_t mma_unit(const MmaPlan &sel, int op, const sihmpc_plan_desc *d, const DevPlan &P, MmaPlan &qp,
                            const sihmpc_io *io, long long batch, int n_steps, int grid, cudaStream_t st, bool allow_resume) {
#define SIH_MMA_CASE(N_, X_) if (sel.N == N_ && sel.nxp == X_) return SIH_MMA_NAME(N_, X_)(op, sel.nve, d, P, qp, io, batch, n_steps, grid, st, allow_resume);
    SIH_MMA_FOR_ALL(SIH_MMA_CASE)
    return cudaErrorInvalidValue;
}

// fills N / nxp / nve when the controller belongs to the compiled family; returns false otherwise
static bool mma_match_shape(const DevPlan &P, const sihmpc_plan_desc *d, MmaPlan *qp) {
    for (int i = 0; i < P.nx; ++i)
        if (d->pl_rowptr[i + 1] - d->pl_rowptr[i] > 4) return false;      // ELL width of the fast path
    if (!(P.nu == 2 && P.ny == 2 && P.nd == 4 && P.m == 4 && P.ncd == 0 && P.mh <= 3 && P.nv == P.nvm)) return false;
    if (!(P.N == 6 || P.N == 8 || P.N == 10 || P.N == 12) || P.nx > 32) return false;
    qp->N = P.N;
    qp->nxp = P.nx <= 16 ? 16 : (P.nx <= 24 ? 24 : 32);
    const bool presolve = !getenv("SIHMPC_NO_PRESOLVE") && mma_dead_variables(P, d).size() >= 4;
    qp->nve = presolve ? 2 * P.N - 4 : 2 * P.N;
    return true;
}

cudaError_t mma_configure(const sihmpc_plan_desc *d, const DevPlan &P, MmaPlan *qp) {
    memset(qp, 0, sizeof *qp);
    qp->shape = -1;
    if (!mma_match_shape(P, d, qp)) return cudaSuccess;
    qp->shape = 0;
    const MmaPlan sel = *qp;
    cudaError_t e = mma_unit(sel, 0, d, P, *qp, nullptr, 0, 0, 0, nullptr, false);
    if (e != cudaSuccess) return e;
    if (qp->shape >= 0) {            // (mma_configure_shape sets shape = -1 itself when the kernel cannot be resident)
        cudaDeviceProp prop;
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return e;
        qp->sm_count = prop.multiProcessorCount;
        qp->stream_workers = getenv("SIHMPC_NO_STREAM_RESUME") ? 0 : 2;
        if (const char *wv = getenv("SIHMPC_STREAM_WORKERS")) { const int v = atoi(wv); if (v >= 0 && v <= 32) qp->stream_workers = v; }
        qp->stream_min_batch = 16384;
        if (const char *bv = getenv("SIHMPC_STREAM_MIN_BATCH")) { const long long v = atoll(bv); if (v >= 0) qp->stream_min_batch = v; }
        e = cudaStreamCreateWithFlags(&qp->aux, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&qp->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&qp->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t mma_launch(MmaPlan &qp, const DevPlan &P, const sihmpc_io &io, long long batch, int n_steps, int sm_count,
                       cudaStream_t st, bool allow_resume) {
    const long long per_block = MMA_WARPS * 8;
    long long need = (batch + per_block - 1) / per_block;
    const long long cap = (long long)sm_count * qp.blocks_per_sm;
    int grid = (int)(need < cap ? need : cap);
    if (qp.grid_cap > 0 && qp.grid_cap < grid) grid = qp.grid_cap;      // debug override (SIHMPC_QUAD_GRID, read at plan creation)
    const MmaPlan sel = qp;
    return mma_unit(sel, 1, nullptr, P, qp, &io, batch, n_steps, grid, st, allow_resume);
}

void mma_destroy(MmaPlan *qp) {
    if (qp->qbuf) cudaFree(qp->qbuf);
    if (qp->aux) cudaStreamDestroy(qp->aux);
    if (qp->ev_fork) cudaEventDestroy(qp->ev_fork);
    if (qp->ev_join) cudaEventDestroy(qp->ev_join);
    memset(qp, 0, sizeof *qp);
    qp->shape = -1;
}
#endif

}  // namespace sih
