// capi.cu -- C ABI of libsihmpc_b200.so (declared in include/sihmpc_b200.h).
// Host side only: plan upload (with the transposed copies the kernels want), launch
// configuration, staging for the host-buffer entry point.  No torch, no Python.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <string>

#include "../../include/sihmpc_b200.h"
#include "devplan.h"
#include "step_kernel.cuh"
#include "dist_kernel.cuh"
#include "mma_api.h"
#include "tile_api.h"
#include "log_kernel.cuh"

namespace {

thread_local std::string g_err = "";

int fail(int code, const char *fmt, const char *detail = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, detail);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            char b__[512];                                                                   \
            snprintf(b__, sizeof b__, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
            g_err = b__;                                                                     \
            return SIHMPC_E_CUDA;                                                            \
        }                                                                                    \
    } while (0)

}  // namespace

struct sihmpc_plan {
    DevPlan d;
    int device;
    int sm_count;
    int nvr;                 // rows per lane of the general kernel
    int wpb;                 // warps per block
    int blocks_per_sm;
    size_t smem_bytes;
    int max_grid;
    int scratch_in_smem;     // general kernel: active-set workspace in shared memory (small controllers)
    sih::QuadPlan quad;      // register-tiled fast path v2 (dist_kernel.cuh), shape<0: none
    sih::MmaPlan mma;        // register-tiled + FP64 tensor-core fast path v3 (mma_kernel.cuh), shape<0: none
    sih::TilePlan tile;      // CTA-cooperative tile path v4 for mid-size controllers (tile_kernel.cuh), ok=0: none
    int use_fast;            // 0: general kernel only, 1: quad, 2: mma where it matches, else tile (default), 3: tile
                             // (SIHMPC_NO_FAST=1 / SIHMPC_FAST=quad|tile for A/B)
    std::vector<void *> allocs;
    // host-buffer entry point: copy streams + events for the chunked pipeline, pinned pack buffer
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> events;
    void *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    // host copies needed by the host-buffer entry point
    int nw;
    // environment overrides (A/B measurements only), read ONCE at plan creation
    int env_bounce = 0;            // SIHMPC_BOUNCE: bounce rounds of the tile path
    int env_no_zerocopy = 0;       // SIHMPC_NO_ZEROCOPY
    long long env_zc_max = 1LL << 40;   // SIHMPC_ZEROCOPY_MAX
    int env_chunks = 0;            // SIHMPC_HOST_CHUNKS (0: default)
    int env_grid = 0;              // SIHMPC_QUAD_GRID (debug: force many rounds)
    long long resident_steps = -1; // controller-resident batch: steps taken since sihmpc_resident_init (-1: none)
    long long resident_batch = 0;
};

// packed-upload path of sihmpc_step_host: largest batch served from the plan's pinned pack buffer
#define SIH_PACK_MAX 512

extern "C" { static size_t staging_layout(const sihmpc_plan *p, int64_t batch, size_t off[12]); }

namespace {

template <typename T>
int upload(sihmpc_plan *p, const T *host, size_t n, const T **dev) {
    *dev = nullptr;
    if (n == 0) {  // keep a valid pointer so kernels can index conditionally
        n = 1;
        void *q = nullptr;
        CUDA_TRY(cudaMalloc(&q, sizeof(T)));
        CUDA_TRY(cudaMemset(q, 0, sizeof(T)));
        p->allocs.push_back(q);
        *dev = static_cast<const T *>(q);
        return 0;
    }
    if (!host) return fail(SIHMPC_E_INVALID, "plan descriptor has a NULL block%s");
    void *q = nullptr;
    CUDA_TRY(cudaMalloc(&q, n * sizeof(T)));
    p->allocs.push_back(q);
    CUDA_TRY(cudaMemcpy(q, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T *>(q);
    return 0;
}

int upload_T(sihmpc_plan *p, const double *host, int rows, int cols, const double **dev) {
    std::vector<double> t((size_t)rows * cols);
    if ((size_t)rows * cols && !host) return fail(SIHMPC_E_INVALID, "plan descriptor has a NULL block%s");
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = host[(size_t)r * cols + c];
    return upload(p, t.data(), t.size(), dev);
}

bool any_finite(const double *v, int n) {
    for (int i = 0; i < n; ++i)
        if (v[i] > -1e300 && v[i] < 1e300) return true;
    return false;
}

template <int NVR>
int configure(sihmpc_plan *p) {
    auto kern = sih::ihmpc_step_kernel<NVR>;
    const WarpLayout lay = make_layout(p->d);
    const size_t per_warp = (size_t)lay.total * sizeof(double);
    const size_t limit = 227 * 1024;
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > limit) --wpb;
    if (per_warp * wpb > limit) return fail(SIHMPC_E_INVALID, "controller too large for shared memory%s");
    // prefer several small CTAs per SM over one large one
    while (wpb > 2 && per_warp * wpb > limit / 2) wpb >>= 1;
    p->wpb = wpb;
    p->smem_bytes = per_warp * wpb;
    // small controllers: the active-set workspace lives in shared memory as well
    const size_t sbytes = scratch_doubles_per_warp(p->d.nv) * sizeof(double);
    p->scratch_in_smem = 0;
    if ((per_warp + sbytes) * wpb <= limit / 2) {
        p->scratch_in_smem = 1;
        p->smem_bytes += sbytes * wpb;
    }
    CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(kern), p->smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, p->smem_bytes));
    if (occ < 1) return fail(SIHMPC_E_INVALID, "step kernel cannot be resident (registers/shared memory)%s");
    p->blocks_per_sm = occ;
    p->max_grid = occ * p->sm_count;
    return 0;
}

template <int NVR>
int launch_step(sihmpc_plan *p, const sihmpc_io &io, long long batch, int n_steps, double *scratch,
                cudaStream_t st, int *start_step = nullptr, const int *n_handover = nullptr,
                int *bounce_list = nullptr, int *n_bounce = nullptr, unsigned long long *warm_mask = nullptr,
                int bounce_steps = 0, int bounce_space = -1) {
    long long need = (batch + p->wpb - 1) / p->wpb;
    int grid = (int)(need < p->max_grid ? need : p->max_grid);
    if (grid < 1) grid = 1;
    const int max_outer = 8 * (2 * p->d.nv + p->d.ncd) + 16;
    sih::ihmpc_step_kernel<NVR><<<grid, p->wpb * 32, p->smem_bytes, st>>>(p->d, io, batch, n_steps,
                                                                           scratch, max_outer, start_step,
                                                                           n_handover, p->scratch_in_smem, bounce_list,
                                                                           n_bounce, warm_mask, bounce_steps, bounce_space);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Hand-over block of the fast paths (counters, start_step[], list[], list2[], warm masks): it lives in the CALLER's
// workspace right after the general kernel's active-set scratch, so nothing is allocated after plan creation.
size_t general_scratch_bytes(const sihmpc_plan *p) {
    return ((size_t)p->max_grid * p->wpb * scratch_doubles_per_warp(p->d.nv) * sizeof(double) + 255) & ~(size_t)255;
}
size_t handover_bytes(int64_t batch) {
    const size_t ints = ((size_t)sih::TL_CTR + 3 * (size_t)batch + 3) & ~(size_t)3;
    return ints * sizeof(int) + 2 * (size_t)batch * sizeof(unsigned long long) + 256;
}
void bind_handover(sihmpc_plan *p, char *base, int64_t batch) {
    int *ctr = reinterpret_cast<int *>(base);
    const size_t ints = ((size_t)sih::TL_CTR + 3 * (size_t)batch + 3) & ~(size_t)3;
    unsigned long long *warm = reinterpret_cast<unsigned long long *>(base + ints * sizeof(int));
    sih::tile_bind(&p->tile, ctr, warm, batch);
    sih::mma_bind(&p->mma, ctr, batch);
    sih::quad_bind(&p->quad, ctr, batch);
}

int check_io(const sihmpc_plan *p, const sihmpc_io *io, int n_steps) {
    if (!io->x || !io->u || !io->pesos) return fail(SIHMPC_E_INVALID, "io.x, io.u and io.pesos are required%s");
    if (!io->ysp && !(io->ysp_seq && io->seg_start && io->n_seg > 0))
        return fail(SIHMPC_E_INVALID, "io.ysp or io.ysp_seq/seg_start is required%s");
    if (n_steps < 1) return fail(SIHMPC_E_INVALID, "n_steps must be >= 1%s");
    if (n_steps > 1 && !(io->flags & SIHMPC_F_ADVANCE))
        return fail(SIHMPC_E_INVALID, "n_steps > 1 needs SIHMPC_F_ADVANCE%s");
    if (p->d.n_vin > 0 && (io->flags & SIHMPC_F_UPDATE_VIN) && !io->vinant && n_steps == 1)
        return fail(SIHMPC_E_INVALID, "SIHMPC_F_UPDATE_VIN needs io.vinant%s");
    return 0;
}

}  // namespace

extern "C" {

int sihmpc_abi_version(void) { return SIHMPC_ABI_VERSION; }

const char *sihmpc_last_error(void) { return g_err.c_str(); }

int sihmpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sihmpc_plan_destroy(sihmpc_plan *plan) {
    if (!plan) return 0;
    for (void *q : plan->allocs) cudaFree(q);
    sih::quad_destroy(&plan->quad);
    sih::mma_destroy(&plan->mma);
    sih::tile_destroy(&plan->tile);
    for (cudaEvent_t e : plan->events) cudaEventDestroy(e);
    if (plan->s_h2d) cudaStreamDestroy(plan->s_h2d);
    if (plan->s_d2h) cudaStreamDestroy(plan->s_d2h);
    if (plan->h_stage) cudaFreeHost(plan->h_stage);
    delete plan;
    return 0;
}

int sihmpc_plan_create(const sihmpc_plan_desc *d, sihmpc_plan **out) {
    if (!d || !out) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    *out = nullptr;
    if (d->abi_version != SIHMPC_ABI_VERSION) return fail(SIHMPC_E_INVALID, "ABI version mismatch%s");
    if (d->nx < 1 || d->nu < 1 || d->ny < 1 || d->N < 1 || d->m < 1 || d->R < 1 || d->nd < 0 ||
        d->mh < 0 || d->ncd < 0 || d->n_vin < 0 || d->pl_nnz < 1)
        return fail(SIHMPC_E_INVALID, "bad dimension in plan descriptor%s");
    if (d->n_extra < 0) return fail(SIHMPC_E_INVALID, "negative n_extra%s");
    const int nv = d->N * d->nu + d->n_extra;
    if (nv > 64) return fail(SIHMPC_E_INVALID, "N*nu + n_extra > 64 is not supported by the step kernel%s");
    if (d->R != d->N * d->ny + 2 * d->ny + d->nd) return fail(SIHMPC_E_INVALID, "R must be N*ny+2*ny+nd%s");
    {   // every block a kernel or the shape matching below dereferences must be there BEFORE anything reads it
        const void *req[] = {d->E, d->Px, d->Py, d->e0, d->Wrow, d->w0row, d->Wdu, d->Ht, d->Myx, d->Myu, d->Mi1x, d->Mi1u,
                             d->tau, d->gamma, d->dulb, d->duub, d->ulb, d->uub, d->pl_rowptr, d->pl_col, d->pl_val, d->C, d->D};
        for (const void *q : req)
            if (!q) return fail(SIHMPC_E_INVALID, "plan descriptor has a NULL block%s");
        if (d->mh > 0 && (!d->HJ || !d->hj_index)) return fail(SIHMPC_E_INVALID, "plan descriptor: mh > 0 needs HJ and hj_index%s");
        if (d->ncd > 0 && (!d->Gc || !d->Gcx || !d->Gcy || !d->gc0 || !d->Lc || !d->Uc || !d->gc_vin || !d->gc_q))
            return fail(SIHMPC_E_INVALID, "plan descriptor: ncd > 0 needs the dense-row blocks%s");
        if (d->n_vin > 0 && (!d->vin_row || !d->vin_q)) return fail(SIHMPC_E_INVALID, "plan descriptor: n_vin > 0 needs vin_row and vin_q%s");
        if ((d->Phi == nullptr) != (d->Gam == nullptr)) return fail(SIHMPC_E_INVALID, "plan descriptor: Phi and Gam come together%s");
    }
    int ndev = sihmpc_device_count();
    if (ndev < 1) return fail(SIHMPC_E_CUDA, "no CUDA device: this library has no CPU fallback%s");

    sihmpc_plan *p = new sihmpc_plan();
    p->quad.shape = -1;
    p->mma.shape = -1;
    memset(&p->tile, 0, sizeof p->tile);
    int rc = 0;
    auto bail = [&](int code) { sihmpc_plan_destroy(p); return code; };
    if (cudaGetDevice(&p->device) != cudaSuccess) return bail(fail(SIHMPC_E_CUDA, "cudaGetDevice failed%s"));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess)
        return bail(fail(SIHMPC_E_CUDA, "cudaGetDeviceProperties failed%s"));
    p->sm_count = prop.multiProcessorCount;

    DevPlan &P = p->d;
    memset(&P, 0, sizeof P);
    P.nx = d->nx; P.nu = d->nu; P.ny = d->ny; P.nd = d->nd; P.N = d->N; P.nv = nv; P.nvm = d->N * d->nu; P.m = d->m;
    P.mh = d->mh; P.R = d->R; P.ncd = d->ncd; P.n_vin = d->n_vin; P.pl_nnz = d->pl_nnz;
    P.Ny = d->N * d->ny; P.o_sy = P.Ny; P.o_si = P.Ny + d->ny; P.o_t = P.Ny + 2 * d->ny;
    P.ld = nv | 1; P.Ts = d->Ts;
    p->nw = d->N * d->nu + 2 * d->ny;
    P.has_box = any_finite(d->dulb, d->nu) || any_finite(d->duub, d->nu);
    P.has_cum = any_finite(d->ulb, d->nu) || any_finite(d->uub, d->nu);
    P.eval_term = 0;
    for (int r = P.o_t; r < d->R && d->Wrow; ++r)
        for (int i = 0; i < d->m; ++i)
            if (d->Wrow[(size_t)r * d->m + i] != 0.0) P.eval_term = 1;

    const int R = d->R, nx = d->nx, ny = d->ny, nu = d->nu, m = d->m, Ny = P.Ny, ncd = d->ncd;
#define UP(call) do { rc = (call); if (rc) return bail(rc); } while (0)
    UP(upload(p, d->E, (size_t)R * nv, &P.E));
    UP(upload_T(p, d->E, R, nv, &P.ET));
    UP(upload_T(p, d->Px, R, nx, &P.PxT));
    UP(upload_T(p, d->Py, R, ny, &P.PyT));
    UP(upload(p, d->e0, (size_t)R, &P.e0));
    UP(upload(p, d->Wrow, (size_t)R * m, &P.Wrow));
    UP(upload_T(p, d->Wrow, R, m, &P.WrowT));
    UP(upload(p, d->w0row, (size_t)R, &P.w0row));
    UP(upload(p, d->Wdu, (size_t)nv * m, &P.Wdu));
    UP(upload_T(p, d->Wdu, nv, m, &P.WduT));
    UP(upload(p, d->HJ, (size_t)d->mh * nv * nv, &P.HJ));
    UP(upload(p, d->hj_index, (size_t)d->mh, &P.hj_index));
    UP(upload(p, d->Ht, (size_t)nv * nv, &P.Ht));
    UP(upload_T(p, d->Myx, Ny, nx, &P.MyxT));
    UP(upload_T(p, d->Myu, Ny, nv, &P.MyuT));
    UP(upload_T(p, d->Mi1x, ny, nx, &P.Mi1xT));
    UP(upload_T(p, d->Mi1u, ny, nv, &P.Mi1uT));
    UP(upload(p, d->tau, (size_t)d->N, &P.tau));
    UP(upload(p, d->gamma, (size_t)m, &P.gamma));
    UP(upload(p, d->dulb, (size_t)nu, &P.dulb));
    UP(upload(p, d->duub, (size_t)nu, &P.duub));
    UP(upload(p, d->ulb, (size_t)nu, &P.ulb));
    UP(upload(p, d->uub, (size_t)nu, &P.uub));
    UP(upload(p, d->Gc, (size_t)ncd * nv, &P.Gc));
    if (ncd) {
        UP(upload_T(p, d->Gc, ncd, nv, &P.GcT));
        UP(upload_T(p, d->Gcx, ncd, nx, &P.GcxT));
        UP(upload_T(p, d->Gcy, ncd, ny, &P.GcyT));
    } else {
        UP(upload<double>(p, nullptr, 0, &P.GcT));
        UP(upload<double>(p, nullptr, 0, &P.GcxT));
        UP(upload<double>(p, nullptr, 0, &P.GcyT));
    }
    UP(upload(p, d->gc0, (size_t)ncd, &P.gc0));
    UP(upload(p, d->Lc, (size_t)ncd, &P.Lc));
    UP(upload(p, d->Uc, (size_t)ncd, &P.Uc));
    UP(upload(p, d->gc_q, (size_t)ncd, &P.gc_q));
    UP(upload(p, d->gc_vin, (size_t)ncd, &P.gc_vin));
    UP(upload(p, d->vin_row, (size_t)d->n_vin, &P.vin_row));
    UP(upload(p, d->vin_q, (size_t)d->n_vin, &P.vin_q));
    UP(upload(p, d->pl_rowptr, (size_t)nx + 1, &P.pl_rowptr));
    UP(upload(p, d->pl_col, (size_t)d->pl_nnz, &P.pl_col));
    UP(upload(p, d->pl_val, (size_t)d->pl_nnz, &P.pl_val));
    UP(upload_T(p, d->C, ny, nx, &P.CT));
    UP(upload_T(p, d->D, ny, nu, &P.DT));
    if (d->Phi && d->Gam) {
        UP(upload_T(p, d->Phi, d->N * nx, nx, &P.PhiT));
        UP(upload_T(p, d->Gam, d->N * nx, nv, &P.GamT));
    }
#undef UP

    p->nvr = (nv + 31) / 32;
    rc = (p->nvr == 1) ? configure<1>(p) : configure<2>(p);
    if (rc) return bail(rc);
    {
        cudaError_t qe = sih::quad_configure(d, p->d, &p->quad);
        if (qe != cudaSuccess) return bail(fail(SIHMPC_E_CUDA, "fast-path configuration failed: %s", cudaGetErrorString(qe)));
        qe = sih::mma_configure(d, p->d, &p->mma);
        if (qe != cudaSuccess) return bail(fail(SIHMPC_E_CUDA, "tensor fast-path configuration failed: %s", cudaGetErrorString(qe)));
        qe = sih::tile_configure(d, p->d, &p->tile);
        if (qe != cudaSuccess) return bail(fail(SIHMPC_E_CUDA, "tile fast-path configuration failed: %s", cudaGetErrorString(qe)));
        const char *nf = getenv("SIHMPC_NO_FAST");
        const char *ff = getenv("SIHMPC_FAST");
        p->use_fast = 2;
        if (ff && !strcmp(ff, "quad")) p->use_fast = 1;
        if (ff && !strcmp(ff, "tile")) p->use_fast = 3;
        if (nf && nf[0] == '1') p->use_fast = 0;
        p->env_bounce = getenv("SIHMPC_BOUNCE") ? 1 : 0;
        p->env_no_zerocopy = getenv("SIHMPC_NO_ZEROCOPY") ? 1 : 0;
        if (const char *zv = getenv("SIHMPC_ZEROCOPY_MAX")) p->env_zc_max = atoll(zv);
        if (const char *cv = getenv("SIHMPC_HOST_CHUNKS")) { const int v = atoi(cv); if (v >= 1 && v <= 64) p->env_chunks = v; }
        if (const char *gv = getenv("SIHMPC_QUAD_GRID")) { const int v = atoi(gv); if (v > 0) p->env_grid = v; }
        p->mma.grid_cap = p->env_grid; p->quad.grid_cap = p->env_grid;
    }
    {   // host-buffer entry point: copy streams, events and the pinned pack buffer exist from here on
        const int nchunk_max = p->env_chunks > 4 ? p->env_chunks : 4;
        auto cu = [&](cudaError_t e, const char *what) { return e == cudaSuccess ? 0 : fail(SIHMPC_E_CUDA, "%s failed at plan creation", what); };
        if (cu(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(SIHMPC_E_CUDA);
        if (cu(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(SIHMPC_E_CUDA);
        for (int i = 0; i < 2 * nchunk_max + 2; ++i) {
            cudaEvent_t ev;
            if (cu(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate")) return bail(SIHMPC_E_CUDA);
            p->events.push_back(ev);
        }
        size_t off[12];
        p->h_stage_bytes = staging_layout(p, SIH_PACK_MAX, off);
        if (cu(cudaHostAlloc(&p->h_stage, p->h_stage_bytes, cudaHostAllocDefault), "cudaHostAlloc")) return bail(SIHMPC_E_CUDA);
    }
    *out = p;
    return 0;
}

size_t sihmpc_workspace_bytes(const sihmpc_plan *plan, int64_t batch) {
    if (!plan) return 0;
    if (batch < 0) batch = 0;
    return general_scratch_bytes(plan) + handover_bytes(batch) + 256;
}

const char *sihmpc_plan_kernel(const sihmpc_plan *plan) {
    if (!plan) return "none";
    if (plan->tile.ok && (plan->use_fast == 3 || (plan->use_fast == 2 && plan->mma.shape < 0))) return "tile_step_kernel";
    if (plan->use_fast == 2 && plan->mma.shape >= 0) return "mma_step_kernel";
    if (plan->use_fast == 1 && plan->quad.shape >= 0) return "quad_step_kernel";
    return "ihmpc_step_kernel";
}

int sihmpc_plan_counters(sihmpc_plan *plan, int64_t *handed_over, int64_t *solved_in_kernel, void *stream) {
    if (!plan) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaStreamSynchronize(st);
    // every counter starts a launch at the baseline 0x7f7f7f7f (one memset); a resume pass that ran counts what it
    // could not resolve in n_fail, otherwise the first pass' own counter is what the general kernel received
    const int base = 0x7f7f7f7f;
    int c[3] = {base, base, base};          // n_handover, n_fail, n_inkernel
    const int *ctr = nullptr;
    const bool want_tile = plan->tile.ok && (plan->use_fast == 3 || (plan->use_fast == 2 && plan->mma.shape < 0));
    bool resumed = false;
    int final_tile = base;
    if (want_tile) {
        ctr = plan->tile.ctr; resumed = plan->tile.last_resumed != 0;
        if (resumed && plan->tile.last_final && e == cudaSuccess)
            e = cudaMemcpy(&final_tile, plan->tile.last_final, sizeof(int), cudaMemcpyDeviceToHost);
    }
    else if (plan->use_fast == 2 && plan->mma.shape >= 0) { ctr = plan->mma.ctr; resumed = plan->mma.last_resumed != 0; }
    else if (plan->use_fast == 1 && plan->quad.shape >= 0 && plan->quad.n_handover) {
        if (e == cudaSuccess) e = cudaMemcpy(c, plan->quad.n_handover, sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess && ctr) e = cudaMemcpy(c, ctr, 3 * sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(SIHMPC_E_CUDA, "reading the fast-path counters failed: %s", cudaGetErrorString(e));
    if (want_tile && resumed && plan->tile.last_final) c[1] = final_tile;      // after the bounce rounds
    if (handed_over) *handed_over = (int64_t)(resumed ? c[1] : c[0]) - base;
    if (solved_in_kernel) *solved_in_kernel = (int64_t)c[2] - base;
    return 0;
}

int sihmpc_step(sihmpc_plan *plan, const sihmpc_io *io, int64_t batch, int32_t n_steps,
                void *workspace, size_t workspace_bytes, void *stream) {
    if (!plan || !io) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch < 0) return fail(SIHMPC_E_INVALID, "negative batch%s");
    if (batch == 0) return 0;
    int rc = check_io(plan, io, n_steps);
    if (rc) return rc;
    if (!workspace || workspace_bytes < sihmpc_workspace_bytes(plan, batch))
        return fail(SIHMPC_E_WORKSPACE, "workspace smaller than sihmpc_workspace_bytes()%s");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *scratch = reinterpret_cast<double *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    bind_handover(plan, reinterpret_cast<char *>(scratch) + general_scratch_bytes(plan), batch);
    const bool want_lam = io->lam != nullptr;     // multipliers: the dual active-set kernel reports them, no resume pass
    if (want_lam) CUDA_TRY(cudaMemsetAsync(io->lam, 0, (size_t)batch * (2 * plan->d.nv + plan->d.ncd) * sizeof(double), st));
    int *start_step = nullptr;
    const int *n_handover = nullptr;
    const bool want_tile = plan->tile.ok && (plan->use_fast == 3 || (plan->use_fast == 2 && plan->mma.shape < 0));
    if (want_tile) {
        cudaError_t e = sih::tile_reset(plan->tile, batch, st);
        if (e == cudaSuccess) e = sih::tile_launch(plan->tile, plan->d, *io, batch, n_steps, plan->sm_count, st, !want_lam);
        if (e != cudaSuccess) return fail(SIHMPC_E_CUDA, "tile fast-path launch failed: %s", cudaGetErrorString(e));
        if (!(plan->d.has_box || plan->d.has_cum)) return 0;      // nothing can be handed over
        start_step = plan->tile.start_step;
        n_handover = plan->tile.n_handover;
        if (!want_lam && sih::tile_resumes(plan->tile, batch, n_steps)) {
            // What the resume pass gave up on (large changes of the active set right after a set-point change): the
            // general kernel solves ONE step of those scenarios and sends them back with the active set it found, the
            // tile kernel resumes them from there -- TL_BOUNCES rounds, every launch of a round leaves at once when its
            // list is empty -- and the final general pass below takes what is still flagged.
            n_handover = plan->tile.n_fail;
            // Measured on the 10x10 workload (65536 x 200): three rounds resolve 23155 constrained scenario-steps
            // instead of 2437 and leave 1316 instead of 1871 scenarios to the final pass, but every round is two more
            // latency-bound launches and the launch got slower (791 vs 745 ms): opt-in (SIHMPC_BOUNCE=1).
            const int rounds = plan->env_bounce ? sih::TL_BOUNCES : 0;
            for (int r = 0; r < rounds; ++r) {
                int *nb = sih::tile_bounce_counter(plan->tile, r);
                rc = (plan->nvr == 1) ? launch_step<1>(plan, *io, batch, n_steps, scratch, st, start_step, n_handover, plan->tile.list2, nb,
                                                       plan->tile.warm_mask, 1, plan->tile.cfg.cspace)
                                      : launch_step<2>(plan, *io, batch, n_steps, scratch, st, start_step, n_handover, plan->tile.list2, nb,
                                                       plan->tile.warm_mask, 1, plan->tile.cfg.cspace);
                if (rc) return rc;
                e = sih::tile_launch_bounced(plan->tile, plan->d, *io, batch, n_steps, plan->sm_count, r, st);
                if (e != cudaSuccess) return fail(SIHMPC_E_CUDA, "tile resume launch failed: %s", cudaGetErrorString(e));
                n_handover = sih::tile_fail_counter(plan->tile, r);
            }
            plan->tile.last_final = n_handover;
        }
    } else if (plan->use_fast == 2 && plan->mma.shape >= 0) {
        cudaError_t e = sih::mma_reset(plan->mma, batch, st);
        if (e == cudaSuccess) e = sih::mma_launch(plan->mma, plan->d, *io, batch, n_steps, plan->sm_count, st, !want_lam);
        if (e != cudaSuccess) return fail(SIHMPC_E_CUDA, "tensor fast-path launch failed: %s", cudaGetErrorString(e));
        if (!(plan->d.has_box || plan->d.has_cum)) return 0;      // nothing can be handed over
        start_step = plan->mma.start_step;
        n_handover = (!want_lam && sih::mma_resumes(plan->mma, batch, n_steps)) ? plan->mma.n_fail : plan->mma.n_handover;
    } else if (plan->use_fast == 1 && sih::quad_supported(plan->quad, *io)) {
        cudaError_t e = cudaSuccess;
        if (e == cudaSuccess) e = cudaMemsetAsync(plan->quad.n_handover, 0x7f, (size_t)(batch + 1) * sizeof(int), st);
        if (e == cudaSuccess) e = sih::quad_launch(plan->quad, plan->d, *io, batch, n_steps, plan->sm_count, st);
        if (e != cudaSuccess) return fail(SIHMPC_E_CUDA, "fast-path launch failed: %s", cudaGetErrorString(e));
        if (!(plan->d.has_box || plan->d.has_cum)) return 0;      // nothing can be handed over
        start_step = plan->quad.start_step;
        n_handover = plan->quad.n_handover;
    }
    return (plan->nvr == 1) ? launch_step<1>(plan, *io, batch, n_steps, scratch, st, start_step, n_handover)
                            : launch_step<2>(plan, *io, batch, n_steps, scratch, st, start_step, n_handover);
}

// ---- host-buffer variant -----------------------------------------------------------------
static size_t staging_layout(const sihmpc_plan *p, int64_t batch, size_t off[12]) {
    const DevPlan &P = p->d;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    off[0] = take((size_t)batch * P.nx * 8);                       // x
    off[1] = take((size_t)batch * P.nu * 8);                       // u
    off[2] = take((size_t)batch * P.m * 8);                        // pesos
    off[3] = take((size_t)batch * (P.n_vin > 0 ? P.n_vin : 1) * 8);// vinant
    off[4] = take((size_t)batch * P.ny * 8);                       // ysp
    off[5] = take((size_t)batch * P.nu * 8);                       // du
    off[6] = take((size_t)batch * 8);                              // J
    off[7] = take((size_t)batch * p->nw * 8);                      // w_opt
    off[8] = take((size_t)batch * 4);                              // status
    off[9] = take((size_t)batch * 4);                              // n_active
    off[10] = take((size_t)batch * (2 * P.nv + P.ncd) * 8);        // lam
    return o;
}

size_t sihmpc_host_staging_bytes(const sihmpc_plan *plan, int64_t batch) {
    if (!plan) return 0;
    size_t off[12];
    return sihmpc_workspace_bytes(plan, batch) + staging_layout(plan, batch, off) + 512;
}

// Chunk c of the batch: element offsets of every per-scenario array.
static sihmpc_io slice_io(const sihmpc_io &d, const DevPlan &P, int nw, long long lo) {
    sihmpc_io r = d;
    r.x = d.x + lo * P.nx; r.u = d.u + lo * P.nu; r.pesos = d.pesos + lo * P.m;
    if (d.vinant) r.vinant = d.vinant + lo * P.n_vin;
    r.ysp = d.ysp + lo * P.ny;
    r.du = d.du + lo * P.nu; r.J = d.J + lo;
    if (d.w_opt) r.w_opt = d.w_opt + lo * nw;
    r.status = d.status + lo; r.n_active = d.n_active + lo;
    if (d.lam) r.lam = d.lam + lo * (2 * P.nv + P.ncd);
    return r;
}

int sihmpc_step_host(sihmpc_plan *plan, const sihmpc_io *h, int64_t batch, void *workspace,
                     size_t workspace_bytes, void *stream) {
    if (!plan || !h) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return batch == 0 ? 0 : fail(SIHMPC_E_INVALID, "negative batch%s");
    if (!h->x || !h->u || !h->pesos || !h->ysp) return fail(SIHMPC_E_INVALID, "host io needs x,u,pesos,ysp%s");
    if (!workspace || workspace_bytes < sihmpc_host_staging_bytes(plan, batch))
        return fail(SIHMPC_E_WORKSPACE, "workspace smaller than sihmpc_host_staging_bytes()%s");
    const DevPlan &P = plan->d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t off[12];
    const size_t stage_bytes = staging_layout(plan, batch, off);
    const size_t wsb = sihmpc_workspace_bytes(plan, batch);
    char *base = reinterpret_cast<char *>((((uintptr_t)workspace + wsb) + 255) & ~(uintptr_t)255);
    const bool want_vin = (P.n_vin > 0 && h->vinant);
    const bool adv = (h->flags & SIHMPC_F_ADVANCE) != 0;
    sihmpc_io d = *h;
    d.x = reinterpret_cast<double *>(base + off[0]);
    d.u = reinterpret_cast<double *>(base + off[1]);
    d.pesos = reinterpret_cast<double *>(base + off[2]);
    d.vinant = want_vin ? reinterpret_cast<double *>(base + off[3]) : nullptr;
    d.ysp = reinterpret_cast<double *>(base + off[4]);
    d.du = reinterpret_cast<double *>(base + off[5]);
    d.J = reinterpret_cast<double *>(base + off[6]);
    d.w_opt = h->w_opt ? reinterpret_cast<double *>(base + off[7]) : nullptr;
    d.status = reinterpret_cast<int32_t *>(base + off[8]);
    d.n_active = reinterpret_cast<int32_t *>(base + off[9]);
    d.lam = h->lam ? reinterpret_cast<double *>(base + off[10]) : nullptr;
    d.ysp_seq = nullptr; d.seg_start = nullptr; d.n_seg = 0;
    d.dist_step = nullptr; d.dist = nullptr; d.plant_val = nullptr;
    d.stats = nullptr; d.traj_du = nullptr; d.traj_J = nullptr; d.traj_y = nullptr;
    const size_t bx = (size_t)P.nx * 8, bu = (size_t)P.nu * 8, bp = (size_t)P.m * 8, bv = (size_t)P.n_vin * 8,
                 by = (size_t)P.ny * 8, bw = (size_t)plan->nw * 8, bl = (size_t)(2 * P.nv + P.ncd) * 8;

    if (batch <= plan->env_zc_max && !plan->env_no_zerocopy && h->du && h->J && h->status && h->n_active) {
        // ---- every buffer pinned (cudaHostAlloc / cudaHostRegister): ZERO-COPY.  The kernel reads the caller's
        // arrays over PCIe through their device aliases and writes the results straight back: no staging
        // copies, no memcpy launches -- a step is one kernel plus one synchronise (batch 64: 39 us instead of
        // 52 us; 131072 scenarios: 1.13 ms instead of 1.24 ms with the chunked DMA pipeline below).
        {
            void *hp[11] = {h->x, h->u, h->pesos, want_vin ? h->vinant : nullptr, const_cast<double *>(h->ysp),
                            h->du, h->J, h->w_opt, h->status, h->n_active, h->lam};
            void *dp[11];
            bool all = true;
            for (int i = 0; i < 11 && all; ++i) {
                dp[i] = nullptr;
                if (!hp[i]) continue;
                // asked of the driver on EVERY call: an address can be unpinned, freed and reused between two calls, so a
                // device alias remembered from an earlier call may no longer be valid (a stale one is a device fault)
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, hp[i]) != cudaSuccess) { cudaGetLastError(); all = false; break; }
                if (at.type != cudaMemoryTypeHost || !at.devicePointer) { all = false; break; }
                dp[i] = at.devicePointer;
            }
            if (all) {
                sihmpc_io z = d;
                z.x = static_cast<double *>(dp[0]); z.u = static_cast<double *>(dp[1]); z.pesos = static_cast<double *>(dp[2]);
                z.vinant = static_cast<double *>(dp[3]); z.ysp = static_cast<double *>(dp[4]);
                z.du = static_cast<double *>(dp[5]); z.J = static_cast<double *>(dp[6]); z.w_opt = static_cast<double *>(dp[7]);
                z.status = static_cast<int32_t *>(dp[8]); z.n_active = static_cast<int32_t *>(dp[9]);
                z.lam = static_cast<double *>(dp[10]);
                int rc = sihmpc_step(plan, &z, batch, 1, workspace, wsb, stream);
                if (rc) return rc;
                CUDA_TRY(cudaStreamSynchronize(st));
                return 0;
            }
        }
    }
    if (batch <= SIH_PACK_MAX && stage_bytes <= plan->h_stage_bytes) {
        // ---- small batch (latency path): pack on the host (pinned buffer of the plan), ONE upload, ONE download
        char *hs = static_cast<char *>(plan->h_stage);
        memcpy(hs + off[0], h->x, batch * bx); memcpy(hs + off[1], h->u, batch * bu);
        memcpy(hs + off[2], h->pesos, batch * bp); memcpy(hs + off[4], h->ysp, batch * by);
        if (want_vin) memcpy(hs + off[3], h->vinant, batch * bv);
        CUDA_TRY(cudaMemcpyAsync(base, hs, off[5], cudaMemcpyHostToDevice, st));
        int rc = sihmpc_step(plan, &d, batch, 1, workspace, wsb, stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(hs, base, stage_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h->du) memcpy(h->du, hs + off[5], batch * bu);
        if (h->J) memcpy(h->J, hs + off[6], batch * 8);
        if (h->w_opt) memcpy(h->w_opt, hs + off[7], batch * bw);
        if (h->status) memcpy(h->status, hs + off[8], batch * 4);
        if (h->n_active) memcpy(h->n_active, hs + off[9], batch * 4);
        if (h->lam) memcpy(h->lam, hs + off[10], batch * bl);
        if (want_vin && (h->flags & SIHMPC_F_UPDATE_VIN)) memcpy(h->vinant, hs + off[3], batch * bv);
        if (adv) { memcpy(h->x, hs + off[0], batch * bx); memcpy(h->u, hs + off[1], batch * bu); memcpy(h->pesos, hs + off[2], batch * bp); }
        return 0;
    }

    // ---- large batch (throughput path): chunks pipelined over an upload stream, the caller's
    // stream (all kernels, in order: they share the plan's hand-over buffers) and a download stream
    int nchunk = batch >= 4096 ? 4 : 2;
    if (plan->env_chunks) nchunk = plan->env_chunks;
    cudaEvent_t ev_start = plan->events[2 * nchunk], ev_done = plan->events[2 * nchunk + 1];
    CUDA_TRY(cudaEventRecord(ev_start, st));                 // uploads start after prior work of the caller's stream
    CUDA_TRY(cudaStreamWaitEvent(plan->s_h2d, ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(plan->s_d2h, ev_start, 0));
    const long long per = (batch + nchunk - 1) / nchunk;
    for (int c = 0; c < nchunk; ++c) {
        const long long lo = (long long)c * per, hi = (lo + per < batch) ? lo + per : batch;
        if (lo >= hi) break;
        const size_t n = (size_t)(hi - lo);
        cudaStream_t su = plan->s_h2d, sd = plan->s_d2h;
        CUDA_TRY(cudaMemcpyAsync(d.x + lo * P.nx, h->x + lo * P.nx, n * bx, cudaMemcpyHostToDevice, su));
        CUDA_TRY(cudaMemcpyAsync(d.u + lo * P.nu, h->u + lo * P.nu, n * bu, cudaMemcpyHostToDevice, su));
        CUDA_TRY(cudaMemcpyAsync(d.pesos + lo * P.m, h->pesos + lo * P.m, n * bp, cudaMemcpyHostToDevice, su));
        CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(d.ysp) + lo * P.ny, h->ysp + lo * P.ny, n * by, cudaMemcpyHostToDevice, su));
        if (want_vin) CUDA_TRY(cudaMemcpyAsync(d.vinant + lo * P.n_vin, h->vinant + lo * P.n_vin, n * bv, cudaMemcpyHostToDevice, su));
        CUDA_TRY(cudaEventRecord(plan->events[2 * c], su));
        CUDA_TRY(cudaStreamWaitEvent(st, plan->events[2 * c], 0));
        sihmpc_io dc = slice_io(d, P, plan->nw, lo);
        int rc = sihmpc_step(plan, &dc, (int64_t)n, 1, workspace, wsb, stream);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(plan->events[2 * c + 1], st));
        CUDA_TRY(cudaStreamWaitEvent(sd, plan->events[2 * c + 1], 0));
        if (h->du) CUDA_TRY(cudaMemcpyAsync(h->du + lo * P.nu, d.du + lo * P.nu, n * bu, cudaMemcpyDeviceToHost, sd));
        if (h->J) CUDA_TRY(cudaMemcpyAsync(h->J + lo, d.J + lo, n * 8, cudaMemcpyDeviceToHost, sd));
        if (h->w_opt) CUDA_TRY(cudaMemcpyAsync(h->w_opt + lo * plan->nw, d.w_opt + lo * plan->nw, n * bw, cudaMemcpyDeviceToHost, sd));
        if (h->status) CUDA_TRY(cudaMemcpyAsync(h->status + lo, d.status + lo, n * 4, cudaMemcpyDeviceToHost, sd));
        if (h->n_active) CUDA_TRY(cudaMemcpyAsync(h->n_active + lo, d.n_active + lo, n * 4, cudaMemcpyDeviceToHost, sd));
        if (h->lam) CUDA_TRY(cudaMemcpyAsync(h->lam + lo * (2 * P.nv + P.ncd), d.lam + lo * (2 * P.nv + P.ncd), n * bl, cudaMemcpyDeviceToHost, sd));
        if (want_vin && (h->flags & SIHMPC_F_UPDATE_VIN))
            CUDA_TRY(cudaMemcpyAsync(h->vinant + lo * P.n_vin, d.vinant + lo * P.n_vin, n * bv, cudaMemcpyDeviceToHost, sd));
        if (adv) {
            CUDA_TRY(cudaMemcpyAsync(h->x + lo * P.nx, d.x + lo * P.nx, n * bx, cudaMemcpyDeviceToHost, sd));
            CUDA_TRY(cudaMemcpyAsync(h->u + lo * P.nu, d.u + lo * P.nu, n * bu, cudaMemcpyDeviceToHost, sd));
            CUDA_TRY(cudaMemcpyAsync(h->pesos + lo * P.m, d.pesos + lo * P.m, n * bp, cudaMemcpyDeviceToHost, sd));
        }
    }
    CUDA_TRY(cudaEventRecord(ev_done, plan->s_d2h));
    CUDA_TRY(cudaStreamWaitEvent(st, ev_done, 0));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// ---- controller-resident stepping --------------------------------------------------------
struct ResidentLayout { size_t u, pesos, vinant, wstart, ysp_prev, x, ysp, du, J, w_opt, status, n_active, ws, total; };

static ResidentLayout resident_layout(const sihmpc_plan *p, int64_t batch) {
    const DevPlan &P = p->d;
    ResidentLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    const size_t B = (size_t)batch;
    L.u = take(B * P.nu * 8); L.pesos = take(B * P.m * 8); L.vinant = take(B * (P.n_vin > 0 ? P.n_vin : 1) * 8);
    L.wstart = take(B * p->nw * 8); L.ysp_prev = take(B * P.ny * 8);
    L.x = take(B * P.nx * 8); L.ysp = take(B * P.ny * 8); L.du = take(B * P.nu * 8); L.J = take(B * 8);
    L.w_opt = take(B * p->nw * 8); L.status = take(B * 4); L.n_active = take(B * 4);
    L.ws = take(sihmpc_workspace_bytes(p, batch));
    L.total = o + 256;
    return L;
}

size_t sihmpc_resident_bytes(const sihmpc_plan *plan, int64_t batch) {
    if (!plan || batch <= 0) return 0;
    return resident_layout(plan, batch).total;
}

int sihmpc_resident_weights_offset(const sihmpc_plan *plan, int64_t batch, size_t *offset) {
    if (!plan || !offset || batch <= 0) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    *offset = resident_layout(plan, batch).pesos;
    return 0;
}

int sihmpc_resident_init(sihmpc_plan *plan, void *state, size_t state_bytes, int64_t batch, const double *u0,
                         const double *pesos0, void *stream) {
    if (!plan || !state || !u0 || !pesos0) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return fail(SIHMPC_E_INVALID, "batch must be positive%s");
    if (plan->d.nv != plan->d.nvm) return fail(SIHMPC_E_INVALID, "resident stepping: ranged terminal residuals are not carried%s");
    const ResidentLayout L = resident_layout(plan, batch);
    if (state_bytes < L.total) return fail(SIHMPC_E_WORKSPACE, "state smaller than sihmpc_resident_bytes()%s");
    const DevPlan &P = plan->d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *base = reinterpret_cast<char *>(((uintptr_t)state + 255) & ~(uintptr_t)255);
    CUDA_TRY(cudaMemcpyAsync(base + L.u, u0, (size_t)batch * P.nu * 8, cudaMemcpyHostToDevice, st));
    std::vector<double> pes((size_t)batch * P.m), vin((size_t)batch * (P.n_vin > 0 ? P.n_vin : 1), 1.0 / 0.0);
    for (int64_t b = 0; b < batch; ++b)
        for (int i = 0; i < P.m; ++i) pes[(size_t)b * P.m + i] = pesos0[i];
    CUDA_TRY(cudaMemcpyAsync(base + L.pesos, pes.data(), pes.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(base + L.vinant, vin.data(), vin.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(base + L.wstart, 0, (size_t)batch * plan->nw * 8, st));
    CUDA_TRY(cudaMemsetAsync(base + L.ysp_prev, 0, (size_t)batch * P.ny * 8, st));
    CUDA_TRY(cudaStreamSynchronize(st));           // pes / vin are stack-lifetime host buffers
    plan->resident_steps = 0;
    plan->resident_batch = batch;
    return 0;
}

int sihmpc_resident_step(sihmpc_plan *plan, void *state, size_t state_bytes, const sihmpc_resident_io *h,
                         int64_t batch, void *stream) {
    if (!plan || !state || !h) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (!h->x || !h->ysp || !h->du || !h->J) return fail(SIHMPC_E_INVALID, "resident io needs x, ysp, du, J%s");
    if (plan->resident_steps < 0 || plan->resident_batch != batch)
        return fail(SIHMPC_E_INVALID, "sihmpc_resident_init has not been called for this batch%s");
    const ResidentLayout L = resident_layout(plan, batch);
    if (state_bytes < L.total) return fail(SIHMPC_E_WORKSPACE, "state smaller than sihmpc_resident_bytes()%s");
    const DevPlan &P = plan->d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *base = reinterpret_cast<char *>(((uintptr_t)state + 255) & ~(uintptr_t)255);
    auto D = [&](size_t off) { return reinterpret_cast<double *>(base + off); };
    sihmpc_io d;
    memset(&d, 0, sizeof d);
    d.x = D(L.x); d.u = D(L.u); d.pesos = D(L.pesos); d.vinant = P.n_vin > 0 ? D(L.vinant) : nullptr;
    d.ysp = D(L.ysp); d.du = D(L.du); d.J = D(L.J); d.w_opt = D(L.w_opt);
    d.status = reinterpret_cast<int32_t *>(base + L.status); d.n_active = reinterpret_cast<int32_t *>(base + L.n_active);
    d.flags = h->flags & SIHMPC_F_UPDATE_VIN;      // the solve only: plant, warm start and weights are the caller's / ours below
    d.alfa = h->alfa;
    const bool first = plan->resident_steps == 0;
    const int wflags = h->flags & (SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1);
    const size_t per_w = (size_t)(P.nx + P.nv + P.ny + 2 * P.ny + P.R + P.nv + P.m + 8) * 8;
    const int wpb = 4;
    if (per_w * wpb > 200 * 1024) return fail(SIHMPC_E_INVALID, "controller too large%s");
    // constants of the weight update staged in shared memory when they are small (not with terminal rows in a
    // weighted sub-objective: those read Px / E as well)
    const size_t cst_w = (size_t)(P.nx * P.Ny + P.nv * P.Ny + (P.Ny + 2 * P.ny) * P.m + P.nv * P.m + P.N + P.m) * 8;
    const int stage_w = (!P.eval_term && cst_w <= 48 * 1024) ? 1 : 0;
    const size_t smem_w = per_w * wpb + (stage_w ? cst_w : 0);
    CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(sih::resident_weights_kernel), smem_w));

    int nchunk = batch >= 16384 ? 4 : 1;
    if (plan->env_chunks) nchunk = plan->env_chunks;
    if (2 * nchunk + 2 > (int)plan->events.size()) nchunk = ((int)plan->events.size() - 2) / 2;
    cudaEvent_t ev_start = plan->events[2 * nchunk], ev_done = plan->events[2 * nchunk + 1];
    CUDA_TRY(cudaEventRecord(ev_start, st));
    CUDA_TRY(cudaStreamWaitEvent(plan->s_h2d, ev_start, 0));
    CUDA_TRY(cudaStreamWaitEvent(plan->s_d2h, ev_start, 0));
    const long long per = (batch + nchunk - 1) / nchunk;
    const size_t wsb = sihmpc_workspace_bytes(plan, batch);
    for (int c = 0; c < nchunk; ++c) {
        const long long lo = (long long)c * per, hi = (lo + per < batch) ? lo + per : batch;
        if (lo >= hi) break;
        const long long n = hi - lo;
        CUDA_TRY(cudaMemcpyAsync(D(L.x) + lo * P.nx, h->x + lo * P.nx, (size_t)n * P.nx * 8, cudaMemcpyHostToDevice, plan->s_h2d));
        CUDA_TRY(cudaMemcpyAsync(D(L.ysp) + lo * P.ny, h->ysp + lo * P.ny, (size_t)n * P.ny * 8, cudaMemcpyHostToDevice, plan->s_h2d));
        CUDA_TRY(cudaEventRecord(plan->events[2 * c], plan->s_h2d));
        CUDA_TRY(cudaStreamWaitEvent(st, plan->events[2 * c], 0));
        const bool in_kernel_weights = plan->use_fast == 2 && plan->mma.shape >= 0;      // MODE 3 of the tensor fast path
        plan->mma.res_active = 0;
        if (in_kernel_weights) {
            plan->mma.res_active = 1;
            plan->mma.res_flags = first ? 0 : wflags;
            plan->mma.res_wstart = D(L.wstart) + lo * plan->nw;
            plan->mma.res_ysp_prev = D(L.ysp_prev) + lo * P.ny;
        } else if (!first && wflags) {
            long long blocks = (n + wpb - 1) / wpb;
            if (blocks > plan->sm_count * 8) blocks = plan->sm_count * 8;
            sih::resident_weights_kernel<<<(int)blocks, wpb * 32, smem_w, st>>>(
                P, D(L.x) + lo * P.nx, D(L.wstart) + lo * plan->nw, D(L.ysp_prev) + lo * P.ny, D(L.pesos) + lo * P.m, n, wflags,
                h->alfa, stage_w);
            CUDA_TRY(cudaGetLastError());
        }
        sihmpc_io dc = slice_io(d, P, plan->nw, lo);
        int rc = sihmpc_step(plan, &dc, (int64_t)n, 1, base + L.ws, wsb, stream);
        plan->mma.res_active = 0;
        if (rc) return rc;
        {
            long long blocks = (n + 3) / 4;                      // one warp per scenario, four per CTA
            if (blocks > plan->sm_count * 16) blocks = plan->sm_count * 16;
            sih::resident_advance_kernel<<<(int)blocks, 128, 0, st>>>(
                P, D(L.x) + lo * P.nx, D(L.w_opt) + lo * plan->nw, D(L.ysp) + lo * P.ny, D(L.u) + lo * P.nu,
                D(L.wstart) + lo * plan->nw, D(L.ysp_prev) + lo * P.ny, n);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(plan->events[2 * c + 1], st));
        CUDA_TRY(cudaStreamWaitEvent(plan->s_d2h, plan->events[2 * c + 1], 0));
        CUDA_TRY(cudaMemcpyAsync(h->du + lo * P.nu, D(L.du) + lo * P.nu, (size_t)n * P.nu * 8, cudaMemcpyDeviceToHost, plan->s_d2h));
        CUDA_TRY(cudaMemcpyAsync(h->J + lo, D(L.J) + lo, (size_t)n * 8, cudaMemcpyDeviceToHost, plan->s_d2h));
        if (h->status)
            CUDA_TRY(cudaMemcpyAsync(h->status + lo, reinterpret_cast<int32_t *>(base + L.status) + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, plan->s_d2h));
    }
    CUDA_TRY(cudaEventRecord(ev_done, plan->s_d2h));
    CUDA_TRY(cudaStreamWaitEvent(st, ev_done, 0));
    CUDA_TRY(cudaStreamSynchronize(st));
    ++plan->resident_steps;
    return 0;
}

int sihmpc_predict(sihmpc_plan *plan, const double *x, const double *u, const double *dU,
                   int64_t batch, double *x_pred, double *y_pred, double *u_pred, void *stream) {
    if (!plan || !x || !u || !dU) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return batch == 0 ? 0 : fail(SIHMPC_E_INVALID, "negative batch%s");
    if (x_pred && !plan->d.PhiT) return fail(SIHMPC_E_INVALID, "plan was created without Phi/Gam%s");
    const int threads = 128;
    long long blocks = (batch * 32 + threads - 1) / threads;
    if (blocks > plan->sm_count * 16) blocks = plan->sm_count * 16;
    sih::predict_kernel<<<(int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        plan->d, x, u, dU, batch, x_pred, y_pred, u_pred);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sihmpc_subobjectives(sihmpc_plan *plan, const double *x, const double *w, const double *ysp,
                         int64_t batch, double *V, void *stream) {
    if (!plan || !x || !w || !ysp || !V) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return batch == 0 ? 0 : fail(SIHMPC_E_INVALID, "negative batch%s");
    const DevPlan &P = plan->d;
    // w is [N*nu ; syN ; siN] (sol['w_opt']): a plan with ranged terminal residuals carries them as extra decision
    // variables, so syN/siN are no longer functions of that vector alone -- evaluate through a plan without ranges
    if (P.nv != P.nvm) return fail(SIHMPC_E_INVALID, "sihmpc_subobjectives: plan has ranged terminal residuals (n_extra > 0)%s");
    const int wpb = 4;
    const size_t per = (size_t)(P.nx + P.nv + P.ny + 2 * P.ny + P.R + P.nv + P.m + 8) * 8;
    const size_t smem = per * wpb;
    if (smem > 200 * 1024) return fail(SIHMPC_E_INVALID, "controller too large%s");
    CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(sih::subobjectives_kernel), smem));
    long long blocks = (batch + wpb - 1) / wpb;
    if (blocks > plan->sm_count * 8) blocks = plan->sm_count * 8;
    sih::subobjectives_kernel<<<(int)blocks, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        plan->d, x, w, ysp, batch, V);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sihmpc_stats_reduce(const double *stats, int64_t batch, int32_t nstat, double *out, void *stream) {
    if (!stats || !out || nstat < 1) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sih::stats_init_kernel<<<1, 128, 0, st>>>(out, nstat);       // sums 0, max -1e308: no host buffer, no synchronise
    CUDA_TRY(cudaGetLastError());
    if (batch <= 0) return 0;
    long long bx = (batch + 255) / 256;
    if (bx > 256) bx = 256;
    dim3 grid((unsigned)bx, (unsigned)nstat);
    sih::stats_reduce_kernel<<<grid, 256, 0, st>>>(stats, batch, nstat, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int log_launch(sihmpc_plan *plan, const sihmpc_io &io, const double *w0, double *w_warm, int64_t batch,
                      int32_t n_steps, int32_t *iters, void *stream) {
    const sih::LogLayout lay = sih::make_log_layout(plan->d);
    int wpb = 4;
    while (wpb > 1 && (size_t)lay.total * 8 * wpb > 200 * 1024) --wpb;
    const size_t smem = (size_t)lay.total * 8 * wpb;
    if (smem > 220 * 1024) return fail(SIHMPC_E_INVALID, "controller too large for the log kernel%s");
    long long blocks = (batch + wpb - 1) / wpb;
    if (blocks > plan->sm_count * 4) blocks = plan->sm_count * 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // <rows per lane, compile-time number of sub-objectives (0 = any)>: m = 4 is the reference's example and the
    // distillation controller (four sub-objectives, example_distillation_log.py:55-66)
    if (plan->d.nv + 1 <= 32 && plan->d.m == 4) {
        CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(sih::ihmpc_log_kernel<1, 4>), smem));
        sih::ihmpc_log_kernel<1, 4><<<(int)blocks, wpb * 32, smem, st>>>(plan->d, io, w0, w_warm, batch, n_steps, iters, 100);
    } else if (plan->d.nv + 1 <= 32) {
        CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(sih::ihmpc_log_kernel<1, 0>), smem));
        sih::ihmpc_log_kernel<1, 0><<<(int)blocks, wpb * 32, smem, st>>>(plan->d, io, w0, w_warm, batch, n_steps, iters, 100);
    } else {
        CUDA_TRY(sih_raise_smem_limit(reinterpret_cast<const void *>(sih::ihmpc_log_kernel<2, 0>), smem));
        sih::ihmpc_log_kernel<2, 0><<<(int)blocks, wpb * 32, smem, st>>>(plan->d, io, w0, w_warm, batch, n_steps, iters, 100);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sihmpc_log_step(sihmpc_plan *plan, const double *x, const double *u, const double *ysp, const double *w0,
                    int64_t batch, double *w_opt, double *J, int32_t *status, int32_t *iters, void *stream) {
    if (!plan || !x || !u || !ysp || !w_opt) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return batch == 0 ? 0 : fail(SIHMPC_E_INVALID, "negative batch%s");
    sihmpc_io io;
    memset(&io, 0, sizeof io);
    io.x = const_cast<double *>(x); io.u = const_cast<double *>(u); io.ysp = ysp;    // read only: no SIHMPC_F_ADVANCE
    io.w_opt = w_opt; io.J = J; io.status = status;
    return log_launch(plan, io, w0, nullptr, batch, 1, iters, stream);
}

int sihmpc_log_loop(sihmpc_plan *plan, const sihmpc_io *io, double *w_warm, int64_t batch, int32_t n_steps,
                    int32_t *iters, void *stream) {
    if (!plan || !io) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (batch <= 0) return batch == 0 ? 0 : fail(SIHMPC_E_INVALID, "negative batch%s");
    if (!io->x || !io->u) return fail(SIHMPC_E_INVALID, "io.x and io.u are required%s");
    if (!io->ysp && !(io->ysp_seq && io->seg_start && io->n_seg > 0))
        return fail(SIHMPC_E_INVALID, "io.ysp or io.ysp_seq/seg_start is required%s");
    if (n_steps < 1) return fail(SIHMPC_E_INVALID, "n_steps must be >= 1%s");
    if (n_steps > 1 && !(io->flags & SIHMPC_F_ADVANCE)) return fail(SIHMPC_E_INVALID, "n_steps > 1 needs SIHMPC_F_ADVANCE%s");
    return log_launch(plan, *io, w_warm, w_warm, batch, n_steps, iters, stream);
}

int sihmpc_fp64_peak(double *flops_per_s) {
    if (!flops_per_s) return fail(SIHMPC_E_INVALID, "NULL argument%s");
    if (sihmpc_device_count() < 1) return fail(SIHMPC_E_CUDA, "no CUDA device%s");
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    double *out = nullptr;
    CUDA_TRY(cudaMalloc(&out, 8));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        sih::fp64_probe_kernel<<<blocks, threads>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3);
        if (rep > 0 && fl > best) best = fl;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *flops_per_s = best;
    return 0;
}

}  // extern "C"
