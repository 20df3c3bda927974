// mma_api.h -- host interface of the tensor fast path v3 (mma_kernel.cuh), see there.  The kernels are instantiated
// per (horizon, state capacity) in their own translation units (mma_units.cu, built in parallel).
#pragma once
#include <cuda_runtime.h>
#include "devplan.h"

namespace sih {

struct MmaConst {            // offsets in doubles into the packed block
    int be, bw, w0c, bf, br, bh, bomdu, by, byp, bvn, tau, gamma, inv_gamma, bounds, vinq, ell_val, total_d;
    int ell_col, ell_csr, hj_idx, vin_row, vlive, vdead, total_i;
    int ell_w;
    // in-kernel active set on simple bounds (mma_active_set): gradient / Hessian / difference maps in the space
    // where the bounds are simple (cspace 1: cumulative moves c = cumsum(dU), input box; 0: dU itself, move box;
    // -1: not available, a violated bound hands the scenario over)
    int bfc, bhc, bdp, bdpt, du_idx, n_du, cspace;
    // input-box check of the unconstrained optimum as one product: cumulative moves at the live positions (bcum);
    // cum0: bit j = the first move of input j is presolved, i.e. the cumulative move 0 belongs to the checked set
    int bcum, cum0;
};

struct MmaPlan {
    int shape;               // >= 0: served by the tensor fast path (index of the compiled unit), -1: not
    int N, nxp, nve;         // horizon, state capacity (multiple of 8) and live variables of the instantiation
    MmaConst qc;
    double *qbuf;
    int *ctr;                // in the caller's workspace (mma_bind): [n_handover | n_fail | n_inkernel | first-pass CTAs done | entries claimed | pad x3 | start_step[cap] | list[cap]]
    int grid_cap;            // debug: upper limit of the grid (0: none)
    int *start_step;
    int *n_handover;         // scenarios the first pass appended to `list` (counter baseline 0x7f7f7f7f)
    int *n_fail;             // scenarios the resume pass could not resolve either: left to the general kernel
    int *n_inkernel;         // scenario-steps of the last launch whose active bounds were resolved by mma_active_set
    int *list;
    int last_resumed;        // the last launch ran the resume pass (sihmpc_plan_counters)
    long long capacity;
    int blocks_per_sm;
    size_t smem_bytes;
    // controller-resident stepping (sihmpc_resident_step): set for the launches of one call, first pass = MODE 3
    int res_active, res_flags;                 // weight rule to apply before the solve (0: first call, none)
    const double *res_wstart, *res_ysp_prev;   // warm start of the previous solve, the set-point it was made for
    // streaming resume (MODE 4): workers on an auxiliary stream consume the hand-over list while the first pass runs
    cudaStream_t aux;
    cudaEvent_t ev_fork, ev_join;
    int stream_workers;      // CTAs of the worker launch (0: resume pass after the first pass, as in round 1)
    long long stream_min_batch;   // below this batch the resume pass simply follows the first pass: the workers only pay when the
                                  // first pass is long, and their launch + fork/join are a third of a batch-1 step
    int sm_count;
};


// shape family compiled in (2x2 models with one pole per channel): N*nu in {12, 16, 20, 24}, nx <= 32, four weighted
// sub-objectives; `nve` live variables after the presolve of moves without effect inside the horizon
cudaError_t mma_configure(const sihmpc_plan_desc *d, const DevPlan &P, MmaPlan *qp);
cudaError_t mma_launch(MmaPlan &qp, const DevPlan &P, const sihmpc_io &io, long long batch, int n_steps, int sm_count,
                       cudaStream_t st, bool allow_resume = true);
void mma_destroy(MmaPlan *qp);

// the hand-over block lives in the caller's workspace (capi.cu: bind_handover): nothing is allocated at step time
inline void mma_bind(MmaPlan *qp, int *ctr, long long batch) {
    qp->ctr = ctr; qp->n_handover = ctr; qp->n_fail = ctr + 1; qp->n_inkernel = ctr + 2;
    qp->start_step = ctr + 8; qp->list = ctr + 8 + batch; qp->capacity = batch;
}
// counters and start_step back to "nothing handed over" (every counter counts up from the baseline 0x7f7f7f7f)
inline cudaError_t mma_reset(const MmaPlan &qp, long long batch, cudaStream_t st) {
    return cudaMemsetAsync(qp.ctr, 0x7f, (size_t)(8 + 2 * batch) * sizeof(int), st);      // counters, start_step[] AND list[] (entry = "not published")
}
// the resume pass is one more launch: only worth it when the call carries enough work (not the single-step host API)
inline bool mma_resumes(const MmaPlan &qp, long long batch, int n_steps) { return qp.qc.cspace >= 0 && (n_steps > 1 || batch >= 4096); }

}  // namespace sih
