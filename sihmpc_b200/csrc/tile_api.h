// tile_api.h -- host interface of the tile fast path v4 (tile_kernel.cu), see there.
#pragma once
#include <cuda_runtime.h>
#include "devplan.h"

namespace sih {

constexpr int TL_WARPS = 8;          // == scenarios per CTA == M of the MMA
constexpr int TL_NTMAX = 8;          // tile rows of the augmented matrix: nv + 1 <= 64
constexpr int TL_BOUNCES = 3;        // rounds of (general kernel: one hard step) -> (resume pass with its active set)
constexpr int TL_CTR = 4 + 2 * TL_BOUNCES + (2 * TL_BOUNCES) % 4;      // counter words in front of start_step[]
constexpr int TL_KHMAX = 6;          // k-tiles of the Hessian product: 1 + max(hj_index) + 1 <= 24

struct TileCfg {
    const int *ptr;                  // per matrix: stored tiles [ptr[T], ptr[T+1]) of output tile T (a multiple of four)
    const int *kof;                  // first k (A offset in doubles) of every stored tile
    const int *hrows;                // Hessian product: real rows (0..8) of packed tile t
    const double *B;                 // stored tiles, 32 doubles each, B-fragment order
    const double *BH;                // Hessian constants, dense: [tile][row][k-tile][32]
    const double *BHc;               // the same in the bound space of the in-kernel active set: S^T (Ht | HJ_a + 2 diag Wdu_a) S
    int KH, ntile;                   // k-tiles of [1 ; pesos] that carry Hessian blocks; packed tiles
    int KHc;                         // k-tiles of BHc (every weight carries a block there)
    int cspace;                      // 1: bound space = cumulative moves (input box), 0: moves (move box), -1: no in-kernel active set
    int me, mw, mdu, mf, mh, mr, my, myp, mvn, mpl;  // offsets of the row-pointer arrays in ptr[]
    int n_x;                         // output tiles of the shared-plant product (0: CSR path)
    int n_r, n_v, n_f, n_h, n_y, n_yp, n_m;          // output tile counts
    int NT, NP, Ib, rb, HS;
    int S_sv, S_r, S_v, S_p, S_q, S_m, S_sc;         // per-scenario strides (doubles; == 4 mod 16 where read as A)
    int o_H, o_sv, o_e, o_om, o_t, o_omdu, o_pes, o_sq, o_vn, o_z, o_y, o_sc;
    int total;                       // doubles of shared memory per CTA
};

struct TilePlan {
    int ok;                  // 1: the tile path can run this controller
    TileCfg cfg;
    int *ibuf;               // device: ptr | kt | hmap
    double *dbuf;            // device: stored tiles
    int *ctr;                // in the caller's workspace (tile_bind): [n_handover | n_fail | n_inkernel | pad | (n_bounce, n_fail2) x TL_BOUNCES | start_step[cap] | list[cap] | list2[cap]]
    int *start_step;
    int *n_handover;         // scenarios the first pass appended to `list` (counter baseline 0x7f7f7f7f)
    int *n_fail;             // scenarios the resume pass could not resolve: left to the general kernel
    int *n_inkernel;         // scenario-steps with active bounds resolved by the in-kernel active set (last launch)
    int *list, *list2;       // scenarios listed by the first pass / sent back by the general kernel (bounce rounds)
    unsigned long long *warm_mask;   // [cap][2] active set exported by the general kernel for a scenario it sends back
    int last_resumed;        // the last launch ran the resume pass (sihmpc_plan_counters)
    const int *last_final;   // counter the final general pass of the last launch was given
    long long capacity;
    size_t smem_bytes;
    int n_tiles;
};

cudaError_t tile_configure(const sihmpc_plan_desc *d, const DevPlan &P, TilePlan *tp);
void tile_bind(TilePlan *tp, int *ctr, unsigned long long *warm, long long batch);
cudaError_t tile_reset(const TilePlan &tp, long long batch, cudaStream_t st);
// the resume pass is one more launch: only worth it when the call carries enough work (not the single-step host API)
inline bool tile_resumes(const TilePlan &tp, long long batch, int n_steps) { return tp.cfg.cspace >= 0 && (n_steps > 1 || batch >= 4096); }
cudaError_t tile_launch(TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch, int n_steps,
                        int sm_count, cudaStream_t st, bool allow_resume = true);
inline int *tile_bounce_counter(const TilePlan &tp, int round) { return tp.ctr + 4 + 2 * round; }
inline int *tile_fail_counter(const TilePlan &tp, int round) { return tp.ctr + 4 + 2 * round + 1; }
cudaError_t tile_launch_bounced(const TilePlan &tp, const DevPlan &P, const sihmpc_io &io, long long batch, int n_steps,
                                int sm_count, int round, cudaStream_t st);
void tile_destroy(TilePlan *tp);

}  // namespace sih
