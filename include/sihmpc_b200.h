/* sihmpc_b200.h -- C ABI of the B200-native batched IHMPC step engine.
 *
 * The reference (mllima/sihmpc) is pure Python and has NO FFI of its own: its boundary for the
 * per-timestep hot path is the Python class surface of sihmpc/ihmpccontroller.py.  Each entry
 * point below names the reference code it replaces (paths relative to /root/reference).  The
 * Python mirror of that surface (sihmpc_b200/ihmpccontroller.py) binds these with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions: plain C, no torch types.  All matrices are FP64, row-major.  "Device" pointers
 * are CUDA device memory owned by the caller on the plan's device.  Per-scenario arrays are
 * [batch][feature] (scenario-major).  Every call returns 0 on success or a negative
 * SIHMPC_E_* code, never throws, never synchronises the stream unless stated, and never
 * allocates after plan creation: the caller supplies the workspace (sihmpc_workspace_bytes /
 * sihmpc_host_staging_bytes for the batch it steps), which also holds the hand-over lists and
 * counters of the fast kernels; copy streams, events and the pinned pack buffer of the host-buffer
 * entry point are created by sihmpc_plan_create.  (sihmpc_fp64_peak, a measurement helper, is the
 * exception.)  Environment overrides (SIHMPC_FAST, SIHMPC_NO_FAST, SIHMPC_BOUNCE, SIHMPC_NO_ZEROCOPY,
 * SIHMPC_ZEROCOPY_MAX, SIHMPC_HOST_CHUNKS: A/B measurements only) are read once, by
 * sihmpc_plan_create.  A plan is bound to the
 * device that was current at creation and is not thread-safe; distinct plans may be used from
 * distinct threads.  There is no CPU fallback: without a CUDA device every compute entry point
 * fails with SIHMPC_E_CUDA.
 */
#ifndef SIHMPC_B200_H
#define SIHMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIHMPC_ABI_VERSION 3

/* error codes */
#define SIHMPC_OK            0
#define SIHMPC_E_INVALID    -1   /* bad argument / unsupported shape            */
#define SIHMPC_E_CUDA       -2   /* CUDA runtime error (see sihmpc_last_error) */
#define SIHMPC_E_WORKSPACE  -3   /* workspace too small                        */
#define SIHMPC_E_NOMEM      -4

/* per-scenario status word written by the step (bit set) */
#define SIHMPC_ST_OK          0
#define SIHMPC_ST_MAXITER     1  /* active-set iteration cap hit              */
#define SIHMPC_ST_INFEASIBLE  2  /* inequality rows inconsistent              */
#define SIHMPC_ST_NOTPD       4  /* Hessian pivot <= 0 (no positive du weight) */

/* sihmpc_io.flags */
#define SIHMPC_F_ADVANCE      1  /* after the solve apply the first move to the plant, shift the
                                    plan and update the weights in place: the driver-loop body
                                    example_distillation.py:98-113 (dynF, warmStart, satWeights2) */
#define SIHMPC_F_UPDATE_VIN   2  /* store ViNant = q*siN_next^2 (ihmpccontroller.py:581-584)     */
#define SIHMPC_F_WEIGHTS_SAT2 4  /* weight rule satWeights2 (ihmpccontroller.py:619-636)         */
#define SIHMPC_F_WEIGHTS_SAT1 8  /* weight rule satWeights  (ihmpccontroller.py:607-617)         */

/* number of per-scenario closed-loop statistics accumulated in sihmpc_io.stats:
 *   [0..ny)  ISE per output  sum_k (y_k - ysp_k)^2      [ny]   sum_k |du_k|^2
 *   [ny+1]   max inequality violation seen              [ny+2] last J
 *   [ny+3]   OR of status words                         [ny+4] total active-set iterations      */
#define SIHMPC_NSTATS_EXTRA 5

typedef struct sihmpc_plan sihmpc_plan;

/* Host-side description of one compiled controller.  Replaces the per-step rebuild
 * IHMPCController._OptimProbl/_MPC (sihmpc/ihmpccontroller.py:374-531) and the symbolic
 * prediction (ihmpccontroller.py:346-371): the blocks are the constant parts of
 *     r(dU) = E dU + Px x + Py ysp + e0 ,  J = sum_r omega_r r_r^2 + sum_j omega_du_j dU_j^2 ,
 *     omega = Wrow pesos + w0row ,  omega_du = Wdu pesos ,
 *     H = Ht + sum_a pesos[hj_index[a]] HJ[a] + 2 diag(omega_du) ,  f = 2 E^T (omega o r(0)).
 * All pointers are HOST memory, copied to the device by sihmpc_plan_create. */
typedef struct sihmpc_plan_desc {
    int32_t abi_version;
    int32_t nx, nu, ny, nd, N;       /* model and horizon (ihmpccontroller.py:18-34)              */
    int32_t m;                       /* len(VJ): weighted sub-objectives (ihmpccontroller.py:68)  */
    int32_t mh;                      /* how many of them carry a dense Hessian block             */
    int32_t R;                       /* residual rows = N*ny + 2*ny + nd                         */
    int32_t ncd;                     /* dense inequality rows                                    */
    int32_t n_vin;                   /* siN contraction slots (ihmpccontroller.py:265-269)       */
    int32_t pl_nnz;                  /* non-zeros of the plant map [A | B]                       */
    double  Ts;
    const double *E;                 /* [R][nv]                                                  */
    const double *Px;                /* [R][nx]                                                  */
    const double *Py;                /* [R][ny]                                                  */
    const double *e0;                /* [R]                                                      */
    const double *Wrow;              /* [R][m]                                                   */
    const double *w0row;             /* [R]                                                      */
    const double *Wdu;               /* [nv][m]                                                  */
    const double *HJ;                /* [mh][nv][nv]                                             */
    const int32_t *hj_index;         /* [mh]                                                     */
    const double *Ht;                /* [nv][nv]                                                 */
    const double *Myx;               /* [N*ny][nx]  Y_pred = Myx x + Myu dU                      */
    const double *Myu;               /* [N*ny][nv]                                               */
    const double *Mi1x;              /* [ny][nx]    xi(N+1), warm start (ihmpccontroller.py:543-548) */
    const double *Mi1u;              /* [ny][nv]                                                 */
    const double *tau;               /* [N]  (k+1-N)*Ts (ihmpccontroller.py:174)                 */
    const double *gamma;             /* [m]  satisficing limits                                  */
    const double *dulb, *duub;       /* [nu] move box (ihmpccontroller.py:42-43)                 */
    const double *ulb, *uub;         /* [nu] input box (ihmpccontroller.py:40-41)                */
    const double *Gc;                /* [ncd][nv] dense rows: state/slack boxes, contraction     */
    const double *Gcx;               /* [ncd][nx]                                                */
    const double *Gcy;               /* [ncd][ny]                                                */
    const double *gc0, *Lc, *Uc;     /* [ncd]                                                    */
    const int32_t *gc_vin;           /* [ncd] contraction slot or -1                             */
    const double *gc_q;              /* [ncd]                                                    */
    const int32_t *vin_row;          /* [n_vin] residual row holding siN of that slot            */
    const double *vin_q;             /* [n_vin]                                                  */
    const int32_t *pl_rowptr;        /* [nx+1]   CSR of [A | B] (ihmpccontroller.py:101-107)     */
    const int32_t *pl_col;           /* [pl_nnz]                                                 */
    const double *pl_val;            /* [pl_nnz] nominal plant                                   */
    const double *C;                 /* [ny][nx]                                                 */
    const double *D;                 /* [ny][nu]                                                 */
    const double *Phi;               /* [N*nx][nx]  X_pred = Phi x + Gam dU; NULL: sihmpc_predict */
    const double *Gam;               /* [N*nx][nv]  cannot return x_pred                          */
    int32_t n_extra;                 /* terminal residual RANGES (rslb != rsub / rilb != riub, ihmpccontroller.py:441-449):
                                        one decision variable per ranged residual after the N*nu moves; every `nv`
                                        above is then N*nu + n_extra wide, the range itself is a dense row (Gc).  0: none */
} sihmpc_plan_desc;

/* Per-call I/O.  DEVICE pointers; [batch][feature] FP64 unless noted; NULL = not wanted. */
typedef struct sihmpc_io {
    /* controller/plant state, read every step, written when SIHMPC_F_ADVANCE is set */
    double *x;                       /* [batch][nx]  x0 of mpc() (ihmpccontroller.py:567)        */
    double *u;                       /* [batch][nu]  u0                                          */
    double *pesos;                   /* [batch][m]   weights                                     */
    double *vinant;                  /* [batch][n_vin] ViN_ant, +inf = not yet set               */
    /* set-point: either fixed for the call ... */
    const double *ysp;               /* [batch][ny]                                              */
    /* ... or a per-scenario piecewise-constant schedule (example_distillation.py:86-90):
       segment s applies at absolute steps k >= seg_start[b][s]; seg_start[b][0] must be <= k0  */
    const double *ysp_seq;           /* [batch][n_seg][ny]                                       */
    const int32_t *seg_start;        /* [batch][n_seg]                                           */
    int32_t n_seg;
    /* unmeasured output-step disturbance: dist[b] is added to xs after the plant update of
       absolute step dist_step[b]                              */
    const int32_t *dist_step;        /* [batch]                                                  */
    const double *dist;              /* [batch][ny]                                              */
    /* plant-model mismatch: per-scenario values of the plant's CSR (same pattern as the model) */
    const double *plant_val;         /* [batch][pl_nnz]                                          */
    /* results of the LAST step of the call */
    double *du;                      /* [batch][nu]  first move, c.du (ihmpccontroller.py:595)   */
    double *J;                       /* [batch]      sol['J']                                    */
    double *w_opt;                   /* [batch][nv+2ny]  sol['w_opt'] = [dU ; syN ; siN]          */
    int32_t *status;                 /* [batch]                                                  */
    int32_t *n_active;               /* [batch]  active inequality rows at the optimum           */
    /* closed-loop accumulators and optional per-step traces */
    double *stats;                   /* [batch][ny+SIHMPC_NSTATS_EXTRA], accumulated in place    */
    double *traj_du;                 /* [n_steps][batch][nu]                                     */
    double *traj_J;                  /* [n_steps][batch]                                         */
    double *traj_y;                  /* [n_steps][batch][ny]  plant output after the move        */
    int32_t flags;
    int32_t k0;                      /* absolute index of the first step of this call            */
    double alfa;                     /* satWeights2 blending (ihmpccontroller.py:634)            */
    /* multipliers of the inequality rows at the optimum of the LAST step (sol['lam_w'], sol['lam_g'] of
       ihmpccontroller.py:525-529 are assembled from them), casadi's sign convention: > 0 upper bound active,
       < 0 lower bound active, 0 inactive.  Rows: [0,nv) move box dulb <= dU_i <= duub (-> lam_w), [nv,2nv)
       input box ulb <= u + cumsum(dU) <= uub, row nv + k*nu + j = U_pred[k][j] (-> lam_g), [2nv,2nv+ncd) dense
       rows in the order of sihmpc_plan_desc.Gc.  NULL: not wanted.  When set, scenarios with active bounds are
       solved by the dual active-set kernel (the resume passes of the fast kernels do not report multipliers). */
    double *lam;                     /* [batch][2*nv + ncd]                                      */
} sihmpc_io;

/* -- lifecycle ------------------------------------------------------------------------------ */
int sihmpc_abi_version(void);
const char *sihmpc_last_error(void);                  /* thread-local, never NULL               */
int sihmpc_device_count(void);                        /* 0 without a usable CUDA device         */

/* replaces: IHMPCController.__init__ / subObj / subObjComposed / set_terminal_objective and the _MPC() rebuild
 * (ihmpccontroller.py:12-85,151-343,480-531): uploads the compiled blocks once.                */
int sihmpc_plan_create(const sihmpc_plan_desc *desc, sihmpc_plan **out);
int sihmpc_plan_destroy(sihmpc_plan *plan);
size_t sihmpc_workspace_bytes(const sihmpc_plan *plan, int64_t batch);
/* Name of the kernel that serves sihmpc_step for this plan: "mma_step_kernel" (register quads + DMMA,
 * distillation shape), "tile_step_kernel" (tile-CSR DMMA products + tiled block LDL^T, N*nu < 64),
 * "quad_step_kernel" or "ihmpc_step_kernel" (general warp-per-scenario active set; also the second pass
 * for scenarios a fast path hands over).  Static string, never NULL. */
const char *sihmpc_plan_kernel(const sihmpc_plan *plan);

/* Diagnostics of the LAST sihmpc_step launch of this plan (synchronises `stream`): how many scenarios a fast
 * path handed over to the general active-set kernel, and how many scenario-steps with active bounds the
 * tensor fast path resolved itself (mma_active_set).  Either pointer may be NULL.                       */
int sihmpc_plan_counters(sihmpc_plan *plan, int64_t *handed_over, int64_t *solved_in_kernel, void *stream);

/* -- the hot path --------------------------------------------------------------------------- */
/* replaces: IHMPCController.mpc (ihmpccontroller.py:567-604) for `batch` scenarios at once and,
 * with SIHMPC_F_ADVANCE, the rest of the driver-loop body (example_distillation.py:98-113:
 * dynF ihmpccontroller.py:93-113, warmStart :534-564, satWeights2 :619-636), for n_steps
 * consecutive control steps in ONE launch (state stays on chip between steps).
 * `stream` is a cudaStream_t passed as void*.                                                   */
int sihmpc_step(sihmpc_plan *plan, const sihmpc_io *io, int64_t batch, int32_t n_steps,
                void *workspace, size_t workspace_bytes, void *stream);

/* Same step for callers whose state lives in HOST memory (the reference's calling convention:
 * numpy in, numpy out).  Copies x,u,pesos,vinant,ysp host->device, runs one step, copies
 * du,J,w_opt,status (and the advanced state when SIHMPC_F_ADVANCE) device->host, then
 * synchronises `stream`.  io holds HOST pointers here; staging buffers come from `workspace`.
 * When every buffer is page-locked (cudaHostAlloc / cudaHostRegister, mapped) the kernel reads and
 * writes the caller's arrays in place over PCIe (zero-copy); the device aliases are asked of the
 * driver on every call, so buffers may be freed, unregistered or re-registered BETWEEN calls, but
 * must stay valid and pinned for the duration of a call.                                          */
int sihmpc_step_host(sihmpc_plan *plan, const sihmpc_io *host_io, int64_t batch,
                     void *workspace, size_t workspace_bytes, void *stream);
size_t sihmpc_host_staging_bytes(const sihmpc_plan *plan, int64_t batch);

/* -- controller-resident stepping ------------------------------------------------------------- */
/* The driver loop of example_distillation.py:81-113 with the PLANT on the caller's side (a real process, or the
 * caller's own simulator): every call takes the measured state and the set-point of this control step from HOST
 * memory and returns the move and the cost to HOST memory; everything the controller carries from one step to the
 * next -- u (ihmpccontroller.py:108), the weights, ViN_ant, the warm start of the last solve -- stays on the device.
 * One call =  satWeights2 at the measured state with the previous warm start (ihmpccontroller.py:619-636; skipped on
 * the first call)  ->  mpc (:567-604)  ->  u += du, warmStart (:534-564).  Per scenario-step (nx + ny) doubles go
 * up and (nu + 1) doubles + a status word come down, instead of the whole state both ways through sihmpc_step_host.
 * Uploads, the kernels and downloads of consecutive chunks of the batch overlap (three streams).  One resident batch
 * per plan.  `state` is caller-owned DEVICE memory of sihmpc_resident_bytes(); the call synchronises `stream`.    */
typedef struct sihmpc_resident_io {
    const double *x;                 /* HOST [batch][nx]  measured state x0 of this step                 */
    const double *ysp;               /* HOST [batch][ny]  set-point of this step                          */
    double *du;                      /* HOST [batch][nu]  first move (out)                                */
    double *J;                       /* HOST [batch]      cost (out)                                      */
    int32_t *status;                 /* HOST [batch]      SIHMPC_ST_* (out) or NULL                       */
    int32_t flags;                   /* SIHMPC_F_WEIGHTS_SAT2 | SIHMPC_F_WEIGHTS_SAT1 | SIHMPC_F_UPDATE_VIN */
    double alfa;                     /* satWeights2 blending                                               */
} sihmpc_resident_io;
size_t sihmpc_resident_bytes(const sihmpc_plan *plan, int64_t batch);
/* u0 HOST [batch][nu]; pesos0 HOST [m] initial weights of every scenario (init_pesos, ihmpccontroller.py:87-91) */
int sihmpc_resident_init(sihmpc_plan *plan, void *state, size_t state_bytes, int64_t batch,
                         const double *u0, const double *pesos0, void *stream);
int sihmpc_resident_step(sihmpc_plan *plan, void *state, size_t state_bytes, const sihmpc_resident_io *io,
                         int64_t batch, void *stream);
/* byte offset of the weights [batch][m] inside a 256-byte aligned `state` (a loop that is already running hands its
 * per-scenario weights over with one cudaMemcpy there after sihmpc_resident_init)                                  */
int sihmpc_resident_weights_offset(const sihmpc_plan *plan, int64_t batch, size_t *offset);

/* replaces: the `pred` Function of _OptimProbl (ihmpccontroller.py:466-472,523):
 * x_pred [batch][N*nx], y_pred [batch][N*ny], u_pred [batch][N*nu] from (x, u, dU). NULL skips. */
int sihmpc_predict(sihmpc_plan *plan, const double *x, const double *u, const double *dU,
                   int64_t batch, double *x_pred, double *y_pred, double *u_pred, void *stream);

/* replaces: fObj.F(x0,u0,var,ysp) for every weighted sub-objective (ihmpccontroller.py:124-126,
 * 598-599, 628): V [batch][m] evaluated at explicit w = [dU ; syN ; siN].                       */
int sihmpc_subobjectives(sihmpc_plan *plan, const double *x, const double *w, const double *ysp,
                         int64_t batch, double *V, void *stream);      /* n_extra > 0: SIHMPC_E_INVALID */

/* replaces: IHMPCController.mpc of the LOG-BARRIER variant (sihmpc/ihmpccontroller_log.py:552-563 and the
 * NLP of :349-461): minimise Vt + ss^2 - sum_i log(ss*gamma_i - V_i) over (dU, ss) by damped Newton.
 * The plan is compiled from the log controller's declarations (its weighted list VJ defines V_i and
 * gamma_i; `pesos` are not used).  DEVICE pointers: x [batch][nx], u [batch][nu], ysp [batch][ny],
 * w0 [batch][nv+2ny+1] warm start or NULL, w_opt [batch][nv+2ny+1] = [dU ; syN ; siN ; ss], J [batch],
 * status [batch] (SIHMPC_ST_* bits).  When an explicit inequality of the reference NLP -- ss >= 1,
 * ss*gamma_i - V_i >= 0.01, input/move boxes (ihmpccontroller_log.py:390-392,420-424,435-440) -- is not strictly
 * inactive at that minimiser the step is solved again as an interior-point problem (a barrier -mu log c_j per
 * inequality, mu 1e-1 ... 1e-14); bit 8 is only left set when no strictly feasible start exists (u outside its
 * box, 0 outside the move box).
 * iters [batch] Newton iterations.  NULL outputs are skipped.                                      */
#define SIHMPC_ST_ACTIVE_UNSUPPORTED 8   /* log variant: active inequality and no strictly feasible start */
int sihmpc_log_step(sihmpc_plan *plan, const double *x, const double *u, const double *ysp, const double *w0,
                    int64_t batch, double *w_opt, double *J, int32_t *status, int32_t *iters, void *stream);

/* The driver loop of the log-barrier variant (example_distillation_log.py:85-135) for `batch` scenarios and
 * n_steps consecutive control steps in ONE launch: mpc (ihmpccontroller_log.py:552-563), first move to the
 * plant (dynF :103-123), warmStart (:518-549: shifted moves, ss = 1).  `io` as for sihmpc_step (DEVICE
 * pointers; x, u in/out with SIHMPC_F_ADVANCE; ysp or ysp_seq/seg_start; optional dist, plant_val, du, J,
 * status, stats, traj_*; io.w_opt is [batch][nv+2ny+1] here; pesos/vinant are not used).  w_warm
 * [batch][nv+2ny+1] is the warm start on entry and the shifted plan on exit (NULL: cold start, not
 * returned).  iters [batch]: Newton iterations of the last step (NULL skips); io.stats[ny+4] accumulates
 * them.  io.pesos (optional, with SIHMPC_F_ADVANCE): [batch] doubles of solver state carried between launches --
 * ss of the scenario's last optimum, where the Newton iteration of the next step starts (zero-initialised by the
 * caller; the warm start itself keeps ss = 1 as the reference writes it, the optimum is the same, the solve is
 * one iteration shorter).                                                                           */
int sihmpc_log_loop(sihmpc_plan *plan, const sihmpc_io *io, double *w_warm, int64_t batch, int32_t n_steps,
                    int32_t *iters, void *stream);

/* On-device reduction of the per-scenario statistics block to aggregate moments
 * out[3][nstat] = {sum, sum of squares, max} (feeds one NCCL all-reduce). */
int sihmpc_stats_reduce(const double *stats, int64_t batch, int32_t nstat, double *out,
                        void *stream);

/* -- measurement helper --------------------------------------------------------------------- */
/* Dependent-free DFMA loop on every SM: returns measured FP64 FMA throughput in FLOP/s
 * (the roofline denominator for this path; MEASURED_PEAKS.json has no FP64 entry).            */
int sihmpc_fp64_peak(double *flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SIHMPC_B200_H */
