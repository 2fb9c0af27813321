/*
 * nm_rollout.h -- C ABI of the B200-native fused DPC rollout (libnm_rollout.so).
 *
 * The reference (pnnl/neuromancer 1.5.6) is pure Python: it has no FFI for this path.  The
 * boundary the library sits behind is therefore the set of Python call sites listed beside each
 * entry point below; `neuromancer_b200/_lib.py` binds these symbols with ctypes and
 * `INTEGRATION.md` shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C, no C++/torch types.  All tensors are caller-owned, contiguous fp32 DEVICE buffers.
 *   - nothing here allocates device memory or synchronises the host; every launch goes to the
 *     `stream` argument (a cudaStream_t passed as void*).
 *   - every function returns NM_OK (0) or a negative NM_ERR_* code; nm_last_error() gives the text.
 *   - layouts follow the reference: trajectories are [batch, time, dim] row-major
 *     (reference system.py:261-277), nn.Linear weights are [out, in] row-major followed by the
 *     bias [out] (reference blocks.py:131-177, slim/linear.py:104-119).
 */
#ifndef NM_ROLLOUT_H_
#define NM_ROLLOUT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_ABI_VERSION 7

#define NM_MAX_LAYERS 8   /* linear layers per MLP                                   */
#define NM_MAX_SEGS   8   /* concatenated inputs of the policy (blocks.py:39-43)     */
#define NM_MAX_EXO    8   /* exogenous [B,T,d] inputs read at index i (system.py:274) */
#define NM_MAX_TERMS  16  /* objectives + constraints of a PenaltyLoss               */
#define NM_MAX_DIM    16  /* nx, nu, ny                                              */

/* status codes */
#define NM_OK                 0
#define NM_ERR_BAD_PLAN      -1   /* inconsistent / unsupported descriptor                  */
#define NM_ERR_NO_KERNEL     -2   /* no compiled specialisation for this plan's shape class */
#define NM_ERR_CUDA          -3   /* a CUDA runtime call failed                             */
#define NM_ERR_WORKSPACE     -4   /* workspace too small / misaligned                       */
#define NM_ERR_ABI           -5   /* plan.abi_version != NM_ABI_VERSION                     */

/* activations (reference modules/activations.py:230-256; torch.nn semantics) */
#define NM_ACT_IDENTITY 0
#define NM_ACT_RELU     1   /* relu'(0) = 0                                   */
#define NM_ACT_GELU     2   /* exact erf form (nn.GELU() default)             */
#define NM_ACT_LEAKY    3   /* slope act_param (0.01); derivative slope at x<=0 */
#define NM_ACT_TANH     4
#define NM_ACT_ELU      5   /* alpha act_param                                */
#define NM_ACT_SIGMOID  6
#define NM_ACT_SOFTPLUS 7   /* beta=1, threshold=20                           */

/* MLP_bounds output maps (reference blocks.py:555-619, functions.py:9-34) */
#define NM_BOUNDS_NONE          0
#define NM_BOUNDS_SIGMOID_SCALE 1   /* (max-min)*sigmoid(x)+min                */
#define NM_BOUNDS_RELU_CLAMP    2   /* x+relu(-x+min); x-relu(x-max)           */

/* policy input segment sources */
#define NM_SEG_STATE  0   /* x_i  (pre-update state)            */
#define NM_SEG_OUTPUT 1   /* y_i  (same-step output of y = Cx)  */
#define NM_SEG_EXO    2   /* exo[index][:, i, :]                */
#define NM_SEG_PREVIEW 3  /* exo[index][:, i : i+1+preview_len, :] flattened time-major, right-padded
                           * by NmPlan.pad_mode when the window runs past the signal (SystemPreview,
                           * reference system.py:325-367)                                       */

/* torch.nn.functional.pad modes of SystemPreview (system.py:297,341-345) */
#define NM_PAD_CONSTANT  0
#define NM_PAD_REPLICATE 1
#define NM_PAD_CIRCULAR  2
#define NM_PAD_REFLECT   3

/* dynamics */
#define NM_DYN_SSM 0      /* discrete x+ = A x + B u [+ E d]            (ode.py:7-51)        */
#define NM_DYN_ODE 1      /* x+ = integrator(rhs)(x, u)                 (integrators.py)     */

#define NM_INT_EULER 0    /* integrators.py:144-156 */
#define NM_INT_RK4   1    /* integrators.py:196-211 */
#define NM_INT_RK2   2    /* integrators.py:180-193 */
#define NM_INT_EULER_TRAP 3 /* integrators.py:159-177  predictor Euler, corrector trapezoid            */
#define NM_INT_RK4_TRAP   4 /* integrators.py:214-235  predictor RK4, corrector trapezoid              */
#define NM_INT_LUTHER     5 /* integrators.py:238-264  Luther's 6th-order 7-stage method               */
#define NM_INT_RKF        6 /* integrators.py:267-301  Runge-Kutta-Fehlberg, 5th-order output          */
/* second-order schemes  ddx = f(x[, u])  on the state X = [x, dx] (nx even; MLP right-hand side nx/2 [+ nu] -> nx/2)   */
#define NM_INT_LEAPFROG   7 /* integrators.py:340-360  velocity Verlet / leapfrog, 2 right-hand-side evaluations */
#define NM_INT_YOSHIDA4   8 /* integrators.py:363-404  4th-order Yoshida composition, 3 evaluations              */

#define NM_RHS_NONE     0
#define NM_RHS_TWOTANK  1 /* ode.py:216-237; rhs_const = {c1, c2}                            */
#define NM_RHS_VDP      2 /* ode.py:352-368; rhs_const = {mu}                                */
#define NM_RHS_CVDP     3 /* ring of nx/2 coupled VanDerPolControl; rhs_const = {mu, kappa}  */
#define NM_RHS_MLP      4 /* blocks.MLP as the right-hand side (neural ODE)                  */
#define NM_RHS_LINEAR   5 /* dx = A x + B u                                                  */
#define NM_RHS_LORENZ_CTRL 6 /* ode.py:393-414; rhs_const = {rho, sigma, beta}; nu = 2               */
#define NM_RHS_CSTR     7 /* ode.py:417-461; rhs_const = {q/V, Caf, Tf, k0, E/R, mdelH/(rho Cp), UA/V/rho/Cp} */
#define NM_RHS_DUFFING  8 /* ode.py:240-262; rhs_const = {alpha, beta, delta, gamma, omega}; input = time */
#define NM_RHS_BRUSSELATOR 9 /* ode.py:265-281; rhs_const = {alpha, beta}; autonomous                */
#define NM_RHS_LOTKA_VOLTERRA 10 /* ode.py:333-349; rhs_const = {alpha}; autonomous                  */
#define NM_RHS_LORENZ   11 /* ode.py:371-391; rhs_const = {rho, sigma, beta}; autonomous             */
/* hybrid (grey-box) right-hand sides: a white-box equation with one unknown term m = rhs_mlp(x) (blocks.MLP 2 -> 1);  */
/* rhs_mlp.trainable bit 0: the MLP's weights, bit 1: the constants (gradient behind the MLP's parameters)             */
#define NM_RHS_BRUSS_HYBRID 12 /* ode.py:284-304; dx1 = alpha + m - beta x1 - x1, dx2 = beta x1 - m; rhs_const = {alpha, beta} */
#define NM_RHS_LV_HYBRID    13 /* ode.py:307-331; dx1 = alpha x1 - beta m, dx2 = delta m - gamma x2; rhs_const = {alpha, beta, delta, gamma} */
#define NM_RHS_HAS_MLP(k) ((k) == NM_RHS_MLP || (k) == NM_RHS_BRUSS_HYBRID || (k) == NM_RHS_LV_HYBRID)

/* BarrierLoss (reference loss.py:184-257): an entry of a constraint with value v < 0 (satisfied) contributes
 * clamp(barrier(v), 0, ub) (nan / +inf -> 0) instead of nothing, an entry with v >= 0 its penalty as before */
#define NM_BARRIER_NONE     0
#define NM_BARRIER_LOG10    1 /* -log10(-v)                                */
#define NM_BARRIER_LOG      2 /* -log(-v)                                  */
#define NM_BARRIER_INVERSE  3 /* 1 / (-v)                                  */
#define NM_BARRIER_SOFTEXP  4 /* (exp(alpha v) - 1) / alpha + alpha        */
#define NM_BARRIER_SOFTLOG  5 /* -log(1 + alpha (-v - alpha)) / alpha      */
#define NM_BARRIER_EXPSHIFT 6 /* exp(v + shift)                            */
#define NM_TERM_NORM(n)    ((n) & 15)
#define NM_TERM_BARRIER(n) (((n) >> 4) & 15)

/* variables a loss-term atom can refer to */
#define NM_VAR_CONST (-1)
#define NM_VAR_X      0   /* state trajectory   [B, nsteps+1, nx] */
#define NM_VAR_U      1   /* action trajectory  [B, nsteps,   nu] */
#define NM_VAR_Y      2   /* output trajectory  [B, nsteps,   ny] */
#define NM_VAR_EXO0   3   /* exo k is NM_VAR_EXO0 + k             */

#define NM_OP_NONE 0      /* side = a      */
#define NM_OP_ADD  1      /* side = a + b  */
#define NM_OP_SUB  2      /* side = a - b  */
#define NM_OP_MUL  3      /* side = a * b  */

#define NM_CMP_EQ 0       /* constraint.py:137-173 */
#define NM_CMP_LT 1       /* constraint.py:65-98   */
#define NM_CMP_GT 2       /* constraint.py:101-134 */

typedef struct NmMlp {
  int32_t n_layers;                   /* linear layers (hidden + 1)                         */
  int32_t sizes[NM_MAX_LAYERS + 1];   /* insize, hsizes..., outsize                         */
  int32_t act;                        /* NM_ACT_* of the hidden layers (last is Identity)   */
  float   act_param;
  int32_t has_bias;
  int32_t bounds;                     /* NM_BOUNDS_*                                        */
  float   bmin[NM_MAX_DIM];
  float   bmax[NM_MAX_DIM];
  int64_t param_offset;               /* offset (floats) of this MLP inside params_flat     */
  int32_t trainable;                  /* gradients are accumulated for this MLP             */
  int32_t _pad;
} NmMlp;

typedef struct NmSegment {
  int32_t kind;    /* NM_SEG_*                    */
  int32_t index;   /* exo index for NM_SEG_EXO    */
  int32_t width;         /* columns this segment contributes to the policy input         */
  int32_t preview_len;   /* NM_SEG_PREVIEW: L (window = L+1 time entries), else 0        */
} NmSegment;

typedef struct NmExo {
  int32_t width;   /* last dim d_k                 */
  int32_t steps;   /* time length T_k (>= nsteps)  */
} NmExo;

/* atom(b, t, d) = scale * V[b, t_start + (t_len == 1 ? 0 : t), d_start + (d_len == 1 ? 0 : d)] + offset
 * (negative t_start counts from the end, as Python indexing does).  var == NM_VAR_CONST: offset. */
typedef struct NmAtom {
  int32_t var;
  int32_t t_start, t_len;
  int32_t d_start, d_len;
  float   scale, offset;
  int32_t _pad;
} NmAtom;

typedef struct NmSide {
  int32_t op;      /* NM_OP_* */
  int32_t _pad;
  NmAtom  a, b;
} NmSide;

/* one Constraint / Objective of the PenaltyLoss: weight * mean_{b,t,d} penalty(lhs, rhs)
 * (reference constraint.py:307-324, loss.py:61-108) */
typedef struct NmTerm {
  int32_t cmp;           /* NM_CMP_*                                  */
  int32_t norm;          /* bits 0-3: 1 or 2 (Constraint.__xor__); bits 4-7: NM_BARRIER_* of a BarrierLoss constraint (0: none) */
  float   weight;
  int32_t is_objective;  /* listed under objectives (else constraint) */
  int32_t t_len, d_len;  /* broadcast element shape per trajectory    */
  NmSide  lhs, rhs;
} NmTerm;

typedef struct NmPlan {
  int32_t abi_version;
  int32_t nx, nu, ny;
  int32_t nsteps;
  int32_t has_policy;
  int32_t has_output;        /* y_i = C x_i evaluated first in the step        */
  int32_t n_segs;
  NmSegment segs[NM_MAX_SEGS];
  NmMlp   policy;
  int32_t dyn_kind;          /* NM_DYN_*                                       */
  int32_t integrator;        /* NM_INT_*                                       */
  int32_t rhs_kind;          /* NM_RHS_*                                       */
  int32_t dist_exo;          /* exo index of the disturbance d of an SSM, or -1 */
  int32_t act_exo;           /* open loop: exo index the action u_i is read from (no policy), or -1 */
  float   h;                 /* integration step                               */
  float   rhs_const[8];
  NmMlp   rhs_mlp;
  float   A[NM_MAX_DIM * NM_MAX_DIM];   /* [nx][nx] row-major; x+ = x A^T ...   */
  float   Bm[NM_MAX_DIM * NM_MAX_DIM];  /* [nx][nu]                             */
  float   E[NM_MAX_DIM * NM_MAX_DIM];   /* [nx][nd]                             */
  float   C[NM_MAX_DIM * NM_MAX_DIM];   /* [ny][nx]                             */
  int32_t n_exo;
  NmExo   exo[NM_MAX_EXO];
  int32_t n_terms;
  NmTerm  terms[NM_MAX_TERMS];
  int64_t n_params;          /* total floats in params_flat                    */
  int32_t pad_mode;          /* NM_PAD_* for NM_SEG_PREVIEW segments           */
  float   pad_value;         /* NM_PAD_CONSTANT fill value                     */
  float   barrier_alpha;     /* BarrierLoss(alpha=): softexp / softlog bending  (loss.py:196-214) */
  float   barrier_shift;     /* BarrierLoss(shift=): expshift                                      */
  float   barrier_ub;        /* BarrierLoss(upper_bound=): the barrier value is clamped to [0, ub] */
  int32_t reserved0;
} NmPlan;

/* Library / build identification. */
int nm_abi_version(void);
const char* nm_last_error(void);

/* Shape-class signature of a plan: the string the kernel registry is keyed by.  Writes a
 * NUL-terminated string into buf (at most cap bytes); returns its length or NM_ERR_*. */
int nm_plan_signature(const NmPlan* plan, char* buf, int cap);

/* Structure of the plan's fused loss terms (comparators, norms, side ops, variables, broadcast flags).  A
 * specialisation compiled with the same term structure scores the terms inside the rollout / adjoint
 * kernels; any other specialisation of the shape class serves the plan through the runtime-descriptor
 * term kernels.  Same calling convention as nm_plan_signature. */
int nm_plan_term_signature(const NmPlan* plan, char* buf, int cap);

/* 1 if a compiled sm_100a specialisation is registered for the plan's signature, else 0. */
int nm_has_kernel(const NmPlan* plan);

/* Number of registered specialisations and their signatures (for diagnostics/tests). */
int nm_num_kernels(void);
const char* nm_kernel_signature(int i);
/* "tile=.. threads=.. smem_fwd=.. smem_bwd=.." of the specialisation serving `plan`, or NULL. */
const char* nm_kernel_info(const NmPlan* plan);
/* sizeof() of NmMlp, NmSegment, NmExo, NmAtom, NmSide, NmTerm, NmPlan as compiled (binding self-check);
 * writes min(cap, 7) entries and returns 7. */
int nm_struct_sizes(int* out, int cap);

/* Load a shared object holding additional specialisations (built by neuromancer_b200.build
 * for a plan whose shape class was not compiled ahead of time).  Returns NM_OK or NM_ERR_*. */
int nm_load_plugin(const char* path);

/* Size in bytes of the optional checkpoint buffer for a batch of B trajectories (0 when the plan has none).
 * The emitted x trajectory is always the per-step checkpoint of the adjoint.  Additionally:
 *   - plans with an MLP right-hand side keep the stage derivatives k_0..k_{NS-2} of every integrator step
 *     ([B, nsteps, (NS-1)*nx] fp32) so that the adjoint skips its first recompute pass;
 *   - closed-loop plans whose policy runs on the tensor-core kernels keep the raw policy output o_t (before
 *     MLP_bounds) of every step ([B, nsteps, nu] fp32).
 * The buffer is filled by nm_rollout_forward and read by nm_rollout_backward; NULL is always accepted. */
int nm_rollout_checkpoint_bytes(const NmPlan* plan, int64_t B, size_t* bytes);

/* Device workspace sizes in bytes for a batch of B trajectories. */
int nm_rollout_workspace_bytes(const NmPlan* plan, int64_t B, size_t* fwd_bytes, size_t* bwd_bytes);

/*
 * Forward rollout + PenaltyLoss scoring.
 * Replaces: System.forward (system.py:261-277) incl. Node.forward (:48-59), System.cat (:234-246),
 *           Block.forward/MLP.block_eval (blocks.py:39-43,169-177), MLP_bounds (blocks.py:611-619),
 *           RK4/Euler.integrate (integrators.py:144-211), ODESystem/SSM.forward (ode.py:33-72),
 *           and the evaluation of Constraint.forward / PenaltyLoss.forward (constraint.py:307-324,
 *           loss.py:61-108,168-181) on the produced trajectories.
 *
 *   params    [plan->n_params]            flat policy / RHS-MLP parameters
 *   x0        [B, nx]                     initial state (data[x][:, 0])
 *   exo[k]    [B, exo[k].steps, exo[k].width]
 *   x_traj    [B, nsteps+1, nx]  out      (row 0 is written with x0)
 *   u_traj    [B, nsteps, nu]    out      (NULL when nu == 0)
 *   y_traj    [B, nsteps, ny]    out      (NULL when ny == 0)
 *   scalars   [n_terms + 3]      out fp32: weight*mean of each term (list order: objectives then
 *                                constraints as given in plan->terms), objective_loss, penalty_loss, loss
 *   workspace fwd_bytes, 256-byte aligned
 */
int nm_rollout_forward(const NmPlan* plan, const float* params, const float* x0,
                       const float* const* exo, float* x_traj, float* u_traj, float* y_traj,
                       float* scalars, float* checkpoints /* nm_rollout_checkpoint_bytes, or NULL */,
                       void* workspace, size_t workspace_bytes, int64_t B, void* stream);

/*
 * Adjoint / backward-through-time pass.
 * Replaces: output[train_metric].backward() (trainer.py:250) for the graph above.
 *
 *   grad_scalars [n_terms + 3]  upstream gradient of `scalars` (NULL: d loss = 1)
 *   grad_x_ext   [B, nsteps+1, nx] optional extra upstream gradient w.r.t. x_traj (terms evaluated
 *                by the caller outside the kernel), or NULL; likewise grad_u_ext, grad_y_ext
 *   grad_params  [plan->n_params] out (overwritten, not accumulated)
 *   grad_x0      [B, nx] out, or NULL
 *   b_total      global batch size the loss means are taken over (== B on one GPU; B * world_size
 *                when each rank holds an equal shard and gradients are summed by all-reduce)
 */
int nm_rollout_backward(const NmPlan* plan, const float* params, const float* const* exo,
                        const float* x_traj, const float* u_traj, const float* y_traj,
                        const float* checkpoints /* written by the forward call, or NULL: recompute */,
                        const float* grad_scalars, const float* grad_x_ext, const float* grad_u_ext,
                        const float* grad_y_ext, float* grad_params, float* grad_x0,
                        void* workspace, size_t workspace_bytes, int64_t B, int64_t b_total,
                        void* stream);

/* Number of kernel launches the last forward / backward call issued (for bench.py gpu_launches). */
int nm_last_launch_count(void);

/* Fused gradient clipping + AdamW update over the flat parameter buffer the rollout kernels read and the flat gradient
 * buffer nm_rollout_backward wrote (after the data-parallel all-reduce, if any).  Replaces, for the parameters of a fused
 * problem, the tail of the reference's training step -- trainer.py:251-252:
 *     torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip);  self.optimizer.step()
 * with torch.optim.AdamW arithmetic (decoupled weight decay, bias-corrected moments):
 *     coef = min(1, max_norm / (||g||_2 + 1e-6));  g <- coef * g
 *     p <- p (1 - lr wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
 *     p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 * Two launches (sum of squares in a fixed order: deterministic; update), no host synchronisation.
 *   params, exp_avg, exp_avg_sq  [n] device fp32, updated in place;  grads [n] device fp32 (read only)
 *   step        1-based update count t
 *   max_norm    <= 0: no clipping
 *   mask        [n] device bytes or NULL: entries with mask[i] == 0 are left untouched (frozen parameters); they do not
 *               enter the norm either
 *   norm_out    [1] device fp32 or NULL: the gradient norm before clipping
 *   workspace   device memory, >= nm_adamw_workspace_bytes(n) bytes, 256-byte aligned
 */
size_t nm_adamw_workspace_bytes(int64_t n);
int nm_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const unsigned char* mask, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float max_norm,
                  float* norm_out, void* workspace, size_t workspace_bytes, void* stream);

/* Constraint matrices of AggregateLoss.calculate_constraints (reference loss.py:86-106) for the plan's fused constraint
 * terms, in list order: row b = the constraint values (C_values) / violations (C_violations) of trajectory b, each
 * constraint contributing t_len * d_len columns (its [t][d] entries, row-major), and the same columns split into the
 * equality (Eq comparator) and inequality partitions.  One pass over the trajectories, every output written once.
 *   widths      [3] out (may be called with B = 0 and NULL buffers to query them): total, equality, inequality columns
 *   c_values / c_violations            [B, widths[0]] device fp32 or NULL
 *   c_eq_values / c_eq_violations      [B, widths[1]] device fp32 or NULL
 *   c_ineq_values / c_ineq_violations  [B, widths[2]] device fp32 or NULL
 * The violation of a BarrierLoss constraint is its plain penalty, as in the reference.
 */
int nm_constraint_matrices(const NmPlan* plan, const float* const* exo, const float* x_traj, const float* u_traj, const float* y_traj,
                           float* c_values, float* c_violations, float* c_eq_values, float* c_eq_violations,
                           float* c_ineq_values, float* c_ineq_violations, int64_t widths[3], int64_t B, void* stream);

/* Data-parallel exchange of the path: in-kernel SUM over ranks of the flat [gradient | loss scalars]
 * buffer through NVLink / NVSwitch peer memory -- replaces the NCCL all-reduce DDP would issue (reference trainer.py:49).
 * The caller owns a SYMMETRIC buffer of nm_grad_allreduce_buffer_floats(n) floats per rank (zero-initialised, mapped on
 * every peer; torch.distributed._symmetric_memory or cuMem* + IPC), lets nm_rollout_forward / nm_rollout_backward write
 * scalars and gradients of step `epoch` into half (epoch & 1) of its own buffer (offset (epoch & 1) * npad floats,
 * npad = n rounded up to 64), and calls this on every rank with the same epoch = 1, 2, 3, ...
 *   peer_bufs      [world] host array: every rank's symmetric buffer as mapped in THIS process (peer_bufs[rank] = own)
 *   multicast_ptr  the buffer's multicast address (NVLS: the sum is formed inside the switch by multimem.ld_reduce) or
 *                  NULL (peer loads summed in rank order)
 *   out            [n] device fp32, not part of the symmetric buffer; entries >= tail_from are multiplied by tail_scale
 *                  (loss scalars are averaged, gradients summed)
 * No host synchronisation; a peer that does not arrive within ~30 s sets status word
 * ((uint32_t*)(own + 2 * npad))[17] instead of hanging the device.
 */
size_t nm_grad_allreduce_buffer_floats(int64_t n);
int nm_grad_allreduce(void* const* peer_bufs, const void* multicast_ptr, float* out, int64_t n, int rank, int world,
                      uint32_t epoch, int64_t tail_from, float tail_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NM_ROLLOUT_H_ */
