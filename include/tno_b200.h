/*
 * tno_b200.h -- C ABI of the B200-native batched thrust-network evaluator.
 *
 * The reference (compas_tno 0.3.0, pure Python) has no FFI; its seam for this path is the
 * Python callable protocol stored on `Problem` (src/compas_tno/problems/problems.py:214-218,
 * written by src/compas_tno/problems/setup.py:466-469, consumed by
 * src/compas_tno/solvers/solver.py:126-129):
 *
 *     fconstr(x, problem) -> g   (ng,)        constr_wrapper         problems/constraints.py:20
 *     fjac   (x, problem) -> J   (ng, nvar)   sensitivities_wrapper  problems/jacobian.py:53
 *     fobj   (x, problem) -> f   scalar       f_min_thrust / ...     problems/objectives.py:91,151,415,438
 *     fgrad  (x, problem) -> df  (nvar,)      gradient_fmin / ...    problems/derivatives.py:223,336,183,203
 *
 * all built on  q_from_variables / xyz_from_q  (algorithms/equilibrium.py:163-192, :195-236).
 * The entry points below are what a binding for that protocol needs: one plan per form diagram
 * (= everything `Problem` carries that is shared by all candidates) and one batched evaluation
 * returning f, grad f, g, J, z and a status word for N variable vectors at once.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary.
 *   - all floating point data is IEEE fp64; index data is int32; row-major ("C") layouts.
 *   - every function returns 0 on success and a negative TNO_E* code otherwise; the message is
 *     available from tno_last_error() (thread local).  Per-candidate numerical failures are
 *     reported in `status[]`, never through the return code.
 *   - a plan is immutable after creation.  tno_eval_batch is asynchronous on the given CUDA
 *     stream and uses plan-owned scratch memory: concurrent calls on ONE plan must be ordered
 *     by the caller (use one plan per stream for concurrency).
 *   - there is no CPU fallback: without a CUDA device tno_plan_create fails with TNO_ENODEVICE.
 */
#ifndef TNO_B200_H
#define TNO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNO_ABI_VERSION 2

/* error codes */
#define TNO_OK 0
#define TNO_EINVAL (-1)    /* bad argument (null pointer, sizes, alignment, malformed arrays) */
#define TNO_ENODEVICE (-2) /* no usable CUDA device */
#define TNO_ECUDA (-3)     /* CUDA runtime error */
#define TNO_ENOMEM (-4)
#define TNO_EUNSUPPORTED (-5) /* valid arguments, but a combination this library (like the reference: NotImplementedError)
                                 does not provide */

/* variables present in x, in the reference's unpacking order (problems/constraints.py:47-88):
 *   x = [ q_ind (k) | xyb (2 nb: x of the supports, then y) | zb (nb) | t (1) | lambdh (1) | lambdv (1) ]   */
#define TNO_VAR_Q 1u
#define TNO_VAR_ZB 2u
#define TNO_VAR_T 4u      /* 't' or 'n': thickness handed to the parametric envelope */
#define TNO_VAR_LAMBDH 8u
#define TNO_VAR_LAMBDV 16u
#define TNO_VAR_XYB 32u   /* horizontal support positions (needs the envelopexy and envelope constraints, like the
                             reference's Jacobian assembly problems/jacobian.py:269-273; interleaved kernel family) */

/* constraint blocks, stacked in this order (problems/constraints.py:104-155) */
#define TNO_CON_FUNICULAR 1u   /* q - qmin ; qmax - q                      2m rows */
#define TNO_CON_ENVELOPE 2u    /* z - lb ; ub - z                          2n rows */
#define TNO_CON_REAC_BOUNDS 4u /* |b| - (zb - s) |R_xy / R_z|              2nb rows */
#define TNO_CON_ENVELOPEXY 8u  /* x - xlim0 ; xlim1 - x ; y - ylim0 ; ylim1 - y   4n rows, between funicular and envelope */

/* objectives (problems/objectives.py:27-88) */
#define TNO_OBJ_NONE 0
#define TNO_OBJ_MIN 1        /* f_min_thrust  / gradient_fmin */
#define TNO_OBJ_MAX 2        /* f_max_thrust  / gradient_fmax */
#define TNO_OBJ_T 3          /* f_reduce_thk  / gradient_reduce_thk          f = +x[-1] */
#define TNO_OBJ_NEG_LAST 4   /* f_tight_crosssection ('n','s','max_load')    f = -x[-1] */
#define TNO_OBJ_BESTFIT 5    /* f_bestfit     / gradient_bestfit   ('bestfit','target'; variables q [, zb]; needs s)
                                problems/objectives.py:170-200, problems/derivatives.py:355-413 */
#define TNO_OBJ_FEASIBILITY 6 /* f_constant   / gradient_feasibility   f = 1, grad = 0  (objectives.py:395-412) */
#define TNO_OBJ_LOADPATH 7   /* f_loadpath_general / gradient_loadpath  f = sum |q| l^2  (variables q [, zb])
                                problems/objectives.py:240-275, problems/derivatives.py:640-718 */
#define TNO_OBJ_ECOMP_LINEAR 8    /* f_complementary_energy / gradient_complementary_energy   ('Ecomp-linear'; needs dXb)
                                     problems/objectives.py:278-312, problems/derivatives.py:503-582 */
#define TNO_OBJ_ECOMP_NONLINEAR 9 /* f_complementary_energy_nonlinear / its gradient, method 'simplified'
                                     ('Ecomp-nonlinear'; needs dXb and stiff)  objectives.py:315-340, derivatives.py:585-620 */

/* parametric envelopes (arithmetic lives in the un-vendored compas_tna; closed forms recovered from
 * the reference fixtures) */
#define TNO_ENV_FIXED 0      /* ub / lb constant (thickness is not a variable) */
#define TNO_ENV_DOME 1       /* env_params = { xc, yc, R, lb_outside } */
#define TNO_ENV_CROSSVAULT 2 /* env_params = { xc, yc, rho, lb_outside } */

/* kernel families (tno_problem_desc.kernel) */
#define TNO_KERNEL_AUTO 0
#define TNO_KERNEL_INTERLEAVED 1 /* batch-interleaved scratch in HBM, level-scheduled factor and solves; any size */
#define TNO_KERNEL_ONCHIP 2      /* factor + right-hand sides resident in shared memory; size limited */
#define TNO_KERNEL_INTERLEAVED_TPC 3 /* interleaved family with its original thread-per-candidate factor / solve kernels
                                        (kept as an independent cross-check; slow on large problems) */

/* per-candidate status bits */
#define TNO_STATUS_OK 0
#define TNO_STATUS_NOT_PD 1    /* a pivot of -(Ci^T Q Ci) was <= 0: the reference's LU would still run */
#define TNO_STATUS_NONFINITE 2 /* NaN / Inf in an output (the reference propagates these silently) */

typedef struct tno_plan tno_plan;

/* Everything of `Problem` (reference problems/problems.py:39-224) the path reads.  Host pointers,
 * copied during tno_plan_create.  Optional arrays may be NULL as noted. */
typedef struct tno_problem_desc {
  int32_t abi_version; /* = TNO_ABI_VERSION */
  int32_t n;           /* vertices                                      Problem.n  */
  int32_t m;           /* edges                                         Problem.m  */
  int32_t nb;          /* supports                                      Problem.nb */
  int32_t k;           /* independent edges                             Problem.k  */
  const int32_t* edges; /* m x 2 (u, v): C[e,u] = -1, C[e,v] = +1       Problem.C  */
  const int32_t* fixed; /* nb, in the order of form.supports()          Problem.fixed */
  const double* B;      /* m x k                                        Problem.B  */
  const double* d;      /* m   particular solution (d0 when lambdh is a variable)  Problem.d / d0 */
  const double* P;      /* n x 3 loads                                  Problem.P  */
  const double* X;      /* n x 3 coordinates; z of supports is used when zb is not a variable  Problem.X */
  const double* lb;     /* n                                            Problem.lb */
  const double* ub;     /* n                                            Problem.ub */
  const double* s;      /* n   middle surface (reac_bounds only; NULL = 0)          Problem.s  */
  const double* qmin;   /* m                                            Problem.qmin */
  const double* qmax;   /* m                                            Problem.qmax */
  const double* b;      /* nb x 2 reaction bounds (reac_bounds only)    Problem.b  */
  const double* px0;    /* n  (lambdh only)                             Problem.px0 */
  const double* py0;    /* n  (lambdh only)                             Problem.py0 */
  const double* pzv;    /* n  (lambdv only)                             Problem.pzv */
  const double* pz0;    /* n  (lambdv only)                             Problem.pz0 */
  uint32_t variables;   /* TNO_VAR_* mask; TNO_VAR_Q is mandatory */
  uint32_t constraints; /* TNO_CON_* mask */
  int32_t objective;    /* TNO_OBJ_* */
  int32_t envelope_kind; /* TNO_ENV_*; must be parametric when TNO_VAR_T is set */
  double env_params[8];
  double thk;           /* Problem.thk (informational unless the envelope needs it) */
  int32_t device;       /* CUDA device ordinal */
  int32_t kernel;       /* TNO_KERNEL_* */
  int32_t ordering;     /* 0 = auto (minimum degree), 1 = natural, 2 = minimum degree, 3 = nested dissection */
  int32_t reserved;
  const double* dXb;    /* nb x 3 prescribed support displacements (Ecomp objectives only)   Problem.dXb   */
  const double* stiff;  /* m   0.5 l^2 / k per edge (Ecomp-nonlinear only)                   Problem.stiff */
  /* second horizontal load basis for per-candidate load directions (tno_batch_inputs.load_dir); all three or none */
  const double* d_b;    /* m   particular solution of the second basis (d is that of the first) */
  const double* px0_b;  /* n */
  const double* py0_b;  /* n */
  const double* xlimits; /* n x 2 (envelopexy only)                      Problem.xlimits */
  const double* ylimits; /* n x 2 (envelopexy only)                      Problem.ylimits */
} tno_problem_desc;

/* What the symbolic analysis found; all sizes a caller needs to allocate outputs. */
typedef struct tno_plan_info {
  int32_t n, m, ni, nb, k;
  int32_t nvar;        /* length of x */
  int32_t ng;          /* length of g */
  int32_t nrhs;        /* right-hand sides solved per candidate (x, y, z, zb block, q_ind block, lambda) */
  int32_t nnz_a;       /* lower-triangle non-zeros of Ci^T Q Ci */
  int32_t nnz_l;       /* non-zeros of its Cholesky factor */
  int32_t etree_height;
  int32_t max_colcount;
  int32_t kernel;      /* TNO_KERNEL_* actually selected */
  int32_t chunk;       /* candidates processed per launch sequence */
  int64_t scratch_bytes;   /* plan-owned device scratch */
  int64_t factor_flops;    /* flops of one numeric factorisation */
  int64_t solve_flops;     /* flops of all triangular solves of one evaluation */
  int64_t smem_bytes;      /* dynamic shared memory per CTA of the main kernel */
  int32_t last_kernel;     /* TNO_KERNEL_* the most recent tno_eval_batch ran (0 before the first call).  A plan created
                              with TNO_KERNEL_AUTO whose on-chip layout leaves room for one CTA per SM only hands batches
                              of >= 6144 candidates to the interleaved family, which is faster there */
  int32_t onchip_ctas_per_sm;
} tno_plan_info;

/* Optional per-candidate inputs that are not part of x (device pointers, NULL = not used). */
typedef struct tno_batch_inputs {
  const double* pz_scale; /* N   vertical load multiplier: P[:,2] *= pz_scale[j] (Monte-Carlo self-weight) */
  const double* load_dir; /* N x 2 (c, s): horizontal load direction of candidate j (needs the lambdh variable and the
                             second load basis d_b / px0_b / py0_b of the problem):
                             d0 = c d + s d_b,  px0 = c px0 + s px0_b,  py0 = c py0 + s py0_b */
  const double* bounds;   /* N x 2n: per candidate [ub (n) | lb (n)] replacing Problem.ub / Problem.lb in the envelope rows
                             of g (Monte-Carlo geometry perturbations of the intrados / extrados; not with the t variable) */
} tno_batch_inputs;

/* Output buffers (device pointers for tno_eval_batch, host pointers for tno_eval_batch_host).
 * Any pointer may be NULL: that output is then not produced. */
typedef struct tno_outputs {
  double* f;       /* N                       objective */
  double* grad;    /* N x nvar                gradient of the objective */
  double* g;       /* N x ng                  constraints, g >= 0 feasible */
  double* J;       /* N x ng x nvar           dense Jacobian of g, row-major per candidate; 16-byte aligned base */
  double* z;       /* N x n                   node heights (all vertices, supports included) */
  double* q;       /* N x m                   force densities */
  double* gmin;    /* N                       min_r g[r]  (feasibility margin) */
  int32_t* status; /* N                       TNO_STATUS_* bits */
} tno_outputs;

int tno_abi_version(void);
const char* tno_last_error(void);

int tno_plan_create(const tno_problem_desc* desc, tno_plan** out_plan);
void tno_plan_destroy(tno_plan* plan);
int tno_plan_query(const tno_plan* plan, tno_plan_info* info);

/* Copy the symbolic analysis out (host arrays sized from tno_plan_info): perm[ni] maps factor row ->
 * free-vertex position, colptr[ni+1] / rowidx[nnz_l] the pattern of L, parent[ni] the elimination
 * tree, level[ni] the level-set index of each column.  Any pointer may be NULL. */
int tno_plan_symbolic(const tno_plan* plan, int32_t* perm, int32_t* colptr, int32_t* rowidx,
                      int32_t* parent, int32_t* level);

/* Host-only entry to the symbolic analysis (no CUDA device needed; used by the CPU test-suite).  The matrix pattern is
 * given as an edge list between n_free rows.  ordering as in tno_problem_desc; chain_width = widest level still
 * handled as dense chains by the on-chip kernels (0 = none).  Outputs may be NULL; info[8] receives
 * { nnz_l, etree height, max column count, wide levels, wide rows, chains, chain levels, factor flops }. */
int tno_symbolic_analyse(int32_t n_free, int32_t n_edges, const int32_t* edges, int32_t ordering, int32_t chain_width,
                         int32_t* perm, int32_t* colptr, int32_t* rowidx, int32_t rowidx_capacity, int32_t* parent,
                         int32_t* chain_ptr, int32_t chain_capacity, int32_t* info);

/* Evaluate N candidates.  x_dev: device pointer, candidate j at x_dev + j * ldx (ldx >= nvar).
 * `cuda_stream` is a cudaStream_t (NULL = default stream).  Asynchronous. */
int tno_eval_batch(tno_plan* plan, const double* x_dev, int64_t n_candidates, int64_t ldx,
                   const tno_batch_inputs* opt, const tno_outputs* out_dev, void* cuda_stream);

/* Same with HOST buffers: stages x through pinned memory, runs tno_eval_batch on the plan's own
 * stream, copies the requested outputs back and synchronises.  This is the call behind the
 * batch-1 drop-in callbacks and the end-to-end figure of bench.py. */
int tno_eval_batch_host(tno_plan* plan, const double* x_host, int64_t n_candidates, int64_t ldx,
                        const tno_batch_inputs* inputs_host, const tno_outputs* out_host);

/* Number of kernel launches issued by this plan so far (bench.py reports it as gpu_launches). */
int64_t tno_plan_launch_count(const tno_plan* plan);

/* Per-kernel device timing with CUDA events recorded on the launching stream (used by bench.py for the
 * roofline of the dominant kernel; off by default).  enable != 0 clears the counters and starts recording.
 * tno_plan_profile_read synchronises the recorded events and returns the number of kernel kinds; for kind i
 * names + 32 * i is its NUL-terminated name, total_ms[i] the summed duration, launches[i] the launch count. */
int tno_plan_profile(tno_plan* plan, int enable);
int tno_plan_profile_read(tno_plan* plan, int32_t max_kinds, char* names, double* total_ms, int64_t* launches);

/* Bring the dense Jacobians of n_candidates evaluations from device memory (J_dev, layout of tno_outputs.J) to host
 * memory (J_host, same layout; pinned memory makes the copy asynchronous).  The candidate-independent funicular rows
 * [B 0; -B 0] (+ d0 column) do not cross PCIe: host threads write them from the plan's copy while the DMA engine moves
 * the other rows on `cuda_stream`.  Returns when the host threads are done; the caller synchronises the stream.
 * With per-candidate load directions (tno_batch_inputs.load_dir) the funicular rows are not constant: use a plain copy. */
int tno_fetch_jacobian_host(tno_plan* plan, const double* J_dev, double* J_host, int64_t n_candidates, void* cuda_stream);

/* Device peaks measured in place with register-resident kernels (a few milliseconds each), the denominators of the
 * secondary FP64 roofline bench.py reports (north_star: "fraction of the HBM or FP64 roofline", "FP64 tensor-pipe
 * utilisation").  what = TNO_PEAK_FP64_FMA: DFMA throughput;  TNO_PEAK_FP64_MMA: mma.sync m8n8k4 f64 (FP64 tensor core)
 * throughput.  *value receives TFLOP/s. */
#define TNO_PEAK_FP64_FMA 0
#define TNO_PEAK_FP64_MMA 1
int tno_device_peak(int32_t device, int32_t what, double* value);

/* Statistics of the most recent tno_fetch_jacobian_host of a plan: stats[4] = { fraction of the constant (funicular) rows
 * written by the host threads, measured host fill rate [GB/s], measured DMA rate [GB/s], host threads used }.  The split
 * between the host threads and the DMA engine follows the two measured rates (both finish together), starting from a
 * model (52 GB/s PCIe, 4.5 GB/s per thread); TNO_HOST_FILL_FRACTION fixes it. */
int tno_plan_fetch_stats(const tno_plan* plan, double* stats);

/* ---- batched multi-start driver -------------------------------------------------------------------
 * Building blocks of S SQP iterations that advance together with iterates, f, grad f, g and the dense J resident in device
 * memory.  They replace scipy.optimize.fmin_slsqp as the reference calls it (src/compas_tno/solvers/scipy.py:127-174:
 * objective / gradient / f_ieqcons / fprime_ieqcons callables + per-variable bounds); the host loop with the line search is
 * compas_tno_b200/multistart.py.  All pointers are DEVICE pointers, row-major, one CTA per start, asynchronous on
 * `cuda_stream`.  1 <= nvar <= 64.
 *
 * tno_sqp_qp_solve: for every start s solve
 *       min_d 1/2 d'H_s d + grad_s'd    s.t.   J_s d + g_s >= 0 ,   lb <= x_s + d <= ub
 *   with a primal-dual interior point method (Mehrotra predictor-corrector) on the dense J.  Rows [0, nshared) of every J_s
 *   are read from J_shared instead (the candidate-independent funicular rows [B 0; -B 0]: one L2-resident copy).
 *     H (S x nvar x nvar), grad (S x nvar), g (S x ng), J (S x ng x nvar), x (S x nvar), lb / ub (nvar; +-1e20 = none),
 *     active (S bytes or NULL; 0 = skip the start, d = 0)
 *   out: d (S x nvar); lam (S x (ng + 2 nvar)) multipliers of [constraints | lower bounds | upper bounds];
 *     info (S x 8) = { grad'd, sum lam |b|, sum max(0, -g), max lam, iterations, converged (0/1), mu, max |primal residual| };
 *     lagr_grad (S x nvar) = grad - J' lam_g, the Lagrangian gradient at x with the NEW multipliers (for the BFGS update);
 *     work (S x 4 x (ng + 2 nvar)) scratch.
 *   use_mma != 0 forms the normal matrix H + J' diag(lam / s) J on the FP64 tensor pipe (mma.sync m8n8k4.f64).
 *
 * tno_sqp_bfgs_update: H_s <- damped BFGS update with step_s = x_new - x_old and
 *       y = (grad_new - J_new' lam_g) - lagr_grad_old      (Powell damping keeps H positive definite). */
int tno_sqp_qp_solve(int64_t n_starts, int32_t nvar, int32_t ng, int32_t nshared, const double* H, const double* grad,
                     const double* g, const double* J, const double* J_shared, const double* x, const double* lb,
                     const double* ub, const unsigned char* active, double* d, double* lam, double* info, double* lagr_grad,
                     double* work, int32_t max_iter, double tol, int32_t use_mma, void* cuda_stream);
int tno_sqp_bfgs_update(int64_t n_starts, int32_t nvar, int32_t ng, int32_t nshared, double* H, const double* grad_new,
                        const double* J_new, const double* J_shared, const double* lam, const double* lagr_grad_old,
                        const double* step, const unsigned char* active, void* cuda_stream);

/* ---- one-time pre-processing, batched over plan geometries ------------------------------------------
 * Replaces, for G geometries of ONE form-diagram topology at once, what FixedProblem.from_formdiagram does on the host
 * (src/compas_tno/problems/problems.py:516-528):
 *       Edinv = -pinv(E[:, dep]) ;  B[dep] = Edinv E[:, ind] ;  B[ind] = I ;  d[dep] = -Edinv ph
 * with E = [Cit diag(C x) ; Cit diag(C y)] (problems.py:314-326) assembled on the device from the vertex coordinates.  The set
 * of independent edges `ind` is an INPUT (the reference picks it once per diagram with a pivoted QR,
 * algorithms/independents.py:126-154, and the choice is LAPACK tie-break dependent): E[:, dep] must have full column rank,
 * then pinv(Ed) Y is the least-squares solution of Ed X = Y, computed here by a blocked Householder QR (compact WY,
 * 16-column panels; the trailing updates are dense contractions and run on the FP64 tensor pipe unless use_mma = 0).
 *
 *   edges       m x 2 (u, v): C[e, u] = -1, C[e, v] = +1                                   HOST
 *   free_index  n: position of vertex v among the free vertices (ascending), or -1 for a support    HOST
 *   ind         k independent edges                                                          HOST
 *   xy          G x n x 2 plan coordinates                                                   DEVICE
 *   ph          horizontal loads [px_free ; py_free] (2 ni), geometry g at ph + g * ph_stride
 *               (ph_stride = 0: shared by all geometries); null = zero loads                 DEVICE
 *   B           G x m x k row-major;  d  G x m                                               DEVICE, out
 *   info        G x 4: { largest least-squares residual entry (0 when `ind` fits the geometry), min |r_jj|, max |r_jj|,
 *               largest |entry| of the solution }                                            DEVICE, out
 *   work        tno_preprocess_work_bytes(n, m, ni, k, G) bytes of device memory
 * Asynchronous on `cuda_stream` after the index tables have been uploaded.  G <= 65535 per call. */
int64_t tno_preprocess_work_bytes(int32_t n, int32_t m, int32_t ni, int32_t k, int64_t n_geometries);
int tno_preprocess_batch(int32_t n, int32_t m, int32_t ni, int32_t k, const int32_t* edges, const int32_t* free_index,
                         const int32_t* ind, const double* xy, const double* ph, int64_t ph_stride, int64_t n_geometries,
                         double* B, double* d, double* info, void* work, int64_t work_bytes, int32_t use_mma,
                         void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TNO_B200_H */
