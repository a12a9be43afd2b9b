/* mpcb200.h -- C ABI of libmpcb200.so: batched nonlinear-MPC solves on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of TGoldC/Motion-Planning-for-Autonomous-Driving-with-MPC:
 * the per-MPC-step NLP solve that `CasadiOptimizer.optimize()` performs with
 *     sol, f = self.solver();  res = sol(x0=init_control, p=c_p, lbg=..., lbx=..., ubg=..., ubx=...)
 * (/root/reference/MPC_Planner/optimizer.py:605-607), plus the two tiny host steps either side of it
 * (`shift_movement` :645-655, `desired_command_and_trajectory` :657-702) so that a closed loop can stay on the device.
 *
 * Plain pointers and sizes only; every `d_` pointer is a DEVICE pointer (e.g. torch.Tensor.data_ptr()), every
 * `h_` pointer is a HOST pointer.  All calls are asynchronous on the given CUDA stream unless stated.  Return value:
 * 0 on success, negative on an API / CUDA error (text via mpcb200_last_error).  Per-problem solver outcome goes to
 * `d_status[B]` using the Forcespro exit codes the reference already knows (test/FORCESNLPsolver/include/
 * FORCESNLPsolver.h:70-106): 1 optimal, 0 iteration limit, -6 NaN, -7 no progress; additionally 3 = feasible but stalled
 * at the rounding-noise floor of the arithmetic (usable, accuracy below tol_step) and -8 = the pinned
 * stage X_0 violates a constraint (IPOPT would report an infeasible problem).  The batch is never aborted
 * (contrast optimizer.py:330).
 *
 * Array layouts (row-major, float64 like the reference's numpy arrays; arithmetic precision is cfg.precision):
 *     xref  [B][N+1][5]   the reference's X_ref parameter block: row 0 = current state (pinned X_0), rows 1..N =
 *                         [path_x, path_y, 0, v_des, orientation]            (optimizer.py:600, 667-699)
 *     X     [B][N+1][5]   in: state warm start / out: optimal states          (optimizer.py:602, 617)
 *     U     [B][N][2]     in: control warm start / out: optimal controls      (optimizer.py:602, 616)
 *   state = [sx, sy, delta, v, psi], control = [delta_dot, a]                 (configuration.py:354-368)
 */
#ifndef MPCB200_H
#define MPCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB200_ABI_VERSION 4

enum { MPCB200_F32 = 0, MPCB200_F64 = 1 };
enum { MPCB200_HESS_GAUSS_NEWTON = 0, MPCB200_HESS_EXACT = 1 };

/* Everything `Optimizer.__init__` pulls out of `configuration` (optimizer.py:34-68) + solver options.
 * Replaces: the Python attributes of Optimizer, the lbg/ubg/lbx/ubx lists of inequal_constraints (optimizer.py:413-491)
 * and the IPOPT option dict (optimizer.py:556). */
typedef struct mpcb200_config {
  int32_t abi_version;      /* = MPCB200_ABI_VERSION */
  int32_t device;           /* CUDA device ordinal */
  int32_t N;                /* predict_horizon (optimizer.py:56) */
  int32_t max_batch;        /* largest B a call will pass */
  int32_t precision;        /* MPCB200_F32 | MPCB200_F64 : arithmetic type of the kernels */
  int32_t hessian;          /* MPCB200_HESS_GAUSS_NEWTON | MPCB200_HESS_EXACT */
  int32_t max_iter;         /* SQP iteration limit per solve (ipopt.max_iter = 100, optimizer.py:556) */
  int32_t ls_max;           /* line-search trial limit */
  double dt;                /* configuration.delta_t */
  double l_wb;              /* p.a + p.b = 2.5789128 (configuration.py:362-363) */
  double l_fric;            /* 2.578, literal of the friction row (optimizer.py:378) */
  double Q[5];              /* weight_x, weight_y, weight_steering_angle, weight_velocity, weight_heading_angle */
  double R[2];              /* weight_velocity_steering_angle, weight_long_acceleration (optimizer.py:500-504) */
  double deltav_min, deltav_max;   /* optimizer.py:40-41 */
  double a_max;                    /* optimizer.py:46 */
  double delta_min, delta_max;     /* optimizer.py:37-38 */
  double v_min, v_max;             /* optimizer.py:43-44 */
  double r_sum;             /* radius_ego + radius_obstacle (optimizer.py:439) */
  double ego_offset;        /* ego circle-centre offset along heading, 0.75 m (configuration.py:80-91) */
  double obstacle[6];       /* obstacle circle centres: centre, front, rear (optimizer.py:60-64) */
  double mu0, mu_min, mu_factor;   /* barrier schedule */
  double tol_step, tol_feas;       /* convergence: inf-norm of the Newton step, l1 norm of the defects */
  double tau_min, bound_push;      /* fraction-to-the-boundary, initial interior push */
  double mu_min_alpha;             /* mu is reduced only after an accepted step length >= this */
  double mu_up_alpha, mu_up_factor, mu_max;   /* barrier warm-up: mu *= mu_up_factor (<= mu_max) while the first steps are blocked below mu_up_alpha */
  double mu_factor_full;           /* barrier reduction factor after a FULL primal and dual step (alpha = 1): <= mu_factor */
  double kappa_sigma;              /* multipliers kept within [mu/(kappa s), kappa mu/s] (IPOPT kappa_sigma) */
  double screen_inv_curv;          /* obstacle rows whose barrier curvature mu/s^2 is below 1/this are skipped for the iteration (<= 0: keep all) */
  double trust_step;               /* feasible iterate + step below this: the Newton step is accepted without the merit test */
  double acc_factor;               /* acceptable exit: acc_iters consecutive steps <= acc_factor * tol_step at mu_min */
  int32_t acc_iters;
  int32_t stall_iters;             /* status 3 after this many iterations at mu_min without halving the step (noise floor) */
  int32_t refine_f64;              /* precision F32 only (default 1): instances that did not reach status 1 are queued on the device and re-solved in float64 arithmetic by a second, persistent launch */
  int32_t init_rollout;            /* 1: initial states = Euler rollout of the initial controls from X_0 (X warm start ignored) */
  /* ---- ABI 4 */
  double mu_warm;                  /* dual warm start: barrier parameter a warm-started solve restarts at (instead of mu0) */
  double warm_push;                /* dual warm start: relative interior push of the warm primal point (instead of bound_push) */
  double kappa_warm;               /* dual warm start: carried multipliers are kept within [mu_warm/(kappa s), kappa mu_warm/s] */
  double stiff_slack;              /* float32: an acceptable-level exit while a live obstacle row has a slack below this is reported as status 3 (refined in float64 when refine_f64 is set) */
  int32_t warm_duals;              /* mpcb200_closed_loop: 1 = carry slacks / multipliers across MPC steps (shifted one stage), 0 = IPOPT-like restart every step */
  int32_t warps_per_cta;           /* problems (= warps) per CTA: 0 = library default, else 1 | 2 | 4; 8 | 16: the phase-aligned kernel
                                      (one CTA barrier per SQP iteration, csrc/aligned_solver.cu; float32 Gauss-Newton fused solves only,
                                      bit-identical results, measured no faster: profiles/r02_summary.md section 12) */
  int32_t host_route;              /* mpcb200_solve_host: 0 = zero-copy when every buffer is pinned, else staged; 1 = always staged */
  int32_t host_chunks;             /* staged host route: chunks of the copy / solve / copy pipeline, 0 = by batch size */
} mpcb200_config;

typedef struct mpcb200_handle mpcb200_handle;

/* Fill `cfg` with the reference's constants (vehicle 2 bounds, IPOPT-like tolerances for the chosen precision). */
void mpcb200_default_config(mpcb200_config* cfg, int32_t N, int32_t precision);

/* Create / destroy a solver handle (owns only its scratch). One handle per stream/GPU; not thread-safe. */
int mpcb200_create(const mpcb200_config* cfg, mpcb200_handle** out);
void mpcb200_destroy(mpcb200_handle* h);
const char* mpcb200_last_error(const mpcb200_handle* h);   /* h may be NULL: error of the last failed create */

/* Replaces optimizer.py:605-607 for B independent ego instances: one fused kernel runs ALL SQP iterations of its
 * problems (each lane iterates until its own problem converges).  d_iters/d_status may be NULL. */
int mpcb200_solve(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U,
                  int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* Same solve with the inequality multipliers in and out (SURVEY 8b `d_lam`; IPOPT's `lam_x` / `lam_g` of the bound and obstacle
 * rows, which the reference never reads, optimizer.py:609).  d_lam [B][mpcb200_lam_words] float64, per problem:
 *   [N][14]  stage k: multipliers of deltaDot >= min, <= max, a <= a_max, a >= -sqrt(a_max - s0) (stage 0 only: friction row),
 *            delta_{k+1} >= min, <= max, v_{k+1} >= min, <= max, the three obstacle rows of x_{k+1}; then their three slacks
 *   [1]      barrier parameter at exit          [1]  1.0 = block valid
 * OUT: always written.  IN: a valid block (from a previous call, shifted by the caller as it shifts X / U) warm-starts slacks
 * and multipliers and restarts the barrier at cfg.mu_warm instead of cfg.mu0; anything else (e.g. zeros) = cold duals. */
int mpcb200_solve_dual(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, double* d_lam,
                       int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);
int32_t mpcb200_lam_words(const mpcb200_handle* h);          /* 14 N + 2 */

/* Same solve from the reference's step-0 initial guess (X_0 tiled, zero controls, optimizer.py:578-583): d_X / d_U are
 * outputs only, nothing but d_xref is read. */
int mpcb200_solve_cold(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U,
                       int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* Same solve, one kernel launch per SQP iteration with the per-problem KKT slab (iterate, multipliers, slacks,
 * Riccati blocks) staged HBM -> shared memory by TMA bulk copy and written back each launch:
 *   begin: load problem data, initialise;  iter: `n_iter` iterations per call;  end: write X, U, status, iters.
 * No float64 refinement pass in this mode (cfg.refine_f64 is honoured by solve / solve_cold / solve_dual / solve_host /
 * solve_scenarios): a float32 instance that stalls on an active obstacle row keeps status 3. */
int mpcb200_sqp_begin(mpcb200_handle* h, const double* d_xref, const double* d_X, const double* d_U, int32_t B, void* cuda_stream);
int mpcb200_sqp_iter(mpcb200_handle* h, int32_t n_iter, void* cuda_stream);
int mpcb200_sqp_end(mpcb200_handle* h, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters, void* cuda_stream);

/* shift_movement (optimizer.py:645-655): x+ = x + dt*f(x, U[:,0]); shift U and X one stage, repeating the last.
 * d_x [B][5] in/out (current state), d_U/d_X in/out, d_u_applied [B][2] out (may be NULL). */
int mpcb200_plant_step_shift(mpcb200_handle* h, double* d_x, double* d_U, double* d_X, double* d_u_applied,
                             int32_t B, void* cuda_stream);

/* desired_command_and_trajectory (optimizer.py:657-702): build X_ref for MPC step `i` from the resampled path.
 * d_path [T][2], d_orientation [T], d_x [B][5] current states -> d_xref [B][N+1][5]. */
int mpcb200_build_ref_window(mpcb200_handle* h, int32_t i, int32_t iter_length, const double* d_path,
                             const double* d_orientation, double desired_velocity, const double* d_x,
                             double* d_xref, int32_t B, void* cuda_stream);

/* The whole receding-horizon loop of CasadiOptimizer.optimize() (optimizer.py:596-631) on the device, no host
 * sync between MPC steps: for i in 0..iter_length-1: solve, record u_0, plant step + shift, next window.
 * d_x0 [B][5]; outputs d_traj [B][iter_length][5] (Q12: x0 first, last dropped), d_ctrl [B][iter_length][2],
 * d_status [B][iter_length], d_iters [B][iter_length] (may be NULL).
 * cfg.warm_duals = 1 carries slacks / multipliers across the MPC steps (default 0 = restart every step like IPOPT in the
 * reference).  No float64 refinement pass inside the loop: with float32 arithmetic and an obstacle within reach, stalled steps
 * are reported as status 3 -- use precision F64 for such closed loops (B200Optimizer.optimize() routes them through the
 * per-step host loop, where every solve has its refinement launch). */
int mpcb200_closed_loop(mpcb200_handle* h, int32_t iter_length, const double* d_path, const double* d_orientation,
                        double desired_velocity, const double* d_x0, double* d_traj, double* d_ctrl,
                        int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* Host-buffer entry point (the end-to-end path); synchronous on return.  h_X_out / h_U_out may alias h_X / h_U (in
 * place).  h_X = h_U = NULL: cold start (X_0 tiled, zero controls), only xref is read.
 *   - every data buffer pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory()): ZERO-COPY route -- one launch
 *     whose TMA bulk copies read xref (+ warm start) from and write X / U to host memory directly over PCIe, so each
 *     problem's transfer overlaps the other problems' iterations and no staging copy sits on the critical path;
 *   - otherwise (pageable buffers, or MPCB200_HOST_STAGED=1): H2D, solve and D2H through library-owned device staging,
 *     pipelined in chunks over internal streams.  Same arithmetic, bit-identical results either way. */
int mpcb200_solve_host(mpcb200_handle* h, const double* h_xref, const double* h_X, const double* h_U,
                       double* h_X_out, double* h_U_out, int32_t* h_status, int32_t* h_iters, int32_t B);

/* Per-problem scenarios (SURVEY 8b `optimize_batch(x0, scenario_id[B])`, BASELINE configs[4]): one launch over problems that belong
 * to different scenarios.  A row of the table holds what `Optimizer.__init__` takes from a scenario's `configuration`
 * (optimizer.py:51-68): time step, weights, obstacle circles and r_ego + r_obs; bounds, horizon and solver options stay the
 * handle's.  `mpcb200_set_scenarios` copies `n` rows (HOST pointer) to the device; `mpcb200_solve_scenarios` is
 * `mpcb200_solve_cold` with the constants of problem b taken from row d_scenario_id[b] (device int32 [B]; ids are clamped to
 * the table).  Gauss-Newton Hessian only.  A float32 handle with refine_f64 runs the float64 refinement pass as well. */
typedef struct mpcb200_scenario {
  double dt;             /* configuration.delta_t */
  double Q[5], R[2];     /* weights_setting, order as in mpcb200_config */
  double r_sum;          /* radius_ego + radius_obstacle */
  double obstacle[6];    /* obstacle circle centres: centre, front, rear */
} mpcb200_scenario;
int mpcb200_set_scenarios(mpcb200_handle* h, const mpcb200_scenario* table, int32_t n);
int mpcb200_solve_scenarios(mpcb200_handle* h, const double* d_xref, const int32_t* d_scenario_id, double* d_X, double* d_U,
                            int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* Stage functions of the reference's FORCESPRO formulation and their first derivatives -- the linearisation one SQP stage on that
 * formulation needs (ForcesproOptimizer: RK4 dynamics optimizer.py:90-98, friction circle + nine squared circle distances
 * :119-155, stage / terminal least-squares objective :163-195).  Replaces the generated model callbacks the FORCESPRO solver
 * calls per stage: FORCESNLPsolver_{dynamics,ddynamics,inequalities,dinequalities,objective,dobjective}_0 and _1
 * (test/FORCESNLPsolver/FORCESNLPsolver_model.c:74-1756, dispatched by FORCESNLPsolver_interface.c:84-191).
 *   d_z [n][7]  stage variables [deltaDot, aLong, xPos, yPos, delta, v, psi]
 *   d_p [n][10] stage parameters [path_x, path_y, v_des, psi_ref, obstacle centre / front / rear circle x, y]
 *   d_out [n][136] = c(5) | dc/dz (5x7 row-major) | h(10) | dh/dz (10x7) | f | df/dz(7) | f_terminal | df_terminal/dz(7)
 * Stage weights Q, R, dt, wheelbases and the ego circle offset come from the handle's config; `weights_terminal` [5] (host
 * pointer) are the weight_*_terminate values.  float64 arithmetic.  `mpcb200_forces_solve` below is the solver on top of it. */
int mpcb200_forces_stage_eval(mpcb200_handle* h, const double* weights_terminal, const double* d_z, const double* d_p,
                              double* d_out, int32_t n, void* cuda_stream);

/* The reference's FORCESPRO formulation of the MPC problem (ForcesproOptimizer, optimizer.py:86-246), one NLP per ego instance,
 * solved to convergence on the device (csrc/forces_core.cuh: general Riccati recursion, RK4 Jacobians, friction circle at
 * every stage, nine circle pairs per stage, terminal weights).  Replaces the generated solver's entry point
 *   FORCESNLPsolver_solve(FORCESNLPsolver_params*, FORCESNLPsolver_output*, FORCESNLPsolver_info*, ...)
 * (test/FORCESNLPsolver/include/FORCESNLPsolver.h:108-165, bound by interface/FORCESNLPsolver_py.py:95-179 and called from
 * optimizer.py:320) with the same data, batched:
 *   d_xinit  [B][5]      params.xinit            initial state [xPos, yPos, delta, v, psi]                  (optimizer.py:283)
 *   d_params [B][N][10]  params.all_parameters   stage major: path_x, path_y, v_des, psi_ref, 3 obstacle circle centres (:313-318)
 *   d_z_init [B][N][7]   params.x0               initial guess, stage major [deltaDot, aLong, x(5)]; NULL: xinit tiled, zero inputs
 *   d_z      [B][N][7]   output.x01 .. xN        optimal stage variables; row 0's state is xinit                  (:328-336)
 *   d_status [B]         exitflag                1 optimal, 0 iteration limit, 3 stalled at the float32 noise floor, -6 / -7 / -8
 *   d_iters  [B]         info.it
 * N = cfg.N is model.N (the number of stages, optimizer.py:204); stage weights, bounds (a_max is both the symmetric
 * acceleration bound and the friction-circle radius, optimizer.py:108-111), dt, wheelbases, r_sum and the ego circle offset
 * come from the handle's config; `weights_terminal` [5] (HOST pointer) are the weight_*_terminate values.  The obstacle
 * circle centres are per stage (the reference tiles one obstacle over the stages; a moving obstacle just changes the rows).
 * Deviations that do not change the solution set: the vacuous lower bound 0 <= aLong^2 + (v psiDot)^2 is not a row; the last
 * stage's inputs (no cost term, no dynamics) are returned as their minimum-norm optimum 0; the circle rows are worked on as
 * distance >= r_sum instead of distance^2 >= r_sum^2.  A float32 handle with refine_f64 runs the float64 pass on the
 * stragglers.  Unlike FORCESPRO's SQP_NLP with maxqps = 1 (one QP per call, optimizer.py:237) the NLP is solved to
 * convergence. */
int mpcb200_forces_solve(mpcb200_handle* h, const double* weights_terminal, const double* d_xinit, const double* d_params,
                         const double* d_z_init, double* d_z, int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* ForcesproOptimizer.optimize()'s receding-horizon loop (optimizer.py:286-362) for B egos in ONE launch, one ego per warp: per
 * closed-loop step k the `all_parameters` rows are built on the device (path points / headings k+1 .. k+N replenished with the
 * last one, :296-311; `d_velocity` [iter_length] is the desired-velocity profile of :291-294; the handle's obstacle circles tiled,
 * :306-317), the NLP is solved (warm start: the previous solution shifted one stage), the first input is applied to the RK4
 * plant in float64 (model.eq, :359).  d_x0 [B][5] -> d_traj [B][iter_length][5] (state at the start of each step, the reference's
 * `x` without its last column, :366), d_ctrl [B][iter_length][2], d_status / d_iters [B][iter_length].  Noise-free loop (the
 * `noised` option of :347-356 stays on the host: B200ForcesproOptimizer.optimize_batch).  No float64 refinement pass inside
 * the loop: a float32 handle reports status 3 where it stalls. */
int mpcb200_forces_closed_loop(mpcb200_handle* h, const double* weights_terminal, int32_t iter_length, const double* d_path,
                               const double* d_orientation, const double* d_velocity, const double* d_x0, double* d_traj, double* d_ctrl,
                               int32_t* d_status, int32_t* d_iters, int32_t B, void* cuda_stream);

/* Road-boundary rows for `mpcb200_forces_solve` -- the six constraints per stage the reference states and leaves
 * commented out (`find_closest_distance_with_road_boundary`, optimizer.py:18-30; rows :136-161 and :386-410, bounds :113-117,
 * model.nh = 16 :208): for each of the three ego circle centres and each of the two boundaries, the distance to the CLOSEST
 * VERTEX of the boundary polyline (`ca.mmin` over the vertex distances) >= r_min (radius_ego).  `left` / `right`: HOST pointers
 * to [n][2] vertex lists in absolute coordinates (`configuration.left_road_boundary` / `right_road_boundary`,
 * configuration.py:432-433); copied to the device.  n_left = n_right = 0 switches the rows off again (the default). */
int mpcb200_forces_set_road_boundaries(mpcb200_handle* h, const double* left, int32_t n_left, const double* right, int32_t n_right,
                                       double r_min);

/* Introspection for benches/tests. */
int64_t mpcb200_launch_count(const mpcb200_handle* h);       /* kernels launched by this handle so far */
int32_t mpcb200_workspace_words(const mpcb200_handle* h);    /* words of the per-problem KKT slab */
int32_t mpcb200_slab_in_smem(const mpcb200_handle* h);       /* 1 if the slab lives in shared memory */
int32_t mpcb200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPCB200_H */
