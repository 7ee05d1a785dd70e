/* ftmpc.h -- C ABI of libftmpc.so: the B200-native per-timestep MPC solve of ft_mpc.
 *
 * The reference (DISCOWER/fault-tolerant-mpc, pure Python) has no FFI layer.  The boundary this
 * library replaces is the body of
 *     SpiralingController.get_control      ft_mpc/controllers/spiraling_mpc.py:288-317
 *       -> SpiralModel.robot_to_center     ft_mpc/models/spiral_model.py:91-109
 *       -> solve_mpc (CasADi/IPOPT NLP)    ft_mpc/controllers/spiraling_mpc.py:319-354  (NLP: :87-238)
 *       -> ControlAllocator.get_physical_input   ft_mpc/controllers/tools/control_allocator.py:65-95
 * batched over B independent instances (fault scenario x initial state).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions: all tensors are contiguous row-major DEVICE buffers owned by the caller (PyTorch);
 * fp64 arithmetic; every call only enqueues work on `stream` (no hidden synchronisation);
 * return value 0 = ok, <0 = error
 * (ftmpc_strerror); per-instance outcomes are reported in status[B], never by exit()/exceptions.
 * A handle is bound to the device that was current in ftmpc_create; every entry point switches to that device for
 * its launches and restores the caller's current device.  A handle is not thread-safe.
 */
#ifndef FTMPC_H_
#define FTMPC_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTMPC_NX 13      /* centre state [p_c v_c omega q]            spiral_model.py:12-19   */
#define FTMPC_NU 6       /* generalised wrench                         spiral_model.py:131-133 */
#define FTMPC_NE 9       /* optimised error states (pos, vel, omega)   spiraling_mpc.py:57     */
#define FTMPC_NTHR 16    /* physical thrusters                         sys_model.py:55         */
#define FTMPC_NH 26      /* max hull facets per fault set              input_bounds.py:71      */
#define FTMPC_NF 72      /* terminal-set rows                          config/terminal.yaml    */
#define FTMPC_HULL_STRIDE (FTMPC_NH * FTMPC_NU + FTMPC_NH)   /* one hull-table entry: A_h[26][6] then b_h[26] */
#define FTMPC_MAX_POLY 32
#define FTMPC_MAX_ROOT 16

/* per-instance status codes */
enum {
    FTMPC_ST_OK = 0,          /* NLP converged, feasible                                    */
    FTMPC_ST_MAXITER = 1,     /* SQP iteration cap reached                                  */
    FTMPC_ST_QPFAIL = 2,      /* QP sub-problem infeasible / numerical breakdown / NaN      */
    FTMPC_ST_INFEASIBLE = 3,  /* converged to a point that violates the constraints         */
    FTMPC_ST_ALLOC = 4,       /* NLP ok but thrust allocation infeasible (control_allocator.py:88-93 calls exit()) */
    FTMPC_ST_BADINPUT = 5,    /* rejected inputs: hull_idx outside [0, n_hull_sets); zero command returned   */
    FTMPC_ST_RUNNING = -1
};

/* library error codes */
enum {
    FTMPC_OK = 0,
    FTMPC_ERR_ARG = -1,
    FTMPC_ERR_CUDA = -2,
    FTMPC_ERR_WORKSPACE = -3,
    FTMPC_ERR_UNSUPPORTED = -4,
    FTMPC_ERR_NO_DEVICE = -5,
    FTMPC_ERR_NOMEM = -6
};

typedef struct ftmpc_config {
    int32_t horizon;           /* N                                   reactive.yaml:26, spiraling_mpc.py:38 */
    int32_t dtype;             /* 0 = fp64 (only mode implemented)                                          */
    int32_t max_sqp_iter;      /* outer iteration cap (default 60).  Of the three bench scenarios (of 8192) that hit it, two reach
                                  the oracle's KKT point after 64 / 68 iterations and one has two local minima
                                  (tools/hard_instances.py); a cap of 80 converts the two but costs 2.6 % of the 8-GPU throughput
                                  (every hopeless instance then runs a third longer at the tail of its launch)           */
    int32_t max_qp_iter;       /* active-set iteration cap per QP (default 20*(n+m))                        */
    int32_t stall_window;      /* stall detector (sqp_stalled): give up (FTMPC_ST_MAXITER) when the step has not halved over this many
                                  iterations, checked from 2 windows on (default 10; 0 = run to max_sqp_iter)                  */
    int32_t warm_qp;           /* 1 = start each QP from the previous QP's active set (gi_warm_start): -40 % active-set iterations, same
                                  results; default 0 -- on the B200 the bulk update currently costs what it saves (profiles/README.md) */
    int32_t n_poly, n_root, n_hull_sets;
    int32_t qp_method;         /* bit field; 0 = default.  The QP of every SQP iteration is factorised stage by stage (Riccati
                                  recursion, csrc/ftmpc_riccati.cuh) and solved by a dual active-set method in
                                    - null-space form, J = Phi^-1 blkdiag(C_t^-T) rotated in shared memory (ftmpc_gi.cuh), one CTA per
                                      SM, for horizons N <= 20;
                                    - operator form for N > 20: range-space iteration, K = E E' applied through the stage records,
                                      nothing of size N^2 in memory (ftmpc_gis.cuh: gis_solve_op, ric_apply_g).
                                  bits 0-1: 1 = experimental k_solve2 (two CTAs per SM, range-space form on the packed dense K,
                                            ftmpc_qp2.cuh; N <= 20), 2 = as 1 and the CPU checker runs its dense range-space prototype;
                                  bit 3 (8):  long horizons condense with the O(N^3) form (only with bit 4);
                                  bit 4 (16): round-1 factorisation (condensed Hessian + Cholesky + L^-T) -- host builds and CUDA
                                              builds made with -DFTMPC_DENSE_FACTOR only (A/B measurements);
                                  bit 5 (32): operator form at every horizon;   bit 6 (64): never the operator form.           */
    double dt, mass, inertia[3], r[3], f_virt[3], max_thrust;   /* sys_model.py:52-61, spiral_parameters.py:33-39 */
    double Q[FTMPC_NE], R[FTMPC_NU];                            /* reactive.yaml:32-33                      */
    double D[FTMPC_NU * FTMPC_NTHR];                            /* allocation matrix, sys_model.py:73-123   */
    double Af[FTMPC_NF * FTMPC_NE], bf[FTMPC_NF];               /* terminal set                             */
    /* terminal cost  V_f(e) = c0 + sum_k c_k prod e^p_k + sum_j d_j (prod e^q_j + eps_j)^w_j              */
    double term_const;
    double poly_c[FTMPC_MAX_POLY];
    int8_t poly_e[FTMPC_MAX_POLY][FTMPC_NE];
    double root_c[FTMPC_MAX_ROOT], root_eps[FTMPC_MAX_ROOT], root_pow[FTMPC_MAX_ROOT];
    int8_t root_e[FTMPC_MAX_ROOT][FTMPC_NE];
    double term_quad[FTMPC_NE * FTMPC_NE];   /* Hessian of the pure-quadratic part of V_f (PSD): Gauss-Newton model */
    double sqp_tol;            /* stop when |step|_inf <= sqp_tol (default 1e-8)                            */
    double qp_tol;             /* primal feasibility tolerance of the active-set QP (default 1e-10)         */
    double feas_tol;           /* NLP constraint violation accepted at convergence (default 1e-7)           */
    double act_tol;            /* row i is "active" iff b_i - g_i(z*) <= act_tol (default 1e-7)             */
    double rho_slack;          /* weight of the QP elastic variable (default 1e4)                           */
    double clip_tol;           /* hull membership tolerance of clip_generalized_input (default 1e-9)        */
    double theta_first;        /* Hessian schedule: blend of exact second-order terms in SQP iteration 1 (iteration 0 is   */
    double theta_growth;       /* Gauss-Newton), multiplied by theta_growth per iteration up to 1 (defaults 0.5, 2)       */
    double blend_dmax;         /* the blend is first attempted once the QP step satisfies |d|_inf <= blend_dmax (default 1) */
    double fast_dmax;          /* early stop in the quadratic regime (sqp_fast_converged): after two consecutive exact-Hessian full
                                  steps with |d_k| <= fast_dmax and predicted next step |d_k|^3/|d_{k-1}|^2 <= 0.1 sqp_tol the
                                  confirming QP is skipped (default 1e-5; 0 = always confirm)                              */
} ftmpc_config;

typedef struct ftmpc_ctx* ftmpc_handle;

/* hull_table: HOST pointer, n_hull_sets * FTMPC_HULL_STRIDE doubles (copied to the device). */
int ftmpc_create(ftmpc_handle* out, const ftmpc_config* cfg, const double* hull_table);
void ftmpc_destroy(ftmpc_handle h);
const char* ftmpc_strerror(int code);
int ftmpc_workspace_bytes(ftmpc_handle h, int batch, size_t* out);

/* dimensions derived from the horizon */
int ftmpc_num_var(ftmpc_handle h);        /* 6N + 13(N+1)   spiraling_mpc.py:110-114            */
int ftmpc_num_ineq(ftmpc_handle h);       /* 26N + 72       spiraling_mpc.py:175-177,199-202    */

/* One MPC step for `batch` instances == SpiralingController.get_control (spiraling_mpc.py:288-317).
 *   state      [B,13]  robot state [p v q omega]                      (sim_env.py:82)
 *   xref       [B,N+1,9], uref [B,N+1,6] or NULL (= 0)               (spiraling_mpc.py:356-365)
 *   fault_mask [B] bit i = thruster i failed; fault_force [B,16] stuck-on force (sys_model.py:239)
 *   hull_idx   [B] row of the hull table for this fault set          (input_bounds.py:43-76)
 *   z_warm     [B, 6N+13(N+1)] in/out: optimal [u_0..u_{N-1} | x_0..x_N]; warm != 0 -> shifted
 *              previous solution is the initial guess (spiraling_mpc.py:324-331), else u = 0
 *   thrust     [B,16] out   u0 [B,6] out   active_set [B, ceil((26N+72)/32)] bit mask out
 *   status     [B] out      iters [B,2] out (SQP iterations, total QP iterations)
 */
int ftmpc_step(ftmpc_handle h, int batch, const double* state, const double* xref, const double* uref,
               const uint16_t* fault_mask, const double* fault_force, const int32_t* hull_idx, int warm,
               double* z_warm, double* thrust, double* u0, uint32_t* active_set, int32_t* status,
               int32_t* iters, double* cost, void* workspace, size_t workspace_bytes, void* stream);

/* Closed loop for `batch` instances == SimulationEnvironment.run_simulation (sim_env.py:77-112): `steps` times
 * (ftmpc_step -> plant step -> noise -> quaternion renormalisation), enqueued back to back on `stream` with no host
 * round trip.  Every instance tracks the SAME reference table (SpiralingController.assign_trajectory, :255-286):
 *   trajectory [table_rows,9], nominal [table_rows,6] or NULL; the window of loop step k starts at row start_step + k
 *   (the integer form of int(t/dt), :360) and is read in place -- no per-instance copies;
 *   state [B,13] in/out (robot states); noise [steps,B,13] or NULL (the reference draws unseeded U(0,1e-3), sim_env.py:88-91);
 *   warm start from the second step on, from the first too when warm_first != 0 (spiraling_mpc.py:324-331);
 *   per-step outputs as in ftmpc_step (they hold the LAST step on return); cost_sum [B] = sum of the optimal costs,
 *   worst_status [B] = max status over the steps. */
int ftmpc_closed_loop(ftmpc_handle h, int batch, int steps, int start_step, int table_rows, double* state,
                      const double* trajectory, const double* nominal, const uint16_t* fault_mask,
                      const double* fault_force, const int32_t* hull_idx, const double* noise, int warm_first,
                      double* z_warm, double* thrust, double* u0, uint32_t* active_set, int32_t* status, int32_t* iters,
                      double* cost, double* cost_sum, int32_t* worst_status, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Profiling hooks used by bench.py.  When enabled, ftmpc_step brackets its kernels with CUDA events on `stream` and
 * the solver kernel accumulates a per-phase cycle profile.  ftmpc_profile_read synchronises the stream and returns, for
 * the LAST ftmpc_step:  kernel_ms[0] = k_solve, kernel_ms[1] = k_alloc (device time);
 * phase_cycles[i] = SM cycles summed over all CTAs spent in phase i:
 *   0 step acceptance + rollout, 1 linearisation (Jacobians, costates, stage Hessians), 2 condensing, 3 Cholesky,
 *   4 J = L^-T, 5 QP set-up, 6 dual active-set iterations, 7 QP post-processing, 8 result write-out. */
#define FTMPC_N_PHASES 40   /* 9 coarse phases, 20 sub-phases, 11 event counters: names in ft_mpc_b200._lib.PHASE_NAMES */
int ftmpc_profile_enable(ftmpc_handle h, int enable);
int ftmpc_profile_read(ftmpc_handle h, void* stream, double* kernel_ms, int64_t* phase_cycles, int n_phase);
/* number of kernels ftmpc_step launched on its last call */
int ftmpc_last_launches(ftmpc_handle h);

/* Measurement helper for bench.py: sustained FP64 FMA throughput of this device (TFLOP/s, 2 flop per DFMA),
 * the roofline denominator of the solve (MEASURED_PEAKS.json carries no fp64 figure).  Synchronises. */
int ftmpc_fp64_peak(ftmpc_handle h, double* tflops, void* stream);

/* ---- stage entry points (unit parity tests; same device code as ftmpc_step) ------------------ */
/* K1: RK4 rollout + Jacobians, one warp per instance.  wrench [B,N,6] = total body wrench per stage.
 *     x [B,N+1,13] (x[:,0] given), jac [B,N,13,13] column-major per stage in z-order [w q F tau]
 *     (column c holds d x_{t+1} / d z_c), lam [B,N+1,13] or NULL, hess [B,N,13,13] or NULL.
 *     (SpiralModel.dx_dt spiral_model.py:44-76, rk4_integrator sys_model.py:138-162)             */
int ftmpc_rk4_jac(ftmpc_handle h, int batch, double* x, const double* wrench, double* jac, const double* lam,
                  double* hess, void* stream);
/* robot state -> centre state, spiral_model.py:91-109 */
int ftmpc_robot_to_center(ftmpc_handle h, int batch, const double* state, double* center, void* stream);
/* terminal cost value / gradient / Hessian, e [B,9] -> V [B], grad [B,9], hess [B,81] */
int ftmpc_terminal(ftmpc_handle h, int batch, const double* e, double* V, double* grad, double* hess, void* stream);
/* K2: condensed Hessian/gradient at the linearisation held in the workspace after ftmpc_step
 *     (debug_theta = blend of exact second-order terms, 0 = Gauss-Newton): H [B,6N,6N], g [B,6N] */
int ftmpc_condense(ftmpc_handle h, int batch, const double* jac, const double* hess, const double* x,
                   const double* u, const double* xref, const double* gradV, const double* hessV, double theta,
                   double* H, double* g, void* stream);
/* K3: batched dense QP  min 1/2 x'Hx + g'x  s.t. C x <= b  (n <= 121, rows of C with <= 12 non-zeros, CSR) */
int ftmpc_qp_solve(ftmpc_handle h, int batch, int n, int m, const double* H, const double* g, const int32_t* row_ptr,
                   const int32_t* col_idx, const double* val, const double* b, double* x, double* lam,
                   int32_t* status, void* stream);
/* K5: control allocation  min |u|^2 s.t. D u = u_des, 0 <= u <= ub   (control_allocator.py:28-40) */
int ftmpc_allocate(ftmpc_handle h, int batch, const double* u_des, const double* ub, double* thrust,
                   int32_t* status, void* stream);
/* clip_generalized_input (control_allocator.py:42-63): u [B,6] = u_res + D f_fault -> u_clipped [B,6]; identity when
 *     A_h u <= b_h + clip_tol, else the Euclidean projection onto the hull of hull_idx[B] (the reference's branch is
 *     dimensionally inconsistent and names an absent solver; its docstring states this projection).  status 0 / GI code. */
int ftmpc_clip(ftmpc_handle h, int batch, const int32_t* hull_idx, const double* u, double* u_clipped, int32_t* status,
               void* stream);
/* K6: plant step (16 thrusters, robot state): RK4 of SystemModel.dx_dt (sys_model.py:138-243);
 *     noise [B,13] or NULL is added after the step (sim_env.py:88-91); normalize != 0 renormalises the
 *     quaternion (sys_model.py:164-175, sim_env.py:93).  normalize = 0, noise = NULL == model.dynamics(x,u). */
int ftmpc_plant_step(ftmpc_handle h, int batch, const double* state, const double* thrust,
                     const uint16_t* fault_mask, const double* fault_force, const double* noise, int normalize,
                     double* next, void* stream);

/* Input-bound polytope of n_sets fault sets on the device (analytic replacement of InputBounds.calc_input_bounds,
 *     input_bounds.py:43-76: zonotope facet enumeration instead of Qhull on 2^14 corner wrenches).
 *     fault_mask [n_sets], fault_force [n_sets,16] (intensity * max_thrust at failed thrusters) ->
 *     table [n_sets, FTMPC_HULL_STRIDE] in the layout ftmpc_create takes (A_h 26x6 then b_h 26, padding rows 0 / 1e30),
 *     n_rows [n_sets], status [n_sets] (0 ok, 1 = rank-deficient fault set or too many facets: use the host routine).
 *     Rows come in a canonical order (lexicographic on [A | -b]); the reference's order follows Qhull's rounding noise. */
int ftmpc_hull_facets(ftmpc_handle h, int n_sets, const uint16_t* fault_mask, const double* fault_force, double* table,
                      int32_t* n_rows, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FTMPC_H_ */
