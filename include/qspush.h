/* include/qspush.h — C-ABI of the B200-native batched pusher-slider NMPC engine.
 *
 * This is the drop-in boundary for the acados v0.2.1 solver object that
 * /root/reference/acados_nmpc/NMPC_controller.m drives through acados' MATLAB class `acados_ocp`
 * (NMPC_controller.m:304 create, :334-348 set, :382-384 init, :389 solve, :392-394,:403 get,
 * :420 get_cost; helper.m:253,264-269 status/statistics), plus the numeric entry points of
 * bspline_shape / PusherSliderModel that the controller and the closed loop call
 * (bspline_shape.m:146-152,192-199; PusherSliderModel.m:606-608).
 *
 * Conventions
 *   - plain C, opaque handles, caller-owned buffers, FP64 everywhere;
 *   - every function returns 0 on success, a negative qspush_err on argument / runtime errors;
 *     qspush_solve never fails because a problem did not converge — poll QSPUSH_STATUS per problem
 *     (acados v0.2.1 enum: 0 success, 1 NaN/failure, 2 max iter, 3 min step, 4 QP failure);
 *   - a solver is bound to one CUDA device and owns one stream; calls on one handle must be
 *     serialised by the caller (same contract as one acados_ocp object);
 *   - batched per-problem arrays use the MATLAB-natural layout  [batch][stage][dim]  (i.e. the
 *     column-major dim x stages x batch array); with stage >= 0 the layout is [batch][dim];
 *   - `mem` says where the caller's buffer lives (host or device of the solver's GPU);
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef QSPUSH_H
#define QSPUSH_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qspush_model qspush_model;
typedef struct qspush_solver qspush_solver;

typedef enum {
    QSPUSH_OK = 0,
    QSPUSH_ERR_ARG = -1,      /* bad pointer / size / field / range               */
    QSPUSH_ERR_CUDA = -2,     /* CUDA runtime error (see qspush_last_error)       */
    QSPUSH_ERR_IO = -3,       /* .ply could not be read / parsed                  */
    QSPUSH_ERR_NO_DEVICE = -4 /* no CUDA device: the engine has no CPU path       */
} qspush_err;

typedef enum { QSPUSH_MEM_HOST = 0, QSPUSH_MEM_DEVICE = 1 } qspush_mem;

/* solver modes: acados nlp_solver "sqp_rti" / "sqp" (NMPC_controller.m:272) */
typedef enum { QSPUSH_MODE_RTI = 0, QSPUSH_MODE_SQP = 1 } qspush_mode;

/* Options; zero-initialise then call qspush_opts_default. Mirrors create_ocp_opts
 * (NMPC_controller.m:270-300) plus the acados defaults the reference does not override. */
typedef struct {
    int    mode;                    /* qspush_mode                                            */
    int    max_sqp_iter;            /* nlp_solver_max_iter = 30                               */
    double tol_stat, tol_eq, tol_ineq, tol_comp;   /* nlp_solver_tol_* = 1e-6                 */
    int    qp_max_iter;             /* qp_solver_iter_max = 50                                */
    double qp_tol;                  /* IPM tolerance on stationarity / dynamics / inequality residuals (1e-11) */
    double qp_mu0;                  /* IPM initial barrier parameter                          */
    double qp_thr;                  /* IPM lower clamp on the initial slacks                  */
    double qp_tau;                  /* IPM fixed fraction to the boundary (only used when qp_gamma_f = 0) */
    int    globalization;           /* 1 = merit backtracking, 0 = fixed full step            */
    double alpha_min, alpha_reduction, eps_sufficient_descent;   /* 0.05, 0.7, 1e-4           */
    int    matlab_single_quirk;     /* reproduce MATLAB `single` rounding of mod(s,b) in prepare */
    int    problems_per_warp;       /* thread-per-problem QP kernel packing: 32, 16, 8, 4 (0 = auto) */
    int    qp_kernel;               /* 2 = auto (default: the warp kernel whenever N <= 127); 1 = warp kernel, parallel-in-time Riccati, one or two problems per warp (N <= 127); 0 = one problem per thread */
    int    h_variant;               /* constraint set h of the OCP: 0 = [s; u_n; u_t] (NMPC_controller.m:237, default);
                                       1 = the authors' parked set [u_n; u_t - v_bound(s); u_t + v_bound(s)] (:226-238) with
                                       v_bound from qspush_ctrl; selecting it resets constr_lh / constr_uh to
                                       [u_n_lb, -2 u_t_ub, 0] / [u_n_ub(0.03), 0, 2 u_t_ub] (:247-248).  Both QP kernels
                                       implement it (r02: the one-problem-per-thread kernel too, so any horizon)            */
    /* IPM end game (appended in r02; the fields above keep their offsets).  Complementarity is driven far below the other
     * residuals: multipliers of this QP are as small as 1e-9 (input weight 5e-5), so lam*t <= 1e-12 would leave du 1e-5 away
     * from the QP solution; at 1e-18 the returned point is within 1e-9 of it whatever path the IPM took. */
    double qp_tol_comp;             /* tolerance on max lam * t (1e-18)                                        */
    double qp_t_min;                /* slack floor (1e-12): pairs with t <= 4 t_min count as converged, centering target lam * t_min */
    double qp_gamma_f;              /* step to the boundary: blocking pair keeps gamma_f * predicted mu reduction (0.01; 0 = fixed qp_tau);
                                       primal (z, t) and dual (pi, lam) steps go to their own boundaries with that fraction            */
    int    qp_stall;                /* iterations without halving the normalised residual before a point below 1e-6 is accepted (10) */
} qspush_opts;

/* controller-side constants of NMPC_controller (NMPC_controller.m:23-26, 98-100) */
typedef struct {
    double v_alpha, d_v_bound, t_angle0, u_t_ub, u_n_lb;
} qspush_ctrl;

typedef enum {
    /* per problem, double */
    QSPUSH_X0 = 0,        /* constr_x0, 4                          NMPC_controller.m:170,334 */
    QSPUSH_YREF = 1,      /* cost_y_ref, 6 per stage 0..N-1         :345                      */
    QSPUSH_YREF_E = 2,    /* cost_y_ref_e, 4                        :348                      */
    QSPUSH_X = 3,         /* init_x / get('x'), 4 per stage 0..N    :382,:393                 */
    QSPUSH_U = 4,         /* init_u / get('u'), 2 per stage 0..N-1  :383,:392,:403            */
    QSPUSH_PI = 5,        /* init_pi / get('pi'), 4 per stage 0..N-1  :384,:394               */
    QSPUSH_LAM = 6,       /* inequality multipliers, 6 per stage 0..N-1: [lower(s,un,ut); upper] */
    QSPUSH_COST = 7,      /* get_cost, 1                            :420                      */
    QSPUSH_RES = 8,       /* residuals [stat, eq, ineq, comp], 4                              */
    /* shared by the whole batch, double (batch_lo/batch_hi ignored) */
    QSPUSH_W = 16,        /* cost_W: stage k<N 6x6 col-major in y=[x;u] order; k==N 4x4  :154,157 */
    QSPUSH_LH = 17,       /* constr_lh, 3                           :137,:251                 */
    QSPUSH_UH = 18,       /* constr_uh, 3                           :138,:252                 */
    /* per problem, int (qspush_get_int / qspush_set_int) */
    QSPUSH_STATUS = 32,   /* helper.m:253                                                     */
    QSPUSH_SQP_ITER = 33, /* helper.m:264                                                     */
    QSPUSH_QP_ITER = 34,  /* total IPM iterations of the last solve                           */
    QSPUSH_OBJECT_ID = 35,/* index into the model list given at creation                      */
    QSPUSH_COLD = 36      /* 1 = no previous solution (isempty(utraj), NMPC_controller.m:351) */
} qspush_field;

/* scalar statistics of the last qspush_solve (helper.m:264-269), seconds, CUDA-event timed; in full-SQP mode TIME_LIN
 * (linearisation + residuals) and TIME_QP (QP + line search) are sums over the SQP iterations, like acados' */
typedef enum { QSPUSH_TIME_TOT = 0, QSPUSH_TIME_LIN = 1, QSPUSH_TIME_QP = 2, QSPUSH_TIME_PREP = 3 } qspush_stat;

const char* qspush_last_error(void);
const char* qspush_version(void);
int qspush_device_count(void);

/* ---------------- model: PusherSliderModel + bspline_shape ---------------- */

/* From tables: knots S (n+p+1), control points (n x 2 row-major), degree p, slider/pusher
 * friction and limit-surface constant (bspline_shape.m:25-38; PusherSliderModel.m:53-55).
 * single_coeffs != 0 reproduces the MATLAB `single` arithmetic of the derivative-coefficient
 * tables when S and P come from pcread. */
int qspush_model_create(const double* knots, int nknots, const double* ctrl_xy, int n, int degree,
                        double mu_sp, double c_ellipse, int single_coeffs, qspush_model** out);

/* From an objects_database entry and its .ply outline (PusherSliderModel.m:45-60, 84-132):
 * binary_little_endian float32 vertices -> sortCadPoints (flip_order for montana/pulirapid,
 * :107-109) -> getSpline. */
int qspush_model_create_from_ply(const char* ply_path, int flip_order, int degree, double mu_sg,
                                 double mu_sp, double mass, double tau_max, qspush_model** out);

void qspush_model_free(qspush_model* m);

/* sizes and constants: n control points, nknots, total length b, c_ellipse, mu_sp */
int qspush_model_info(const qspush_model* m, int* n, int* nknots, double* b, double* c_ellipse, double* mu_sp);
/* copy out knots (nknots), control points (n x 2), cj_1 (n x 2), cj_2 (n x 2); any may be NULL */
int qspush_model_tables(const qspush_model* m, double* knots, double* ctrl_xy, double* c1, double* c2);

/* ---------------- stateless batched evaluation (config 2 entry points) ----------------
 * cnt samples; all pointers are in `mem`; outputs may be NULL to skip.  Arrays are [cnt][dim].
 * wrap: 0 none, 1 MATLAB mod(s,b) (evalSpline, bspline_shape.m:193), 2 fmod(s,b)+(s<0)b (PusherSliderModel.m:526) */
int qspush_eval_spline(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* s, int wrap,
                       double* C, double* Cd, double* Cdd, double* tvers, double* nvers, double* kappa);
/* xdot = f(x,u) (evalModelVariableShape, PusherSliderModel.m:606); Jx [cnt][16], Ju [cnt][8] row-major, optional */
int qspush_eval_dynamics(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* x, const double* u,
                         double* f, double* Jx, double* Ju);
/* one ERK4 step with forward sensitivities (acados sim_erk): Phi [cnt][4], A [cnt][16], B [cnt][8] row-major */
int qspush_eval_erk4_sens(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* x, const double* u,
                          double dt, double* Phi, double* A, double* B);
/* v_bound and |kappa| of update_tangential_velocity_bounds (NMPC_controller.m:319-327); outputs [cnt] */
int qspush_eval_v_bound(const qspush_model* m, int device, qspush_mem mem, int cnt, const double* s,
                        const qspush_ctrl* ctrl, int single_quirk, double* v_bound, double* t_angle);

/* ---------------- batched OCP solver: the acados_ocp replacement ---------------- */

void qspush_opts_default(qspush_opts* o);
void qspush_ctrl_default(qspush_ctrl* c);

/* models[nmodels]: the objects a batch may mix (QSPUSH_OBJECT_ID selects per problem, default 0).
 * N = horizon Hp, dt = sample time, batch = number of independent NMPC instances. */
int qspush_solver_create(const qspush_model* const* models, int nmodels, int N, double dt, int batch,
                         int device, const qspush_opts* opts, qspush_solver** out);
void qspush_solver_free(qspush_solver* s);

int qspush_solver_set_opts(qspush_solver* s, const qspush_opts* opts);
int qspush_solver_set_ctrl(qspush_solver* s, const qspush_ctrl* ctrl);

/* stage = -1 addresses all stages of the field at once ([batch][stage][dim]); problems [batch_lo, batch_hi).
 * Fields with a single stage (constr_x0, cost_y_ref_e, cost, residuals) ignore `stage`: the reference calls
 * set('cost_y_ref_e', y, Hp) (NMPC_controller.m:348). */
int qspush_set(qspush_solver* s, qspush_field f, int stage, int batch_lo, int batch_hi, const double* data, qspush_mem mem);
int qspush_get(qspush_solver* s, qspush_field f, int stage, int batch_lo, int batch_hi, double* out, qspush_mem mem);
int qspush_set_int(qspush_solver* s, qspush_field f, int batch_lo, int batch_hi, const int* data, qspush_mem mem);
int qspush_get_int(qspush_solver* s, qspush_field f, int batch_lo, int batch_hi, int* out, qspush_mem mem);

/* NMPC_controller.solve pre-processing on the device (NMPC_controller.m:332, 351-380): x0 wrap,
 * cold start, v_bound clipping of the warm start, Euler rollout. Uses QSPUSH_X0, QSPUSH_COLD. */
int qspush_prepare(qspush_solver* s);
/* ocp_solver.solve() (NMPC_controller.m:389): asynchronous on the solver's stream. */
int qspush_solve(qspush_solver* s);
/* post-processing shift (NMPC_controller.m:397-399): drop stage 0, duplicate the last column. */
int qspush_shift(qspush_solver* s);
/* one forward-Euler plant step x <- x + dt*f(x,u) on caller arrays [batch][4], [batch][2] (helper.m:294,307) */
int qspush_plant_step(qspush_solver* s, double* x, const double* u, qspush_mem mem);
/* Device-resident closed loop = helper.closed_loop_matlab (helper.m:219-313) for the whole batch, without host
 * round trips: per control period i = 1..steps
 *     disturbance (helper.m:221-236, optional) and state noise (:240-242, optional) on the plant state,
 *     constr_x0 <- state, reference window of period idx0 + i - 1 (NMPC_controller.m:307-313, 343-348),
 *     qspush_prepare, qspush_solve, u = get('u', 0) (:403), Euler plant step (helper.m:294, 307), qspush_shift.
 * In RTI mode nothing synchronises with the host until the final copy-out.  Input delays (helper.m:205-212, 244, 250,
 * 290-298; NMPC_controller.m:106-120), counted in control periods: with delay_plant = ceil(plant.time_delay / dt) the plant
 * applies the input of delay_plant periods ago (zeros at first); with delay_comp = controller.delay_buff_comp the controller
 * is handed the state rolled forward through the delay_comp inputs already sent (delay_buffer_sim) and the caller passes the
 * padded reference of set_reference_trajectory with idx0 = 1 + delay_comp (helper.m:248).
 *   traj  [T][6]       reference columns [x_ref(4); u_ref(2)] shared by all problems
 *   offset[batch][6]   added to every column per problem (NULL: none)
 *   x     [batch][4]   plant state, in: initial, out: after `steps` periods
 *   log_x [steps][batch][4] state handed to the controller (after delay_buffer_sim), log_u [steps][batch][2], log_status [steps][batch] (each may be NULL)
 * All arrays live in `mem`. */
typedef struct {
    int    idx0;                 /* reference index of the first period (1-based; helper.m: i = 1)            */
    double noise_sigma[4];       /* helper.m:241 uses 1e-5, 1e-5, 1e-3, 1e-4; all zero = no noise             */
    unsigned long long seed;     /* counter-based generator: same seed, same noise                            */
    int    t_dist;               /* period of the lateral shove (1-based), 0 = none       helper.m:222        */
    double amplitude_dist;       /* helper.m:224                                                               */
    double xwidth;               /* slider_params.xwidth: contact target of the re-projection  helper.m:228   */
    int    delay_plant;          /* delay_buff_plant, control periods (helper.m:211), 0 = none                */
    int    delay_comp;           /* controller.delay_buff_comp, control periods (NMPC_controller.m:108)        */
} qspush_loop_opts;
int qspush_closed_loop(qspush_solver* s, const double* traj, int T, const double* offset, double* x, int steps,
                       const qspush_loop_opts* lo, double* log_x, double* log_u, int* log_status, qspush_mem mem);
/* NMPC_controller.set_reference_trajectory (NMPC_controller.m:425-431) for the batch: the reference columns stay on
 * the device; traj [T][6] = [x_ref(4); u_ref(2)] shared by all problems, offset [batch][6] added per problem (NULL:
 * none); both in `mem`, copied (the caller may free them).  Replaces a previously set trajectory. */
int qspush_set_reference_trajectory(qspush_solver* s, const double* traj, int T, const double* offset, qspush_mem mem);
/* get_y_ref + the per-stage cost_y_ref / cost_y_ref_e calls of NMPC_controller.solve (NMPC_controller.m:307-313,
 * 343-348) for control period idx (1-based): stage k gets column min(idx + k, T), the terminal reference the x part
 * of the last window column.  Asynchronous on the solver's stream; with it a control period moves only x0 (in) and
 * u0 / status (out) between host and device. */
int qspush_set_reference_window(qspush_solver* s, int idx);
/* One control period at the controller boundary = NMPC_controller.solve(x0, index_time) for the whole batch
 * (NMPC_controller.m:329-423) as ONE CUDA graph launch (RTI mode):
 *     constr_x0 <- x0, reference window of period idx (when a trajectory is set, else the references as last set),
 *     x0 wrap / cold start / v_bound clip / Euler rollout (:332, 351-380), linearisation, QP, full step (:389),
 *     u0 = get('u', 0) (:403) and status, optionally the shift of the warm start (:397-399).
 * Kernels per period: k_prepare, k_linearise, k_qp_warp, k_step_out (+ k_shift) — against 9 launches and as many host
 * round trips through qspush_set / _set_reference_window / _prepare / _solve / _get.  Same arithmetic, bit-identical u0.
 *   x0 [batch][4] in, u0 [batch][2] and status [batch] out, all in `mem`; with host memory the call returns after the
 *   results have arrived (one synchronisation), with device memory it is asynchronous on the solver's stream.
 * flags: QSPUSH_STEP_SHIFT          shift the warm start after the solve (what the controller does every period)
 *        QSPUSH_STEP_RESTORE_GUESS  before the period, restore the initial guess saved by qspush_snapshot_guess
 *                                   (benchmarks: every period solves the same problems)                              */
enum { QSPUSH_STEP_SHIFT = 1, QSPUSH_STEP_RESTORE_GUESS = 2 };
int qspush_step(qspush_solver* s, const double* x0, int idx, unsigned flags, double* u0, int* status, qspush_mem mem);
/* save the current input trajectory (QSPUSH_U) on the device for QSPUSH_STEP_RESTORE_GUESS */
int qspush_snapshot_guess(qspush_solver* s);
/* Measured FP64 roofline denominator of this device: DFMA issue-rate microbenchmark (16 independent chains per thread, all
 * SMs, no memory traffic), best of 5 event-timed launches, in TFLOP/s (2 flop per DFMA).  SURVEY.md 7 step 3. */
int qspush_measure_fp64_peak(int device, double* tflops);
/* wait for everything queued on the solver's stream */
int qspush_sync(qspush_solver* s);
/* the solver's cudaStream_t (as void*) so callers can order their own work / events on it */
void* qspush_stream(qspush_solver* s);
int qspush_get_stat(qspush_solver* s, qspush_stat which, double* out);
/* number of kernels launched by this solver since creation (bench bookkeeping) */
long long qspush_launch_count(const qspush_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* QSPUSH_H */
