/* mpc_b200.h -- C ABI of the B200-native batched convex-MPC engine (libmpc_b200.so).
 *
 * This is the ONLY route from host code into CUDA.  Plain pointers and sizes, no C++/torch types.
 * The host-side C++ facade (mpc_limx_control_b200/host/) mirrors the reference classes
 * QPSolver / mpcQP / MPC and calls nothing but these entry points.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   include/QPSolver.h:13-37 + src/QPSolver.cpp:21-116   discretize / buildQPParams / solveQP / updateState
 *   include/mpcQP.h:10-11,35-182                         TRON1 problem setup, buildSystemModel
 *   include/MPCController.h:61-75                        calculateGait (horizon contact schedule)
 * plus the batch entry point the reference does not have (BASELINE.json north_star).
 *
 * Conventions
 *   - all floating point data is FP64; matrices are column-major (Eigen default) unless stated
 *   - batch arrays are instance-major and dense:
 *       x0     [B][13]         state [roll,pitch,yaw, px,py,pz, wx,wy,wz, vx,vy,vz, g]  (include/mpcQP.h:66-71)
 *       x_ref  [B][N+1][13]    == column-major 13 x (N+1) per instance (include/mpcQP.h:74)
 *       feet   [B][2][3]       world foot positions (left,right); [B][N][2][3] when per_step_feet
 *       contact[B][N][2]       uint8, 1 = stance;  or  iter[B] int32 -> schedule from the gait clock
 *       forces [B][N][6]       == column-major 6 x N U_opt per instance (src/QPSolver.cpp:104)
 *       status [B] int32       0 solved (KKT certified), 1 iteration limit, 2 failed (non-finite / not PD)
 *       iters  [B] int32       active-face solves + ADMM iterations spent
 *   - every function returns 0 on success or a negative MPC_B200_E* code; nothing prints
 *   - an engine is bound to one CUDA device and is single-caller (not re-entrant); its calls must not
 *     overlap on the device (issue them in one stream, or synchronise between streams)
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream)
 *   - there is no CPU fallback: without a usable CUDA device create() fails with MPC_B200_ENODEV
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_B200_OK 0
#define MPC_B200_EINVAL (-1)   /* bad argument (NULL, size, unsupported horizon, misaligned pointer) */
#define MPC_B200_ENODEV (-2)   /* no CUDA device / wrong architecture */
#define MPC_B200_ECUDA (-3)    /* CUDA runtime error (see mpc_b200_last_error) */
#define MPC_B200_ENOMEM (-4)
#define MPC_B200_ECAPACITY (-5) /* batch larger than the engine's max_batch */

#define MPC_B200_INFTY 1.0e20  /* stands in for qpOASES::INFTY (src/QPSolver.cpp:72-73) */

typedef struct mpc_b200_engine mpc_b200_engine;

typedef struct mpc_b200_tron1_params {
    double Ts;            /* MPC step; default dtMPC = 0.005 (include/MPCParam.h:47) */
    double mass;          /* 9.585 (include/mpcQP.h:18) */
    double inertia[9];    /* body inertia (include/mpcQP.h:20-22), symmetric */
    double q[13];         /* diag(Q) (include/mpcQP.h:54) */
    double r;             /* R = r I (include/mpcQP.h:55) */
    double p_scale;       /* P = p_scale Q (include/mpcQP.h:56) */
    double mu;            /* friction pyramid coefficient (not defined by the reference; 0.5) */
    double f_max;         /* normal force cap per foot (not defined by the reference; 2 m g) */
    int32_t ltv;          /* 0: one model at x0 (reference LTI structure), 1: per-step model */
    int32_t per_step_feet;/* feet given per horizon step */
    /* gait clock (include/MPCParam.h:44-49) */
    float gait_dt;        /* 0.001f */
    int32_t gait_mpc_step;/* 5 */
    float gait_swing_time;  /* 0.5f */
    float gait_stance_time; /* 0.5f */
    /* solver */
    int32_t max_newton;   /* active-face iterations before the ADMM fallback (default 12) */
    int32_t max_admm;     /* ADMM iteration cap (default 2000) */
    double tol;           /* natural-residual tolerance relative to max(1,|u|_inf) (default 1e-9) */
    /* nominal base->foot offsets (include/MPCParam.h:64-73), used by the closed-loop rollout */
    double foot_offset_left[3];
    double foot_offset_right[3];
} mpc_b200_tron1_params;

/* library / device */
int mpc_b200_version(void);
const char *mpc_b200_strerror(int code);
int mpc_b200_device_count(void);
/* measured FP64 FMA peak of `device` (register-resident DFMA chains), TFLOP/s */
int mpc_b200_measure_fp64_peak(int device, double *tflops);

/* engine lifetime.  horizon N must be one of the compiled horizons (10, 20, 50; closed-loop rollout: 10, 20).
 * Device memory per engine: the batch buffers of the host entry points (about 2 KB x max_batch at N = 10), two overflow lists
 * (8 B x 7 x max_batch) and, for horizon 50, 40 MB of gain / force slabs for the Riccati class (8 x SM count slabs of 34 KB, handed out
 * through a ring of free slabs, so any number of concurrent launches shares them) plus the factor slabs of the dense class. */
int mpc_b200_tron1_default_params(mpc_b200_tron1_params *p);
int mpc_b200_create(const mpc_b200_tron1_params *p, int horizon, int max_batch, int device,
                    mpc_b200_engine **out);
int mpc_b200_destroy(mpc_b200_engine *e);
const char *mpc_b200_last_error(const mpc_b200_engine *e);
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
int64_t mpc_b200_launch_count(const mpc_b200_engine *e);

/* MPC::calculateGait over the horizon (include/MPCController.h:61-75): contact[b][k][foot] for
 * iter[b] + k*mpc_step; iter < 0 = standing (both feet in contact).  Device pointers. */
int mpc_b200_contact_schedule_device(mpc_b200_engine *e, int B, const int32_t *d_iter,
                                     uint8_t *d_contact, void *stream);

/* The hot path on device-resident data: linearise -> discretise -> condense -> QP solve.
 * Exactly one of d_contact / d_iter may be NULL (d_iter: schedule evaluated in-kernel).
 * d_status / d_iters may be NULL.
 * Stream contract: an engine is single-caller.  All calls on ONE engine must be stream-ordered and must not
 * overlap (same stream, or streams ordered by events): the capacity-overflow list, its counters and (horizon 50)
 * the global factor slabs are per-engine scratch.  Use one engine per concurrent stream.
 * Small batches: with B <= the device's SM count, double-support instances of horizon 10 (a standing robot) are
 * solved by a latency-oriented kernel class (8 warps per instance).  Their forces agree with those of the same
 * instance inside a larger batch to rounding, not bit for bit; all other results are independent of B. */
int mpc_b200_tron1_solve_device(mpc_b200_engine *e, int B, const double *d_x0, const double *d_x_ref,
                                const double *d_feet, const uint8_t *d_contact, const int32_t *d_iter,
                                double *d_forces, int32_t *d_status, int32_t *d_iters, void *stream);

/* Pipelined variant for streams of INDEPENDENT batches (a server working through a queue, a benchmark loop): the solve
 * is ordered after everything already queued on `stream` at the time of the call (so its inputs are ready), but it runs
 * on one of six engine-owned streams, used in rotation, and is NOT ordered against the previous pipelined calls: the
 * CTAs of batch k+1 fill the SM slots that the last, multi-iteration instances of batch k leave idle (at B = 4096 a third
 * of a batch's time is that tail).  Results become visible to `stream` only after mpc_b200_join(e, stream); a batch that
 * consumes the previous batch's forces must use mpc_b200_tron1_solve_device (or join in between).  Output buffers of
 * calls that may overlap (six consecutive calls) must be distinct if all of them are wanted. */
int mpc_b200_tron1_solve_device_pipelined(mpc_b200_engine *e, int B, const double *d_x0, const double *d_x_ref,
                                          const double *d_feet, const uint8_t *d_contact, const int32_t *d_iter,
                                          double *d_forces, int32_t *d_status, int32_t *d_iters, void *stream);
/* Makes `stream` wait for every pipelined solve issued so far (no host synchronisation). */
int mpc_b200_join(mpc_b200_engine *e, void *stream);

/* Same with HOST buffers.  Two data paths:
 *   zero-copy  every buffer is pinned / registered (device-addressable under UVA): the solve kernel reads its
 *              inputs from, and writes its results to, host memory directly (TMA bulk copies over PCIe per CTA,
 *              overlapped with the solves of the other resident CTAs); no cudaMemcpy, one stream sync.
 *   staged     pageable buffers: H2D copies, solve, D2H copies; large transfers are split into chunks
 *              pipelined over several streams, batches <= 64 take a packed single-copy path.
 * MPC_B200_HOST_AUTO (default) picks zero-copy whenever all buffers qualify. */
#define MPC_B200_HOST_AUTO 0
#define MPC_B200_HOST_STAGED 1
#define MPC_B200_HOST_ZEROCOPY 2   /* fail with MPC_B200_EINVAL instead of falling back to staged copies */
int mpc_b200_set_host_mode(mpc_b200_engine *e, int mode);
/* data path the last host-buffer call took: 1 zero-copy, 0 staged */
int mpc_b200_last_host_path(const mpc_b200_engine *e);
/* Page-lock an existing host allocation (malloc, std::vector, Eigen, numpy ...) so that calls using it take the
 * zero-copy path; for callers that do not link the CUDA runtime themselves.  Pin once at start-up (it costs about a
 * millisecond per few MB), unpin before freeing the memory.  Pinning an already pinned range is not an error. */
int mpc_b200_pin_host_buffer(void *ptr, size_t bytes);
int mpc_b200_unpin_host_buffer(void *ptr);
int mpc_b200_tron1_solve_host(mpc_b200_engine *e, int B, const double *x0, const double *x_ref,
                              const double *feet, const uint8_t *contact, const int32_t *iter,
                              double *forces, int32_t *status, int32_t *iters);

/* Asynchronous host-buffer entry for a stream of independent batches (SURVEY.md section 8b "async enqueue + wait pair"):
 * returns as soon as the work is queued on one of the engine's six lanes; mpc_b200_wait blocks until every queued batch
 * is solved and its results are in the caller's arrays.  Every buffer must be pinned / registered.  Inputs are read
 * zero-copy by the kernel; results go through per-lane device buffers and the copy engine, so that one batch's
 * copy-back overlaps the next batch's reads.  Buffers of a batch must not be touched between the call and the wait. */
int mpc_b200_tron1_solve_host_async(mpc_b200_engine *e, int B, const double *x0, const double *x_ref, const double *feet,
                                    const uint8_t *contact, const int32_t *iter, double *forces, int32_t *status,
                                    int32_t *iters);
/* controller-shaped counterpart (arguments as mpc_b200_tron1_control_host): 228 bytes per instance cross PCIe, the first-step
 * forces are written by the kernel straight into the pinned u0 array */
int mpc_b200_tron1_control_host_async(mpc_b200_engine *e, int B, const double *x0, const double *omega_yaw,
                                      const double *velocity_x, const double *feet, const uint8_t *contact,
                                      const int32_t *iter, double *u0, int32_t *status, int32_t *iters);
int mpc_b200_wait(mpc_b200_engine *e);

/* Single-process multi-GPU batch entry (the reference has no counterpart: it solves one robot per
 * call, include/mpcQP.h:63,113).  engines[0..G) are engines of the SAME horizon and parameters on DIFFERENT devices, each
 * with max_batch >= its share.  The batch is block-partitioned by instance -- GPU g owns [g*B/G + min(g, B%G), ...), the
 * first B%G GPUs one instance more -- and every GPU runs mpc_b200_tron1_solve_host on its slice of the caller's arrays
 * from its own host thread and stream; results land in the caller's arrays (one gather-free pinned result array when the
 * buffers are pinned: every GPU writes its rows over PCIe itself).  No collective, no peer traffic.
 * Returns the first non-zero status of any GPU (all GPUs are always waited for). */
int mpc_b200_tron1_solve_host_multi(mpc_b200_engine *const *engines, int G, int B, const double *x0, const double *x_ref,
                                    const double *feet, const uint8_t *contact, const int32_t *iter,
                                    double *forces, int32_t *status, int32_t *iters);

/* Controller-shaped host entry, the closest analogue of the reference's mpcQP constructor
 * (include/mpcQP.h:35-119): state, commanded yaw rate / forward velocity (the reference generator of
 * include/mpcQP.h:74-97 runs on the device), feet and gait in; u = U_opt.col(0) (include/mpcQP.h:118)
 * out: u0 [B][6].  Moves 172 B in and 56 B out per instance instead of 1,300 / 488. */
int mpc_b200_tron1_control_host(mpc_b200_engine *e, int B, const double *x0, const double *omega_yaw,
                                const double *velocity_x, const double *feet, const uint8_t *contact,
                                const int32_t *iter, double *u0, int32_t *status, int32_t *iters);

/* Parity dump of the condensed problem (any output may be NULL), device pointers:
 *   H [B][6N x 6N], f [B][6N], A_aug [B][13(N+1) x 13], B_aug [B][13(N+1) x 6N]  (column-major)
 * H and f do not depend on the contact schedule (src/QPSolver.cpp:58-60). */
int mpc_b200_tron1_condense_device(mpc_b200_engine *e, int B, const double *d_x0, const double *d_x_ref,
                                   const double *d_feet, double *d_H, double *d_f, double *d_A_aug,
                                   double *d_B_aug, void *stream);

/* mpcQP's reference generator (include/mpcQP.h:74-97) for a batch: x_ref[b] from x0[b] and the
 * per-instance commanded yaw rate / forward velocity.  Device pointers. */
int mpc_b200_tron1_reference_device(mpc_b200_engine *e, int B, const double *d_x0, const double *d_omega_yaw,
                                    const double *d_velocity_x, double *d_x_ref, void *stream);

/* Closed-loop rollout (BASELINE configs[4]): `steps` control steps of
 *   reference(x) -> contact schedule(iter0 + s*mpc_step) -> linearise -> condense -> solve -> x <- Ad x + Bd u_0
 * with the state resident on the device, feet at their nominal offsets under the base, and the
 * previous step's optimal face (shifted by one horizon step) as warm start.  The plant is the
 * model (QPSolver::updateState, src/QPSolver.cpp:108-111).  Device pointers:
 *   d_x [B][13] in/out (initial -> final state), d_omega_yaw/d_velocity_x [B], d_iter0 [B] (<0 standing),
 *   d_u_traj [B][steps][6] first-step forces (may be NULL), d_uncertified [B] number of steps whose
 *   solve was not certified (may be NULL), d_iters [B] total solver iterations (may be NULL). */
int mpc_b200_tron1_rollout_device(mpc_b200_engine *e, int B, int steps, double *d_x, const double *d_omega_yaw,
                                  const double *d_velocity_x, const int32_t *d_iter0, double *d_u_traj,
                                  int32_t *d_uncertified, int32_t *d_iters, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Leg kinematics around the force MPC: the rest of MPC::run (include/MPCController.h:183-196).
 *   include/pinocchio_kinematics.h:30-43,153-157  forwardKinematics / getLinkPosition / setBaseLinkPose  -> leg_fk
 *   include/MPCController.h:61-75,106-175         calculateGait, computeFootPlacement, computeSwingFootDesiredPosition
 *   include/pinocchio_kinematics.h:61-149         inverseKinematics                                       -> swing_step
 *   include/MPCController.h:178-180               computeSupportFootForce (empty stub): tau = -J' f       -> grf_to_torque
 * Layouts (instance-major, FP64): base_pos [B][3]; base_quat [B][4] stored [x,y,z,w]
 * (include/state_estimator_fake.h:22); q / q_cmd / tau [B][6] = left abad,hip,knee, right abad,hip,knee
 * (RobotCmd order, include/MPCController.h:164-174); feet [B][2][3] world (the `feet` input of the solve);
 * jac [B][2][3][3] row-major world-aligned d(foot)/d(q_leg); des_vel [B][3]; u0 [B][6] first-step forces.
 * The joint axes are NOT in the reference (its URDF lives in an external repository): they are model
 * parameters, default abad = x, hip = knee = y; the link offsets are include/MPCParam.h:13-38. */
typedef struct mpc_b200_leg_model {
    double offset[2][5][3];  /* leg (0 left, 1 right) x {base->abad, abad->hip, hip->knee, knee->foot, foot->contact} */
    double axis[2][3][3];    /* leg x joint: unit axis in the parent frame */
} mpc_b200_leg_model;

typedef struct mpc_b200_swing_params {
    float dt, swing_time, stance_time, gait_height;   /* include/MPCParam.h:44-51 */
    double p_rel_max;                                  /* 0.3, include/MPCController.h:111 */
    double foot_offset_left[3], foot_offset_right[3];  /* include/MPCParam.h:64-73 */
    double ik_tol, ik_dt, ik_damp;                     /* 1e-3, 1e-1, 1e-6: include/pinocchio_kinematics.h:61,76-77 */
    int32_t ik_max_iter;                               /* 10 */
    int32_t ik_mode;                                   /* 0 (default): position task; 1: the reference's 6-D log6 task as written,
                                                          include/pinocchio_kinematics.h:92-132 */
} mpc_b200_swing_params;

int mpc_b200_leg_default_model(mpc_b200_leg_model *m);
int mpc_b200_swing_default_params(mpc_b200_swing_params *p);

/* World positions of contact_L_Link / contact_R_Link (and, if d_jac != NULL, the leg Jacobians). Device pointers,
 * launched on `stream` on the current device. */
int mpc_b200_leg_fk_device(const mpc_b200_leg_model *m, int B, const double *d_base_pos, const double *d_base_quat,
                           const double *d_q, double *d_feet, double *d_jac, void *stream);

/* Swing-leg step: gait state of iter[b] -> landing point from the desired velocity -> next swing-foot position
 * (linear xy, sine height) -> damped least-squares IK from the current joint angles -> the swing leg's three
 * entries of q_cmd (the stance leg's entries are left untouched, as in the reference).  Optional outputs (may be
 * NULL): feet [B][2][3] (FK at the current q: feed it to the solve), next_foot [B][3], swing_leg [B] (0 left,
 * 1 right), ik_err [B] (last position error norm), ik_iters [B]. */
int mpc_b200_swing_step_device(const mpc_b200_leg_model *m, const mpc_b200_swing_params *p, int B,
                               const double *d_base_pos, const double *d_base_quat, const double *d_q,
                               const double *d_des_vel, const int32_t *d_iter, double *d_q_cmd, double *d_feet,
                               double *d_next_foot, int32_t *d_swing_leg, double *d_ik_err, int32_t *d_ik_iters,
                               void *stream);

/* PinocchioKinematics::inverseKinematics (include/pinocchio_kinematics.h:61-149) for a batch: joint angles that bring
 * contact_L_Link (leg[b] = 0) / contact_R_Link (1) to target [B][3] (world), starting from q_init [B][6]; q_out [B][6]
 * (the other leg's entries are copied through).  p->ik_mode selects the task; tolerance / step / damping / iteration cap
 * are the reference's constants in *p.  ik_err [B] = norm of the last error the loop computed (3-D or 6-D), ik_iters [B]
 * (both may be NULL). */
int mpc_b200_leg_ik_device(const mpc_b200_leg_model *m, const mpc_b200_swing_params *p, int B, const double *d_base_pos,
                           const double *d_base_quat, const int32_t *d_leg, const double *d_target, const double *d_q_init,
                           double *d_q_out, double *d_ik_err, int32_t *d_ik_iters, void *stream);

/* tau = -J(q)' f for both legs from the first-step forces u0 (a swing foot has f = 0, hence tau = 0). */
int mpc_b200_grf_to_torque_device(const mpc_b200_leg_model *m, int B, const double *d_base_quat, const double *d_q,
                                  const double *d_u0, double *d_tau, void *stream);

/* The same three calls with HOST buffers (copies in, kernel, copies out, synchronise) on `device`; used by the
 * single-robot C++ facade.  One caller per device. */
int mpc_b200_leg_fk_host(int device, const mpc_b200_leg_model *m, int B, const double *base_pos, const double *base_quat,
                         const double *q, double *feet, double *jac);
int mpc_b200_swing_step_host(int device, const mpc_b200_leg_model *m, const mpc_b200_swing_params *p, int B,
                             const double *base_pos, const double *base_quat, const double *q, const double *des_vel,
                             const int32_t *iter, double *q_cmd, double *feet, double *next_foot, int32_t *swing_leg,
                             double *ik_err, int32_t *ik_iters);
int mpc_b200_leg_ik_host(int device, const mpc_b200_leg_model *m, const mpc_b200_swing_params *p, int B,
                         const double *base_pos, const double *base_quat, const int32_t *leg, const double *target,
                         const double *q_init, double *q_out, double *ik_err, int32_t *ik_iters);
int mpc_b200_grf_to_torque_host(int device, const mpc_b200_leg_model *m, int B, const double *base_quat, const double *q,
                                const double *u0, double *tau);

/* ------------------------------------------------------------------------------------------------
 * Base-state Kalman filter for a batch of robots: the reference's unbuilt, OCS2-derived
 * stateEstimator::update (include/stateEstimator.h:217-337).  State xhat [B][12] = base position, base velocity,
 * left and right foot position (world); covariance P [B][12 x 12] row-major (symmetric); one call = one update with
 * time step dt from IMU orientation quat [B][4] ([x,y,z,w]), body-frame angular velocity gyro [B][3] and
 * acceleration accel [B][3], joint angles/velocities q, dq [B][6] and contact flags contact [B][2] (uint8, 1 = in
 * contact).  odom [B][13] (may be NULL) = RobotOdomState of include/state_estimator_fake.h:19-25 as filled at
 * include/stateEstimator.h:319-333: pos 3, quat 4, v_pos 3 (body frame), v_ori 3.  Noise constants :124-130. */
typedef struct mpc_b200_kf_params {
    double foot_radius;                   /* 0.02 */
    double imu_process_noise_position;    /* 0.02 */
    double imu_process_noise_velocity;    /* 0.02 */
    double foot_process_noise_position;   /* 0.002 */
    double foot_sensor_noise_position;    /* 0.005 */
    double foot_sensor_noise_velocity;    /* 0.1 */
    double foot_height_sensor_noise;      /* 0.01 */
    double high_suspect_number;           /* 100: noise inflation of a foot that is not in contact (:262) */
    int32_t accel_transpose;              /* 1: accel = R' a_local + g as written at :281; 0: R a_local + g */
} mpc_b200_kf_params;
int mpc_b200_kf_default_params(mpc_b200_kf_params *p);
/* xhat = 0, P = p0 I (the reference starts from 100 I, include/stateEstimator.h:206-207) */
int mpc_b200_kf_reset_device(int B, double p0, double *d_xhat, double *d_P, void *stream);
int mpc_b200_kf_update_device(const mpc_b200_kf_params *p, const mpc_b200_leg_model *m, int B, double dt,
                              const double *d_quat, const double *d_gyro_local, const double *d_accel_local,
                              const double *d_q, const double *d_dq, const uint8_t *d_contact, double *d_xhat,
                              double *d_P, double *d_odom, void *stream);
int mpc_b200_kf_update_host(int device, const mpc_b200_kf_params *p, const mpc_b200_leg_model *m, int B, double dt,
                            const double *quat, const double *gyro_local, const double *accel_local, const double *q,
                            const double *dq, const uint8_t *contact, double *xhat, double *P, double *odom);

/* ------------------------------------------------------------------------------------------------
 * Generic condensed-MPC path: the reference class QPSolver (include/QPSolver.h:13-37), any NX/NU/N.
 * HOST pointers, column-major (Eigen layout), B independent instances per call (B = 1 for the
 * facade).  Copies in, runs one CTA per instance on the device, copies out, synchronises. */
typedef struct mpc_b200_lti mpc_b200_lti;
int mpc_b200_lti_create(int device, mpc_b200_lti **out);
int mpc_b200_lti_destroy(mpc_b200_lti *c);
const char *mpc_b200_lti_last_error(const mpc_b200_lti *c);
int64_t mpc_b200_lti_launch_count(const mpc_b200_lti *c);

/* QPSolver::discretizeSystem (src/QPSolver.cpp:21-29): Ad, Bd = blocks of exp([[Ac,Bc],[0,0]] Ts).
 * Ac [B][NX x NX], Bc [B][NX x NU]. */
int mpc_b200_lti_discretize(mpc_b200_lti *c, int B, int NX, int NU, double Ts, const double *Ac,
                            const double *Bc, double *Ad, double *Bd);

/* QPSolver::buildQPParams (src/QPSolver.cpp:31-81).  One system (Ad, Bd, Q, R, P, x_min, x_max,
 * u_min, u_max), B initial states xi0 [B][NX] and references xi_ref [B][NX x (N+1)].
 * Outputs per instance (any may be NULL): H [n x n], f [n], A_eq [NX N x n], b_eq [NX N], lb/ub [n],
 * A_ineq [2 NX N x n], lbA/ubA [2 NX N], A_aug [NX(N+1) x NX], B_aug [NX(N+1) x n], n = NU N.
 * A_eq/b_eq reproduce the reference block as written; it is spurious and the
 * facade does not pass it on to the solver. */
int mpc_b200_lti_build_qp(mpc_b200_lti *c, int B, int NX, int NU, int N, const double *Ad,
                          const double *Bd, const double *Q, const double *R, const double *P,
                          const double *x_min, const double *x_max, double u_min, double u_max,
                          const double *xi0, const double *xi_ref, double *H, double *f, double *A_eq,
                          double *b_eq, double *lb, double *ub, double *A_ineq, double *lbA, double *ubA,
                          double *A_aug, double *B_aug);

/* QPSolver::solveQP (src/QPSolver.cpp:83-106): min 1/2 u'Hu + f'u, lb <= u <= ub, lbA <= A u <= ubA.
 * H [B][n x n] symmetric positive definite, A [B][m x n] (m may be 0), bounds beyond
 * +-MPC_B200_INFTY/2 are infinite, all-zero rows of A are ignored.  status/iters may be NULL.
 * Unlike the reference (which ignores qpOASES' return value) the status is reported. */
int mpc_b200_qp_solve_dense(mpc_b200_lti *c, int B, int n, int m, const double *H, const double *f,
                            const double *A, const double *lb, const double *ub, const double *lbA,
                            const double *ubA, double *U, int32_t *status, int32_t *iters);

/* QPSolver::updateState (src/QPSolver.cpp:108-111): xi <- Ad xi + Bd u, xi [B][NX] in place. */
int mpc_b200_lti_update_state(mpc_b200_lti *c, int B, int NX, int NU, const double *Ad,
                              const double *Bd, double *xi, const double *u);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
