/* mpc_oracle.h -- CPU oracle for the batched convex-MPC hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path
 * (mpc_limx_control_b200/, include/mpc_b200.h) never links or calls it.
 *
 * It is a plain-C restatement of the reference algorithm (paths relative to /root/reference):
 *   src/QPSolver.cpp:21-29    discretizeSystem  (ZOH via expm of the augmented matrix)
 *   src/QPSolver.cpp:31-81    buildQPParams     (A_aug, B_aug, dense Q_bar/R_bar, H, f, bounds)
 *   src/QPSolver.cpp:83-106   solveQP           (qpOASES cold-start active set -> restated as a
 *                                                cold-start dense dual active-set solver)
 *   src/QPSolver.cpp:108-111  updateState
 *   include/mpcQP.h:35-182    TRON1 problem setup / model (literal and intended variants)
 *   include/MPCController.h:61-75 + include/MPCParam.h:44-49   calculateGait (float semantics)
 *
 * PARITY UNPINNED at the third-party boundary: the reference holds no golden vectors and its
 * dependencies (Eigen unsupported MatrixFunctions, qpOASES; both unpinned in CMakeLists.txt:26-30)
 * are absent from the container, so the oracle is pinned instead against an independent
 * numpy/scipy restatement (tests/golden/npref.py) and the fixtures it generated
 * (tests/golden/ npz files), plus KKT self-certification of every QP solution.
 *
 * All matrices are column-major (Eigen default) unless stated otherwise.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INFTY 1.0e20 /* qpOASES::INFTY (src/QPSolver.cpp:72-73) */

/* ---- dense helpers restating the Eigen calls the reference makes ---- */
/* E = exp(A), n x n.  Pade approximant with scaling and squaring (Higham 2005 degree
 * selection 3/5/7/9/13), the algorithm behind Eigen's MatrixBase::exp() (src/QPSolver.cpp:26). */
void orc_expm(int n, const double *A, double *E);
/* P = A^k (k>=0) by binary powering; stands in for MatrixBase::pow(int) (src/QPSolver.cpp:45,75) */
void orc_matpow(int n, const double *A, int k, double *P);

/* ---- QPSolver restatement ---- */
/* src/QPSolver.cpp:21-29.  Ac NXxNX, Bc NXxNU -> Ad, Bd */
void orc_discretize(int NX, int NU, const double *Ac, const double *Bc, double Ts, double *Ad, double *Bd);

/* src/QPSolver.cpp:31-81 (literal, including the dense Q_bar product order).
 * xi_ref is NX x (N+1).  Outputs (any may be NULL):
 *   A_aug NX(N+1) x NX, B_aug NX(N+1) x NU*N, H n x n, f n, A_eq NX*N x n, b_eq NX*N,
 *   lb n, ub n, A_ineq 2*NX*N x n, lbA 2*NX*N, ubA 2*NX*N          (n = NU*N) */
void orc_build_qp_params(int NX, int NU, int N, const double *Ad, const double *Bd,
                         const double *Q, const double *R, const double *P,
                         const double *x_min, const double *x_max, double u_min, double u_max,
                         const double *xi0, const double *xi_ref,
                         double *A_aug, double *B_aug, double *H, double *f,
                         double *A_eq, double *b_eq, double *lb, double *ub,
                         double *A_ineq, double *lbA, double *ubA);

/* src/QPSolver.cpp:108-111:  xi <- Ad xi + Bd u */
void orc_update_state(int NX, int NU, const double *Ad, const double *Bd, double *xi, const double *u);

/* QP  min 1/2 u'Hu + f'u  s.t. lb<=u<=ub, lbA<=A u<=ubA   (src/QPSolver.cpp:83-106 problem form).
 * Cold-start dense dual active-set (Goldfarb-Idnani) standing in for qpOASES::QProblem::init.
 * A is mA x n column-major (may be NULL when mA==0).  Bounds beyond +-ORC_INFTY/2 are infinite.
 * y_bnd[n], y_row[mA] (may be NULL): multipliers, >0 lower bound active, <0 upper bound active,
 *   so that H u + f - y_bnd - A' y_row = 0.
 * returns 0 solved, 1 iteration limit, 2 infeasible / not positive definite. */
int orc_qp_solve(int n, const double *H, const double *f, int mA, const double *A,
                 const double *lbA, const double *ubA, const double *lb, const double *ub,
                 double *u, double *y_bnd, double *y_row, int *iters);

/* KKT residuals of (u, y_bnd, y_row): res[0] stationarity |Hu+f-y_bnd-A'y_row|_inf,
 * res[1] primal infeasibility, res[2] dual sign violation, res[3] complementarity. */
void orc_kkt_residual(int n, const double *H, const double *f, int mA, const double *A,
                      const double *lbA, const double *ubA, const double *lb, const double *ub,
                      const double *u, const double *y_bnd, const double *y_row, double res[4]);

/* ---- gait (include/MPCController.h:61-75, include/MPCParam.h:44-49) ---- */
typedef struct {
    float dt;          /* 0.001f */
    int mpc_step;      /* 5 */
    float swing_time;  /* 0.5f */
    float stance_time; /* 0.5f */
} orc_gait_params;
void orc_gait_defaults(orc_gait_params *g);
/* literal calculateGait: leg states (0 stance, 1 swing), phase, remaining swing time */
void orc_calculate_gait(const orc_gait_params *g, int iter, int *left_leg_state, int *right_leg_state,
                        double *phase, double *remain_swing_time);
/* horizon schedule: contact[k*2+foot] = 1 if foot (0 left, 1 right) is in stance at iter + k*mpc_step.
 * iter < 0 means "standing": both feet in contact for the whole horizon. */
void orc_contact_schedule(const orc_gait_params *g, int iter, int N, uint8_t *contact);

/* ---- TRON1 single-rigid-body problem ---- */
typedef struct {
    double Ts;            /* MPC sampling time */
    double mass;          /* include/mpcQP.h:18 */
    double inertia[9];    /* body inertia, include/mpcQP.h:20-22 (symmetric) */
    double q[13];         /* diag(Q), include/mpcQP.h:54 */
    double r;             /* R = r*I, include/mpcQP.h:55 */
    double p_scale;       /* P = p_scale*Q, include/mpcQP.h:56 */
    double mu;            /* friction coefficient (builder-defined: 0.5) */
    double f_max;         /* max normal force per foot (builder-defined: 2 m g) */
    int ltv;              /* 0: one model at x0 (reference LTI structure); 1: per-step model */
    int per_step_feet;    /* feet given per horizon step ([N][2][3]) instead of [2][3] */
} orc_tron1_params;
void orc_tron1_defaults(orc_tron1_params *p);

/* intended model (north_star): Ac 13x13, Bc 13x6 */
void orc_tron1_model(const orc_tron1_params *p, double yaw, const double pos[3], const double feet[6],
                     double *Ac, double *Bc);
/* reference-literal model (include/mpcQP.h:139-181): Ac 13x13, Bc 13x3, one support foot */
void orc_tron1_model_literal(const orc_tron1_params *p, const double pos[3], const double foot[3],
                             double *Ac, double *Bc);
/* include/mpcQP.h:74-97 reference generator: x_ref 13 x (N+1) */
void orc_tron1_reference(const double x0[13], int N, double Ts, double omega_yaw, double velocity_x,
                         double *x_ref);
/* prediction matrices and cost: A_aug 13(N+1)x13, B_aug 13(N+1)x6N, H 6Nx6N, f 6N (NULL to skip).
 * Dense evaluation in the reference's product order (src/QPSolver.cpp:50-60). */
void orc_tron1_condense(const orc_tron1_params *p, int N, const double *x0, const double *x_ref,
                        const double *feet, double *A_aug, double *B_aug, double *H, double *f);
/* friction pyramid + contact bounds in the generic QP form. A 8N x 6N. */
void orc_tron1_constraints(const orc_tron1_params *p, int N, const uint8_t *contact,
                           double *A, double *lbA, double *ubA, double *lb, double *ub);
/* natural residual |u - P_C(u - (Hu+f))|_inf with the exact projection on the contact-bounded
 * friction pyramid: a solver-independent KKT measure (0 iff u is the minimiser). */
double orc_tron1_natural_residual(const orc_tron1_params *p, int N, const double *H, const double *f,
                                  const uint8_t *contact, const double *u);
/* end to end for one instance: condense (dense) + constraints + cold-start active set.
 * forces[6N]; returns QP status. */
int orc_tron1_solve(const orc_tron1_params *p, int N, const double *x0, const double *x_ref,
                    const double *feet, const uint8_t *contact, double *forces, int *iters);
/* batch over B independent instances with nthreads host threads (one solve per thread at a time).
 * x0[B][13], x_ref[B][N+1][13], feet[B][2][3] or [B][N][2][3], contact[B][N][2],
 * forces[B][N][6], status[B], iters[B] (may be NULL).  returns number of non-zero statuses. */
int orc_tron1_solve_batch(const orc_tron1_params *p, int N, int B, const double *x0, const double *x_ref,
                          const double *feet, const uint8_t *contact, double *forces,
                          int32_t *status, int32_t *iters, int nthreads);

/* closed-loop rollout of ONE instance (BASELINE configs[4]): per control step s
 *   feet = nominal offsets under the base (z = 0), x_ref = orc_tron1_reference(x), contact schedule at
 *   iter0 + s*mpc_step (iter0 < 0: standing), solve (cold-start active set), x <- Ad x + Bd u_0
 *   (src/QPSolver.cpp:108-111) with Ad, Bd = orc_discretize of the model at x.
 * x[13] in/out, u_traj[steps][6] (may be NULL). returns the number of steps whose QP did not solve. */
int orc_tron1_rollout(const orc_tron1_params *p, const orc_gait_params *g, int N, int steps, double *x,
                      double omega_yaw, double velocity_x, int iter0, const double off_l[3],
                      const double off_r[3], double *u_traj);

#ifdef __cplusplus
}
#endif
#endif
