/* leg_oracle.h -- CPU oracle of the leg kinematics around the force MPC.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/ may load it.
 *
 * Restates, in plain C and with a formulation independent of the product's (homogeneous 4x4 transform
 * chains, rotations from unit quaternions; the product uses closed-form Rodrigues 3x3 products):
 *   include/pinocchio_kinematics.h:30-43,153-157   FK of contact_{L,R}_Link on a fixed-base chain placed at the base pose
 *   include/MPCController.h:106-132                 computeFootPlacement
 *   include/MPCController.h:134-175                 computeSwingFootDesiredPosition (interpolation + IK + cmd.q write)
 *   include/pinocchio_kinematics.h:61-149           inverseKinematics constants (tolerance 1e-3, 10 iterations, DT 0.1,
 *                                                   damping 1e-6), position task
 *   include/MPCController.h:178-180                 computeSupportFootForce stub -> tau = -J' f
 * (paths relative to /root/reference).
 *
 * PARITY UNPINNED: the reference's kinematic model is an external URDF (include/pinocchio_kinematics.h:24) that is
 * not in the repository and Pinocchio is not installed; the reference has no expected values for this block
 * (src/pinocchio_test.cpp only prints).  The link offsets come from include/MPCParam.h:13-38; the joint axes
 * (abad x, hip y, knee y) are this build's assumption and a parameter.  Pinned instead by: the zero pose reproducing
 * static_foot_offset_{left,right} (include/MPCParam.h:64-73) exactly, finite-difference Jacobians, and an independent
 * numpy/scipy witness in the tests.
 */
#ifndef LEG_ORACLE_H
#define LEG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_leg_model {
    double offset[2][5][3];
    double axis[2][3][3];
} orc_leg_model;

typedef struct orc_swing_params {
    float dt, swing_time, stance_time, gait_height;
    double p_rel_max;
    double foot_offset_left[3], foot_offset_right[3];
    double ik_tol, ik_dt, ik_damp;
    int32_t ik_max_iter;
    int32_t ik_mode;   /* 0 position task, 1 the reference's 6-D log6 task (include/pinocchio_kinematics.h:92-132) */
} orc_swing_params;

void orc_leg_defaults(orc_leg_model *m, orc_swing_params *p);
/* world position of the contact point of `leg` and its 3x3 world-aligned Jacobian (row-major; J may be NULL) */
void orc_leg_fk(const orc_leg_model *m, int leg, const double base_pos[3], const double quat_xyzw[4], const double q[3],
                double p[3], double *J);
void orc_foot_placement(const orc_swing_params *p, const double pos[3], const double des_v[3], double remain,
                        int left_leg_state, double fin[3]);
void orc_swing_next(const orc_swing_params *p, const double foot[3], const double fin[3], double remain, double nxt[3]);
int orc_leg_ik(const orc_leg_model *m, const orc_swing_params *p, int leg, const double base_pos[3], const double quat[4],
               const double target[3], double q[3], double *err);
/* The reference's 6-D task as written (include/pinocchio_kinematics.h:92-132): err = log6(oMf^-1 oMdes), oMdes = (I, target);
 * J = -Jlog6(iMd^-1) J_frame; v = -J' (J J' + damp I)^-1 err; q += v DT.  log6 / Jlog6 are computed here by routes that share
 * nothing with the product's closed forms: the rotation vector through a quaternion, the translation part by solving with
 * the SO(3) left-Jacobian matrix, and Jlog6 as the INVERSE of the SE(3) right Jacobian summed as a power series of ad(xi). */
int orc_leg_ik6(const orc_leg_model *m, const orc_swing_params *p, int leg, const double base_pos[3], const double quat[4],
                const double target[3], double q[3], double *err);
/* xi[6] = [v; w] = log6 of the placement with rotation R (row-major 3x3) and translation t */
void orc_se3_log(const double R[9], const double t[3], double xi[6]);
/* Jl (6x6 row-major) = Jlog6 of the same placement */
void orc_se3_jlog(const double R[9], const double t[3], double Jl[36]);
/* whole swing-leg step of MPC::run for one robot; q_cmd[6] in/out; returns the swing leg (0 left, 1 right) */
int orc_swing_step(const orc_leg_model *m, const orc_swing_params *p, const double pos[3], const double quat[4],
                   const double q[6], const double des_v[3], int iter, double q_cmd[6], double feet[6], double next_foot[3],
                   double *ik_err, int *ik_iters);
void orc_grf_to_torque(const orc_leg_model *m, const double quat[4], const double q[6], const double u0[6], double tau[6]);

#ifdef __cplusplus
}
#endif
#endif
