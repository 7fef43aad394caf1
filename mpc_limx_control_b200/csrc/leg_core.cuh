// leg_core.cuh -- per-robot leg kinematics around the force MPC: the "next" rows of SURVEY.md section 8(f).
//
// What it replaces in the reference (paths relative to /root/reference):
//   include/pinocchio_kinematics.h:30-43,153-157  setBaseLinkPose / forwardKinematics / getLinkPosition
//                                                 -> leg_fk(): world position of contact_{L,R}_Link and the
//                                                    world-aligned 3x3 position Jacobian of each leg
//   include/MPCController.h:106-132               computeFootPlacement        -> foot_placement()
//   include/MPCController.h:134-158               swing trajectory (linear xy + sine z) -> swing_next_position()
//   include/pinocchio_kinematics.h:61-149         inverseKinematics (damped least squares, DT 0.1, damping
//                                                 1e-6, tolerance 1e-3, <= 10 iterations) -> leg_ik()
//   include/MPCController.h:178-180               computeSupportFootForce (empty stub): tau = -J' f -> grf_to_torque()
//
// Kinematic model: the reference loads PF_TRON1A/urdf/robot.urdf, which is NOT in the repository
// (include/pinocchio_kinematics.h:24).  The link offsets are the reference's own (include/MPCParam.h:13-38);
// the joint axes are not stated anywhere in the reference and are therefore PARAMETERS of LegModel
// (default: abad about x, hip and knee about y, all link frames axis-aligned at q = 0 -- the zero pose then
// reproduces static_foot_offset_{left,right} of include/MPCParam.h:64-73 exactly).  Parity for this block is
// unpinned (no URDF, no Pinocchio); the oracle (oracle/leg_oracle.c) defines it.
//
// The reference's IK is 6-D (log6 to an identity target orientation) on the 6-joint fixed-base model; a
// 3-DoF point-foot leg cannot track orientation.  Two tasks are provided (SwingParams::ik_mode):
//   0  position-only damped least squares with the reference's constants (what the author evidently wants from a
//      point foot; the default of the swing-leg step)
//   1  REFERENCE-LITERAL: the 6-D task exactly as written at include/pinocchio_kinematics.h:61-149 -- err =
//      log6(oMf^-1 oMdes) with oMdes = (I, target), J = -Jlog6(iMd^-1) J_frame(LOCAL_WORLD_ALIGNED),
//      v = -J' (J J' + damp I)^-1 err, q += v DT -- with log6 / Jlog6 restated from Pinocchio's published formulas
//      (Pinocchio itself is an absent, unpinned dependency).  The frame placement is recomputed every iteration
//      (the reference reads data.oMf without updateFramePlacements inside its loop, :104-106; which placement that
//      yields depends on the Pinocchio version and is not replicated).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LEG_HD __host__ __device__ __forceinline__
#else
#define LEG_HD inline
#endif

namespace mpcb200 {

struct LegModel {
    double offset[2][5][3];   // leg (0 left, 1 right) x {base->abad, abad->hip, hip->knee, knee->foot, foot->contact}
    double axis[2][3][3];     // leg x joint (abad, hip, knee): unit axis in the parent link frame
};

struct SwingParams {          // include/MPCParam.h:44-51,64-73 (float members kept as float)
    float dt, swing_time, stance_time, gait_height;
    double p_rel_max;         // 0.3, include/MPCController.h:111
    double foot_off_l[3], foot_off_r[3];
    // IK constants, include/pinocchio_kinematics.h:61,74-77
    double ik_tol, ik_dt, ik_damp;
    int ik_max_iter;
    int ik_mode;              // 0 position task, 1 reference-literal 6-D task
};

// R = I + sin(q) [a]x + (1 - cos(q)) [a]x^2   (row-major 3x3)
LEG_HD void axis_angle(const double a[3], double q, double R[9]) {
    double s, c;
    sincos(q, &s, &c);
    const double v = 1.0 - c, x = a[0], y = a[1], z = a[2];
    R[0] = c + v * x * x;     R[1] = v * x * y - s * z; R[2] = v * x * z + s * y;
    R[3] = v * x * y + s * z; R[4] = c + v * y * y;     R[5] = v * y * z - s * x;
    R[6] = v * x * z - s * y; R[7] = v * y * z + s * x; R[8] = c + v * z * z;
}
LEG_HD void mat3_vec(const double R[9], const double v[3], double o[3]) {
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
LEG_HD void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
LEG_HD void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// unit quaternion stored [x, y, z, w] (include/state_estimator_fake.h:22) -> rotation matrix; the quaternion is
// normalised first (Eigen's toRotationMatrix assumes a unit quaternion, include/pinocchio_kinematics.h:154)
LEG_HD void quat_to_rot(const double q[4], double R[9]) {
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double in = n > 0.0 ? 1.0 / n : 0.0;
    const double x = q[0] * in, y = q[1] * in, z = q[2] * in, w = n > 0.0 ? q[3] * in : 1.0;
    R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - z * w);       R[2] = 2.0 * (x * z + y * w);
    R[3] = 2.0 * (x * y + z * w);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - x * w);
    R[6] = 2.0 * (x * z - y * w);       R[7] = 2.0 * (y * z + x * w);       R[8] = 1.0 - 2.0 * (x * x + y * y);
}

// Forward kinematics of one leg in the BASE frame: contact point p and, if J != nullptr, the 3x3 position
// Jacobian dp/dq (row-major, columns = abad, hip, knee).
LEG_HD void leg_fk_base(const LegModel& M, int leg, const double q[3], double p[3], double* J) {
    const double (*o)[3] = M.offset[leg];
    double R0[9], R1[9], R2[9], R01[9], R012[9];
    axis_angle(M.axis[leg][0], q[0], R0);
    axis_angle(M.axis[leg][1], q[1], R1);
    axis_angle(M.axis[leg][2], q[2], R2);
    mat3_mul(R0, R1, R01);
    mat3_mul(R01, R2, R012);
    double tip[3] = {o[3][0] + o[4][0], o[3][1] + o[4][1], o[3][2] + o[4][2]};   // knee -> contact, in the shank frame
    double t2[3], t1[3], t0[3];
    mat3_vec(R012, tip, t2);        // knee joint -> contact, base frame
    mat3_vec(R01, o[2], t1);        // hip joint -> knee joint
    mat3_vec(R0, o[1], t0);         // abad joint -> hip joint
    const double c0[3] = {o[0][0], o[0][1], o[0][2]};
    const double c1[3] = {c0[0] + t0[0], c0[1] + t0[1], c0[2] + t0[2]};
    const double c2[3] = {c1[0] + t1[0], c1[1] + t1[1], c1[2] + t1[2]};
    for (int i = 0; i < 3; ++i) p[i] = c2[i] + t2[i];
    if (J) {
        double z1[3], z2[3], d[3], col[3];
        const double* z0 = M.axis[leg][0];
        mat3_vec(R0, M.axis[leg][1], z1);
        mat3_vec(R01, M.axis[leg][2], z2);
        for (int i = 0; i < 3; ++i) d[i] = p[i] - c0[i];
        cross3(z0, d, col); J[0] = col[0]; J[3] = col[1]; J[6] = col[2];
        for (int i = 0; i < 3; ++i) d[i] = p[i] - c1[i];
        cross3(z1, d, col); J[1] = col[0]; J[4] = col[1]; J[7] = col[2];
        for (int i = 0; i < 3; ++i) d[i] = p[i] - c2[i];
        cross3(z2, d, col); J[2] = col[0]; J[5] = col[1]; J[8] = col[2];
    }
}

// World position of the contact point and the world-aligned Jacobian (LOCAL_WORLD_ALIGNED position rows,
// include/pinocchio_kinematics.h:113) of one leg.  Rb = rotation of the base (row-major).
LEG_HD void leg_fk_world(const LegModel& M, int leg, const double base_pos[3], const double Rb[9], const double q[3],
                         double pw[3], double* Jw) {
    double pb[3], Jb[9];
    leg_fk_base(M, leg, q, pb, Jw ? Jb : nullptr);
    double r[3];
    mat3_vec(Rb, pb, r);
    for (int i = 0; i < 3; ++i) pw[i] = base_pos[i] + r[i];
    if (Jw) mat3_mul(Rb, Jb, Jw);
}

// MPC::calculateGait (include/MPCController.h:61-75) including phase and remainSwingTime, float semantics
// as in tron1_core.cuh gait_contact (int * float product rounded to float; float add for the cycle).
LEG_HD void gait_state(const SwingParams& P, int iter, int& left_leg_state, int& right_leg_state, double& phase, double& remain) {
#if defined(__CUDA_ARCH__)
    const float ct = __fmul_rn((float)iter, P.dt);
    const float cy = __fadd_rn(P.swing_time, P.stance_time);
#else
    volatile float ctv = (float)iter * P.dt;
    volatile float cyv = P.swing_time + P.stance_time;
    const float ct = ctv, cy = cyv;
#endif
    const double x = (double)ct, y = (double)cy;
    phase = fmod(x, y);
    if (phase < (double)P.swing_time) { left_leg_state = 1; right_leg_state = 0; remain = (double)P.swing_time - phase; }
    else { left_leg_state = 0; right_leg_state = 1; remain = y - phase; }
}

// MPC::computeFootPlacement (include/MPCController.h:106-132): landing point of the swing foot (world xy; the
// reference never assigns z -- it is overwritten by the sine profile afterwards -- so z = 0 here).
LEG_HD void foot_placement(const SwingParams& P, const double pos[3], const double des_v[3], double remain,
                           int left_leg_state, double fin[3]) {
    double px = pos[0] + des_v[0] * remain, py = pos[1] + des_v[1] * remain;
    double pfx = des_v[0] * 0.5 * (double)P.stance_time, pfy = des_v[1] * 0.5 * (double)P.stance_time;
    pfx = fmin(fmax(pfx, -P.p_rel_max), P.p_rel_max);
    pfy = fmin(fmax(pfy, -P.p_rel_max), P.p_rel_max);
    px += pfx; py += pfy;
    const double* off = (left_leg_state == 1) ? P.foot_off_l : P.foot_off_r;
    fin[0] = px + off[0];
    fin[1] = py + off[1];
    fin[2] = 0.0;
}

// next swing-foot position (include/MPCController.h:155-158): linear interpolation towards the landing point
// by the elapsed fraction of the swing, height = gait_height * sin(pi * fraction)
LEG_HD void swing_next_position(const SwingParams& P, const double foot[3], const double fin[3], double remain, double nxt[3]) {
    const double a = (double)P.swing_time - remain, b = (double)P.swing_time;
    for (int i = 0; i < 3; ++i) nxt[i] = foot[i] + ((fin[i] - foot[i]) * a) / b;
    nxt[2] = (double)P.gait_height * sin(3.14159265358979323846 * a / b);
}

// x = (A + damp I)^-1 b for symmetric positive definite 3x3 A (adjugate form)
LEG_HD void solve_spd3(const double A[9], double damp, const double b[3], double x[3]) {
    const double a = A[0] + damp, d = A[4] + damp, f = A[8] + damp, bq = A[1], c = A[2], e = A[5];
    const double C00 = d * f - e * e, C01 = c * e - bq * f, C02 = bq * e - c * d;
    const double C11 = a * f - c * c, C12 = bq * c - a * e, C22 = a * d - bq * bq;
    const double det = a * C00 + bq * C01 + c * C02;
    const double id = 1.0 / det;
    x[0] = (C00 * b[0] + C01 * b[1] + C02 * b[2]) * id;
    x[1] = (C01 * b[0] + C11 * b[1] + C12 * b[2]) * id;
    x[2] = (C02 * b[0] + C12 * b[1] + C22 * b[2]) * id;
}

// Damped least-squares position IK of one leg (include/pinocchio_kinematics.h:61-149 with a 3-D position task):
//   repeat <= max_iter: e = target - p(q); stop if |e| < tol; v = J' (J J' + damp I)^-1 e; q += dt v
// Returns the number of iterations performed; err_out = |e| at exit (before the last update, as the reference
// reports the error it last computed).
LEG_HD int leg_ik(const LegModel& M, const SwingParams& P, int leg, const double base_pos[3], const double Rb[9],
                  const double target[3], double q[3], double& err_out) {
    int it = 0;
    double en = 0.0;
    for (; it < P.ik_max_iter; ++it) {
        double p[3], J[9], e[3], y[3], JJt[9];
        leg_fk_world(M, leg, base_pos, Rb, q, p, J);
        for (int i = 0; i < 3; ++i) e[i] = target[i] - p[i];
        en = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        if (en < P.ik_tol) break;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) JJt[i * 3 + j] = J[i * 3] * J[j * 3] + J[i * 3 + 1] * J[j * 3 + 1] + J[i * 3 + 2] * J[j * 3 + 2];
        solve_spd3(JJt, P.ik_damp, e, y);
        for (int k = 0; k < 3; ++k) q[k] += P.ik_dt * (J[k] * y[0] + J[3 + k] * y[1] + J[6 + k] * y[2]);
    }
    err_out = en;
    return it;
}

// ---- reference-literal 6-D task (include/pinocchio_kinematics.h:61-149) ---------------------------------------------
// Frame placement of the contact point in the world: position pw, rotation Rf (row-major) and the 6 x 3 frame Jacobian
// in LOCAL_WORLD_ALIGNED convention (rows 0-2 linear = z_k x (p - c_k), rows 3-5 angular = z_k), J[row * 3 + joint].
LEG_HD void leg_frame_world(const LegModel& M, int leg, const double base_pos[3], const double Rb[9], const double q[3],
                            double pw[3], double Rf[9], double J6[18]) {
    const double (*o)[3] = M.offset[leg];
    double R0[9], R1[9], R2[9], R01[9], R012[9];
    axis_angle(M.axis[leg][0], q[0], R0);
    axis_angle(M.axis[leg][1], q[1], R1);
    axis_angle(M.axis[leg][2], q[2], R2);
    mat3_mul(R0, R1, R01);
    mat3_mul(R01, R2, R012);
    const double tip[3] = {o[3][0] + o[4][0], o[3][1] + o[4][1], o[3][2] + o[4][2]};
    double t2[3], t1[3], t0[3], pb[3];
    mat3_vec(R012, tip, t2);
    mat3_vec(R01, o[2], t1);
    mat3_vec(R0, o[1], t0);
    const double c0[3] = {o[0][0], o[0][1], o[0][2]};
    const double c1[3] = {c0[0] + t0[0], c0[1] + t0[1], c0[2] + t0[2]};
    const double c2[3] = {c1[0] + t1[0], c1[1] + t1[1], c1[2] + t1[2]};
    for (int i = 0; i < 3; ++i) pb[i] = c2[i] + t2[i];
    double r[3];
    mat3_vec(Rb, pb, r);
    for (int i = 0; i < 3; ++i) pw[i] = base_pos[i] + r[i];
    mat3_mul(Rb, R012, Rf);            // link frames are axis-aligned at q = 0 (fixed joints carry no rotation)
    double z[3][3], zb1[3], zb2[3];
    mat3_vec(R0, M.axis[leg][1], zb1);
    mat3_vec(R01, M.axis[leg][2], zb2);
    mat3_vec(Rb, M.axis[leg][0], z[0]);
    mat3_vec(Rb, zb1, z[1]);
    mat3_vec(Rb, zb2, z[2]);
    const double* cs[3] = {c0, c1, c2};
    for (int k = 0; k < 3; ++k) {
        double db[3], d[3], lin[3];
        for (int i = 0; i < 3; ++i) db[i] = pb[i] - cs[k][i];
        mat3_vec(Rb, db, d);
        cross3(z[k], d, lin);
        for (int i = 0; i < 3; ++i) { J6[i * 3 + k] = lin[i]; J6[(3 + i) * 3 + k] = z[k][i]; }
    }
}

// below this angle the closed forms cancel catastrophically and the series are used (Pinocchio:
// TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4))
#define LEG_TAYLOR_EPS 1.220703125e-4
// w = log3(R): rotation vector of R (row-major), theta = |w|.  Pinocchio's log3: small-angle series below, the
// symmetric-part formula near pi (where the antisymmetric part of R vanishes).
LEG_HD void so3_log(const double R[9], double w[3], double& theta) {
    const double tr = R[0] + R[4] + R[8];
    double c = 0.5 * (tr - 1.0);
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double ax = R[7] - R[5], ay = R[2] - R[6], az = R[3] - R[1];   // 2 sin(theta) * axis
    const double s2 = sqrt(ax * ax + ay * ay + az * az);                   // 2 sin(theta)
    theta = atan2(0.5 * s2, c);
    if (theta < LEG_TAYLOR_EPS) {
        const double k = 0.5 * (1.0 + theta * theta / 6.0);
        w[0] = k * ax; w[1] = k * ay; w[2] = k * az;
    } else if (3.14159265358979323846 - theta > 1e-4) {
        const double k = theta / s2;
        w[0] = k * ax; w[1] = k * ay; w[2] = k * az;
    } else {
        // near pi: axis from the diagonal of (R + I)/2 = cos^2(theta/2) ... a a' (1 - c) + c I, signs from the antisymmetric part
        const double d = 1.0 / (1.0 - c);
        double x = sqrt(fmax((R[0] - c) * d, 0.0)), y = sqrt(fmax((R[4] - c) * d, 0.0)), zz = sqrt(fmax((R[8] - c) * d, 0.0));
        if (ax < 0.0) x = -x;
        if (ay < 0.0) y = -y;
        if (az < 0.0) zz = -zz;
        w[0] = theta * x; w[1] = theta * y; w[2] = theta * zz;
    }
}

// log6 of the placement (R, p): [v; w] with w = log3(R) and v = alpha p - w x p / 2 + beta (w.p) w  (Pinocchio log6)
LEG_HD void se3_log(const double R[9], const double p[3], double out[6]) {
    double w[3], t;
    so3_log(R, w, t);
    double alpha, beta;
    const double t2 = t * t;
    if (t < LEG_TAYLOR_EPS) { alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0; beta = 1.0 / 12.0 + t2 / 720.0; }
    else {
        double st, ct;
        sincos(t, &st, &ct);
        alpha = t * st / (2.0 * (1.0 - ct));
        beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
    }
    double wxp[3];
    cross3(w, p, wxp);
    const double wp = w[0] * p[0] + w[1] * p[1] + w[2] * p[2];
    for (int i = 0; i < 3; ++i) { out[i] = alpha * p[i] - 0.5 * wxp[i] + beta * wp * w[i]; out[3 + i] = w[i]; }
}

// Jlog3(theta, w): A = alpha w w' + diag + skew(w / 2)
LEG_HD void so3_jlog(double t, const double w[3], double A[9]) {
    double alpha, diag;
    const double t2 = t * t;
    if (t < LEG_TAYLOR_EPS) { alpha = 1.0 / 12.0 + t2 / 720.0; diag = 0.5 * (2.0 - t2 / 6.0); }
    else {
        double st, ct;
        sincos(t, &st, &ct);
        const double st_1mct = st / (1.0 - ct);
        alpha = 1.0 / t2 - st_1mct / (2.0 * t);
        diag = 0.5 * (t * st_1mct);
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i * 3 + j] = alpha * w[i] * w[j];
    A[0] += diag; A[4] += diag; A[8] += diag;
    A[1] -= 0.5 * w[2]; A[2] += 0.5 * w[1];
    A[3] += 0.5 * w[2]; A[5] -= 0.5 * w[0];
    A[6] -= 0.5 * w[1]; A[7] += 0.5 * w[0];
}

// Jlog6 of the placement (R, p), 6 x 6 row-major, blocks [[A, B], [0, A]] (Pinocchio Jlog6)
LEG_HD void se3_jlog(const double R[9], const double p[3], double Jl[36]) {
    double w[3], t, A[9], C[9], B[9];
    so3_log(R, w, t);
    so3_jlog(t, w, A);
    const double t2 = t * t;
    double beta, bdot;
    if (t < LEG_TAYLOR_EPS) { beta = 1.0 / 12.0 + t2 / 720.0; bdot = 1.0 / 360.0; }
    else {
        double st, ct;
        sincos(t, &st, &ct);
        const double tinv = 1.0 / t, t2inv = tinv * tinv, i22 = 1.0 / (2.0 * (1.0 - ct));
        beta = t2inv - st * tinv * i22;
        bdot = -2.0 * t2inv * t2inv + (1.0 + st * tinv) * t2inv * i22;
    }
    const double wp = w[0] * p[0] + w[1] * p[1] + w[2] * p[2];
    double v3[3];
    for (int i = 0; i < 3; ++i) v3[i] = (bdot * wp) * w[i] - (t2 * bdot + 2.0 * beta) * p[i];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i * 3 + j] = v3[i] * w[j] + beta * w[i] * p[j];
    C[0] += wp * beta; C[4] += wp * beta; C[8] += wp * beta;
    C[1] -= 0.5 * p[2]; C[2] += 0.5 * p[1];
    C[3] += 0.5 * p[2]; C[5] -= 0.5 * p[0];
    C[6] -= 0.5 * p[1]; C[7] += 0.5 * p[0];
    mat3_mul(C, A, B);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Jl[i * 6 + j] = A[i * 3 + j];
            Jl[i * 6 + 3 + j] = B[i * 3 + j];
            Jl[(3 + i) * 6 + j] = 0.0;
            Jl[(3 + i) * 6 + 3 + j] = A[i * 3 + j];
        }
}

// x = S^-1 b for a symmetric positive definite 6 x 6 S (LDL' without pivoting; the reference calls Eigen's ldlt())
LEG_HD void solve_spd6(double S[36], const double b[6], double x[6]) {
    double d[6];
    for (int j = 0; j < 6; ++j) {
        double v = S[j * 6 + j];
        for (int k = 0; k < j; ++k) v -= S[j * 6 + k] * S[j * 6 + k] * d[k];
        d[j] = v;
        for (int i = j + 1; i < 6; ++i) {
            double l = S[i * 6 + j];
            for (int k = 0; k < j; ++k) l -= S[i * 6 + k] * S[j * 6 + k] * d[k];
            S[i * 6 + j] = l / v;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= S[i * 6 + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 6; ++i) y[i] /= d[i];
    for (int i = 5; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 6; ++k) v -= S[k * 6 + i] * x[k]; x[i] = v; }
}

// The reference's inverseKinematics as written (include/pinocchio_kinematics.h:92-132) for the chain of one leg (the
// other leg's joints have zero Jacobian columns for this frame, so their velocity is exactly zero and they drop out).
LEG_HD int leg_ik6(const LegModel& M, const SwingParams& P, int leg, const double base_pos[3], const double Rb[9],
                   const double target[3], double q[3], double& err_out) {
    int it = 0;
    double en = 0.0;
    for (; it < P.ik_max_iter; ++it) {
        double p[3], Rf[9], J6[18], Ri[9], pi[3], err[6];
        leg_frame_world(M, leg, base_pos, Rb, q, p, Rf, J6);
        // iMd = oMf^-1 oMdes, oMdes = (I, target): rotation Rf', translation Rf' (target - p)      (:104)
        const double d[3] = {target[0] - p[0], target[1] - p[1], target[2] - p[2]};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Ri[i * 3 + j] = Rf[j * 3 + i];
            pi[i] = Rf[0 * 3 + i] * d[0] + Rf[1 * 3 + i] * d[1] + Rf[2 * 3 + i] * d[2];
        }
        se3_log(Ri, pi, err);                                                                          // (:105)
        en = 0.0;
        for (int i = 0; i < 6; ++i) en += err[i] * err[i];
        en = sqrt(en);
        if (en < P.ik_tol) break;                                                                      // (:108-111)
        // J = -Jlog6(iMd^-1) J_frame   (:113-116); iMd^-1 = (Rf, -Rf pi) = (Rf, -(target - p))
        double Jl[36], Jt[18], JJt[36], y[6];
        const double pinv[3] = {-d[0], -d[1], -d[2]};
        se3_jlog(Rf, pinv, Jl);
        for (int i = 0; i < 6; ++i)
            for (int k = 0; k < 3; ++k) {
                double v = 0.0;
                for (int m2 = 0; m2 < 6; ++m2) v += Jl[i * 6 + m2] * J6[m2 * 3 + k];
                Jt[i * 3 + k] = -v;
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) JJt[i * 6 + j] = Jt[i * 3] * Jt[j * 3] + Jt[i * 3 + 1] * Jt[j * 3 + 1] + Jt[i * 3 + 2] * Jt[j * 3 + 2];
        for (int i = 0; i < 6; ++i) JJt[i * 6 + i] += P.ik_damp;                                         // (:118-120)
        solve_spd6(JJt, err, y);
        for (int k = 0; k < 3; ++k) {                                                                    // (:122-124)
            double v = 0.0;
            for (int i = 0; i < 6; ++i) v += Jt[i * 3 + k] * y[i];
            q[k] += -v * P.ik_dt;
        }
    }
    err_out = en;
    return it;
}

LEG_HD int leg_ik_task(const LegModel& M, const SwingParams& P, int leg, const double base_pos[3], const double Rb[9],
                       const double target[3], double q[3], double& err_out) {
    return P.ik_mode == 1 ? leg_ik6(M, P, leg, base_pos, Rb, target, q, err_out)
                          : leg_ik(M, P, leg, base_pos, Rb, target, q, err_out);
}

// Joint torques that realise the ground-reaction force f (world frame, acting ON the foot) of a stance leg:
// tau = -J' f  (body of the reference's empty stub, include/MPCController.h:178-180)
LEG_HD void grf_to_torque(const double Jw[9], const double f[3], double tau[3]) {
    for (int k = 0; k < 3; ++k) tau[k] = -(Jw[k] * f[0] + Jw[3 + k] * f[1] + Jw[6 + k] * f[2]);
}

}  // namespace mpcb200
