// kf_core.cuh -- per-robot linear Kalman filter of the base position/velocity and the two foot positions:
// the reference's (unbuilt, OCS2-derived) stateEstimator::update, include/stateEstimator.h:217-337
// (SURVEY.md 8f rank 4).  Group-cooperative code like tron1_core.cuh: one warp per robot on the GPU, one serial
// thread in the host test build.
//
// State  xHat[12] = [p(3), v(3), foot_L(3), foot_R(3)],  covariance P (12x12, row-major; symmetric).
// Measurement y[14] = [-eePos_L, -eePos_R (+foot radius on z), -eeVel_L, -eeVel_R, 0, 0] where eePos/eeVel are the
// feet relative to the base in world-aligned axes from the leg kinematics (the reference calls Pinocchio with the
// base at the origin, :232-246; here leg_core.cuh's closed-form FK).
//
// Structure used instead of dense 12x12 / 14x12 products (the constructor, include/stateEstimator.h:194-213):
//   a = I + dt E_{p<-v},  b = [dt^2/2 I; dt I; 0],  c rows: 0-2 p - foot_L, 3-5 p - foot_R, 6-8 v, 9-11 v,
//   12 foot_L z, 13 foot_R z;  q, r diagonal.
// so  a P a'  is two rank-3 row/column updates,  c M  and  M c'  are row / column differences, and only the
// 14x14 innovation covariance needs a factorisation (Cholesky here; the reference uses two LU solves, :297-301).
//
// Restated as written: noise constants (:124-130), `dt/20.f` and `dt*9.81f/20.f` float literals (:225-226), the
// x100 inflation for a foot without contact (:262-270), `accel = Rzyx' * a_local + g` WITH the transpose the
// reference applies (:281; switchable), symmetrisation and the 2x2 covariance reset (:303-310), the body-frame
// twist output (:326).  Repaired (the file is in no build target and does not compile/run as written):
// `q_.block(6,6,6,6) = dt * Matrix<12,12>::Identity()` is taken as dt I_6 (:227), and `y << ps_, vs_, feetHeights_`
// with a 4-vector of heights (:283, :209) is taken as two zero heights.
#pragma once
#include <math.h>
#include <stdint.h>

#include "leg_core.cuh"

#if defined(__CUDACC__)
#define KF_HD __host__ __device__ __forceinline__
#else
#define KF_HD inline
#endif

namespace mpcb200 {

struct KfParams {
    double foot_radius;                  // 0.02   (:124)
    double imu_noise_pos, imu_noise_vel; // 0.02, 0.02
    double foot_noise_pos;               // 0.002
    double foot_sensor_pos, foot_sensor_vel, foot_height_noise;   // 0.005, 0.1, 0.01
    double suspect;                      // 100 (:262)
    int accel_transpose;                 // 1 = as written (:281)
};

struct KfWork {            // per robot, shared memory on the device
    double M[144];         // pm = a P a' + q, then the updated covariance
    double S[14 * 15];     // innovation covariance (lower triangle, Cholesky in place) | column 14: e_y
    double K[12 * 14];     // pm c'  (12 x 14), then the gain rows solved against S
    double x[12], y[14], r[14];
    double ee[12];         // eePos (6), eeVel (6)
    double Rb[9], acc[3];
};

// rows of c M for a 12-column matrix M (row-major 12x12): (c M)[i][j]
KF_HD double kf_cM(const double* M, int i, int j) {
    if (i < 3) return M[i * 12 + j] - M[(6 + i) * 12 + j];
    if (i < 6) return M[(i - 3) * 12 + j] - M[(6 + i) * 12 + j];      // rows 9..11 = foot_R
    if (i < 9) return M[(i - 3) * 12 + j];                             // v rows 3..5
    if (i < 12) return M[(i - 6) * 12 + j];
    return M[(i == 12 ? 8 : 11) * 12 + j];
}
// (M c')[i][j] for symmetric use: column combination, M row-major 12x12
KF_HD double kf_McT(const double* M, int i, int j) {
    const double* r = M + i * 12;
    if (j < 3) return r[j] - r[6 + j];
    if (j < 6) return r[j - 3] - r[6 + j];
    if (j < 9) return r[j - 3];
    if (j < 12) return r[j - 6];
    return r[j == 12 ? 8 : 11];
}

// One filter update.  xhat[12], P[144] in/out (global memory), odom[13] out = pos 3, quat 4 (x,y,z,w), v_pos 3
// (body frame, :326), v_ori 3 (= gyro).  contact[2]: 1 = foot in contact.
template <class G>
KF_HD void kf_update(const KfParams& K, const LegModel& L, double dt, const double* quat, const double* gyro, const double* accel,
                     const double* q, const double* dq, const uint8_t* contact, double* xhat, double* P, double* odom,
                     KfWork& W, const G& g) {
    const int t = g.tid(), nt = g.size();
    // ---- kinematics: feet relative to the base, world-aligned (:232-246, :272-274) -----------------------------
    if (t == 0) quat_to_rot(quat, W.Rb);
    g.sync();
    for (int leg = t; leg < 2; leg += nt) {
        double pb[3], Jb[9], vb[3], wl[3] = {gyro[0], gyro[1], gyro[2]}, tmp[3];
        leg_fk_base(L, leg, q + 3 * leg, pb, Jb);
        for (int i = 0; i < 3; ++i) vb[i] = Jb[i * 3] * dq[3 * leg] + Jb[i * 3 + 1] * dq[3 * leg + 1] + Jb[i * 3 + 2] * dq[3 * leg + 2];
        cross3(wl, pb, tmp);                               // base-frame: omega_local x r + J dq, then rotate
        for (int i = 0; i < 3; ++i) vb[i] += tmp[i];
        mat3_vec(W.Rb, pb, W.ee + 3 * leg);
        mat3_vec(W.Rb, vb, W.ee + 6 + 3 * leg);
    }
    if (t == 0) {
        const double* R = W.Rb;
        for (int i = 0; i < 3; ++i) {
            W.acc[i] = K.accel_transpose ? (R[i] * accel[0] + R[3 + i] * accel[1] + R[6 + i] * accel[2])
                                         : (R[3 * i] * accel[0] + R[3 * i + 1] * accel[1] + R[3 * i + 2] * accel[2]);
        }
        W.acc[2] += -9.81;
    }
    g.sync();
    // ---- measurement, noise, prediction of the mean (:248-284) -------------------------------------------------------
    const double qp = (dt / 20.f) * K.imu_noise_pos, qv = (dt * 9.81f / 20.f) * K.imu_noise_vel, qf = dt * K.foot_noise_pos;
    for (int i = t; i < 14; i += nt) {
        const int foot = i < 12 ? (i % 6) / 3 : i - 12;
        const double infl = contact[foot] ? 1.0 : K.suspect;
        double yv, rv;
        if (i < 6) { yv = -W.ee[i] + ((i % 3) == 2 ? K.foot_radius : 0.0); rv = K.foot_sensor_pos; }
        else if (i < 12) { yv = -W.ee[i]; rv = K.foot_sensor_vel; }
        else { yv = 0.0; rv = K.foot_height_noise; }
        W.y[i] = yv;
        W.r[i] = rv * infl;
    }
    for (int i = t; i < 12; i += nt) {
        double v = xhat[i];
        if (i < 3) v += dt * xhat[3 + i] + 0.5 * dt * dt * W.acc[i];
        else if (i < 6) v += dt * W.acc[i - 3];
        W.x[i] = v;
    }
    // ---- pm = a P a' + q   (a = I + dt E: rows/columns 0..2 pick up dt * rows/columns 3..5) (:286) ---------------------
    for (int idx = t; idx < 144; idx += nt) {
        const int i = idx / 12, j = idx % 12;
        double v = P[idx];
        if (i < 3) v += dt * P[(3 + i) * 12 + j];
        if (j < 3) v += dt * P[i * 12 + 3 + j];
        if (i < 3 && j < 3) v += dt * dt * P[(3 + i) * 12 + 3 + j];
        if (i == j) {
            const double infl = i >= 6 ? (contact[(i - 6) / 3] ? 1.0 : K.suspect) : 1.0;
            v += i < 3 ? qp : (i < 6 ? qv : qf * infl);
        }
        W.M[idx] = v;
    }
    g.sync();
    // ---- K0 = pm c' (12 x 14),  S = c pm c' + r (lower triangle),  e_y = y - c x (:287-292) -----------------------------
    for (int idx = t; idx < 12 * 14; idx += nt) W.K[idx] = kf_McT(W.M, idx / 14, idx % 14);
    for (int i = t; i < 14; i += nt) {
        double cx;
        if (i < 3) cx = W.x[i] - W.x[6 + i];
        else if (i < 6) cx = W.x[i - 3] - W.x[6 + i];
        else if (i < 9) cx = W.x[i - 3];
        else if (i < 12) cx = W.x[i - 6];
        else cx = W.x[i == 12 ? 8 : 11];
        W.S[i * 15 + 14] = W.y[i] - cx;
    }
    g.sync();
    for (int idx = t; idx < 14 * 14; idx += nt) {
        const int i = idx / 14, j = idx % 14;
        if (j > i) continue;
        // (c pm c')[i][j] = (c (pm c'))[i][j]: row combination of K0's column j
        double v;
        if (i < 3) v = W.K[i * 14 + j] - W.K[(6 + i) * 14 + j];
        else if (i < 6) v = W.K[(i - 3) * 14 + j] - W.K[(6 + i) * 14 + j];
        else if (i < 9) v = W.K[(i - 3) * 14 + j];
        else if (i < 12) v = W.K[(i - 6) * 14 + j];
        else v = W.K[(i == 12 ? 8 : 11) * 14 + j];
        if (i == j) v += W.r[i];
        W.S[i * 15 + j] = v;
    }
    g.sync();
    // ---- Cholesky S = L L' in place (column by column), then L z = e_y -------------------------------------------------
    for (int k = 0; k < 14; ++k) {
        if (t == 0) W.S[k * 15 + k] = sqrt(W.S[k * 15 + k]);
        g.sync();
        const double d = W.S[k * 15 + k];
        for (int i = k + 1 + t; i < 14; i += nt) W.S[i * 15 + k] /= d;
        g.sync();
        for (int idx = t; idx < (13 - k) * (13 - k); idx += nt) {
            const int i = k + 1 + idx / (13 - k), j = k + 1 + idx % (13 - k);
            if (j <= i) W.S[i * 15 + j] -= W.S[i * 15 + k] * W.S[j * 15 + k];
        }
        g.sync();
    }
    // ---- gain rows: G = K0 S^-1  (each of the 12 rows solved against L L'; S is symmetric) -----------------------------
    for (int row = t; row < 12; row += nt) {
        double v[14];
        for (int j = 0; j < 14; ++j) v[j] = W.K[row * 14 + j];
        for (int j = 0; j < 14; ++j) {                      // forward: L w = v
            double s = v[j];
            for (int k = 0; k < j; ++k) s -= W.S[j * 15 + k] * v[k];
            v[j] = s / W.S[j * 15 + j];
        }
        for (int j = 13; j >= 0; --j) {                     // backward: L' u = w
            double s = v[j];
            for (int k = j + 1; k < 14; ++k) s -= W.S[k * 15 + j] * v[k];
            v[j] = s / W.S[j * 15 + j];
        }
        for (int j = 0; j < 14; ++j) W.K[row * 14 + j] = v[j];
    }
    g.sync();
    // ---- x += pm c' S^-1 e_y = G e_y  (:297-298) -----------------------------------------------------------------------
    for (int i = t; i < 12; i += nt) {
        double s = W.x[i];
        for (int j = 0; j < 14; ++j) s += W.K[i * 14 + j] * W.S[j * 15 + 14];
        xhat[i] = s;
        W.x[i] = s;
    }
    // ---- P = (I - G c) pm, symmetrised (:300-304) ------------------------------------------------------------------------
    for (int idx = t; idx < 144; idx += nt) {
        const int i = idx / 12, j = idx % 12;
        double s = W.M[idx];
        for (int k = 0; k < 14; ++k) s -= W.K[i * 14 + k] * kf_cM(W.M, k, j);
        P[idx] = s;
    }
    g.sync();
    for (int idx = t; idx < 144; idx += nt) {
        const int i = idx / 12, j = idx % 12;
        if (j < i) { const double s = 0.5 * (P[idx] + P[j * 12 + i]); W.M[idx] = s; W.M[j * 12 + i] = s; }
        else if (i == j) W.M[idx] = P[idx];
    }
    g.sync();
    // ---- the reference's 2x2 covariance reset (:306-310) ----------------------------------------------------------------
    const bool reset = W.M[0] * W.M[13] - W.M[1] * W.M[12] > 0.000001;
    for (int idx = t; idx < 144; idx += nt) {
        const int i = idx / 12, j = idx % 12;
        double v = W.M[idx];
        if (reset) {
            if ((i < 2) != (j < 2)) v = 0.0;
            else if (i < 2 && j < 2) v /= 10.0;
        }
        P[idx] = v;
    }
    // ---- odometry output (:319-333) ----------------------------------------------------------------------------------------
    if (t == 0 && odom) {
        const double* R = W.Rb;
        for (int i = 0; i < 3; ++i) odom[i] = W.x[i];
        for (int i = 0; i < 4; ++i) odom[3 + i] = quat[i];
        for (int i = 0; i < 3; ++i) odom[7 + i] = R[i] * W.x[3] + R[3 + i] * W.x[4] + R[6 + i] * W.x[5];   // R' v
        for (int i = 0; i < 3; ++i) odom[10 + i] = gyro[i];
    }
    g.sync();
}

}  // namespace mpcb200
