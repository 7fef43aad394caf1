/* kf_oracle.c -- CPU oracle of the base-state Kalman filter.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Dense restatement of the reference's stateEstimator::update, include/stateEstimator.h:217-337 (paths relative to
 * /root/reference), with the matrices a, b, c, q, r built exactly as its constructor and update do (:184-213,
 * :221-260) and the two `s.lu().solve(...)` calls (:297,:300) as Gaussian elimination with partial pivoting
 * (Eigen's PartialPivLU).  The product uses the structure of a/c and a Cholesky factorisation instead.
 *
 * PARITY UNPINNED: the file is in none of the reference's build targets, depends on OCS2/Pinocchio/ROS (absent), and
 * does not compile/run as written (`q_.block(6,6,6,6) = dt * Matrix<12,12>::Identity()`, :227; a 4-vector of foot
 * heights streamed into a 14-vector, :209,:283).  Repairs: dt I_6 and two zero heights.  The feet come from this
 * build's leg model (leg_oracle.c), whose joint axes are an assumption (external URDF).
 */
#include <math.h>
#include <string.h>

#include "leg_oracle.h"

typedef struct orc_kf_params {
    double foot_radius, imu_process_noise_position, imu_process_noise_velocity, foot_process_noise_position;
    double foot_sensor_noise_position, foot_sensor_noise_velocity, foot_height_sensor_noise, high_suspect_number;
    int32_t accel_transpose;
} orc_kf_params;

void orc_kf_defaults(orc_kf_params *p) {
    p->foot_radius = 0.02; p->imu_process_noise_position = 0.02; p->imu_process_noise_velocity = 0.02;
    p->foot_process_noise_position = 0.002; p->foot_sensor_noise_position = 0.005; p->foot_sensor_noise_velocity = 0.1;
    p->foot_height_sensor_noise = 0.01; p->high_suspect_number = 100.0; p->accel_transpose = 1;
}

static void quat_rot(const double qx[4], double R[9]) {
    double n = sqrt(qx[0] * qx[0] + qx[1] * qx[1] + qx[2] * qx[2] + qx[3] * qx[3]);
    double x = qx[0] / n, y = qx[1] / n, z = qx[2] / n, w = qx[3] / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

/* X = S^-1 B for S n x n, B n x m (row-major), Gaussian elimination with partial pivoting */
static void lu_solve(int n, int m, const double *S, const double *Bm, double *X) {
    double A[14 * 14], R[14 * 14];
    memcpy(A, S, sizeof(double) * n * n);
    memcpy(R, Bm, sizeof(double) * n * m);
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
        if (piv != c) {
            for (int j = 0; j < n; ++j) { double t = A[c * n + j]; A[c * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < m; ++j) { double t = R[c * m + j]; R[c * m + j] = R[piv * m + j]; R[piv * m + j] = t; }
        }
        for (int r = c + 1; r < n; ++r) {
            double f = A[r * n + c] / A[c * n + c];
            for (int j = c; j < n; ++j) A[r * n + j] -= f * A[c * n + j];
            for (int j = 0; j < m; ++j) R[r * m + j] -= f * R[c * m + j];
        }
    }
    for (int j = 0; j < m; ++j)
        for (int i = n - 1; i >= 0; --i) {
            double s = R[i * m + j];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * X[k * m + j];
            X[i * m + j] = s / A[i * n + i];
        }
}

static void mm(int m, int n, int k, const double *A, const double *B, double *C) {   /* C(m x n) = A(m x k) B(k x n) */
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}

void orc_kf_update(const orc_kf_params *K, const orc_leg_model *L, double dt, const double quat[4], const double gyro[3],
                   const double accel_local[3], const double q[6], const double dq[6], const uint8_t contact[2], double xhat[12],
                   double P[144], double odom[13]) {
    double a[144] = {0}, b[36] = {0}, c[14 * 12] = {0}, q_[144] = {0}, qm[144] = {0}, r[196] = {0};
    /* constructor :194-213 */
    for (int i = 0; i < 12; ++i) { a[i * 12 + i] = 1.0; q_[i * 12 + i] = 1.0; }
    for (int i = 0; i < 3; ++i) {
        c[i * 12 + i] = 1.0; c[(3 + i) * 12 + i] = 1.0;                 /* c1 on rows 0-2 and 3-5 */
        c[(6 + i) * 12 + 3 + i] = 1.0; c[(9 + i) * 12 + 3 + i] = 1.0;   /* c2 on rows 6-8 and 9-11 */
    }
    for (int i = 0; i < 6; ++i) c[i * 12 + 6 + i] = -1.0;
    c[12 * 12 + 8] = 1.0; c[13 * 12 + 11] = 1.0;
    /* update :221-227 */
    for (int i = 0; i < 3; ++i) {
        a[i * 12 + 3 + i] = dt;
        b[i * 3 + i] = 0.5 * dt * dt; b[(3 + i) * 3 + i] = dt;
        q_[i * 12 + i] = dt / 20.f;
        q_[(3 + i) * 12 + 3 + i] = dt * 9.81f / 20.f;
    }
    for (int i = 6; i < 12; ++i) q_[i * 12 + i] = dt;
    /* kinematics with the base at the origin :232-246 */
    double R[9], ee_pos[6], ee_vel[6], zero[3] = {0, 0, 0}, wg[3];
    quat_rot(quat, R);
    for (int i = 0; i < 3; ++i) wg[i] = R[3 * i] * gyro[0] + R[3 * i + 1] * gyro[1] + R[3 * i + 2] * gyro[2];
    for (int l = 0; l < 2; ++l) {
        double J[9];
        orc_leg_fk(L, l, zero, quat, q + 3 * l, ee_pos + 3 * l, J);
        const double *p = ee_pos + 3 * l;
        double wxr[3] = {wg[1] * p[2] - wg[2] * p[1], wg[2] * p[0] - wg[0] * p[2], wg[0] * p[1] - wg[1] * p[0]};
        for (int i = 0; i < 3; ++i) ee_vel[3 * l + i] = wxr[i] + J[3 * i] * dq[3 * l] + J[3 * i + 1] * dq[3 * l + 1] + J[3 * i + 2] * dq[3 * l + 2];
    }
    /* noise :248-270 */
    for (int i = 0; i < 12; ++i) qm[i * 12 + i] = 1.0;
    for (int i = 0; i < 3; ++i) {
        qm[i * 12 + i] = q_[i * 12 + i] * K->imu_process_noise_position;
        qm[(3 + i) * 12 + 3 + i] = q_[(3 + i) * 12 + 3 + i] * K->imu_process_noise_velocity;
    }
    for (int i = 6; i < 12; ++i) qm[i * 12 + i] = q_[i * 12 + i] * K->foot_process_noise_position;
    for (int i = 0; i < 14; ++i) r[i * 14 + i] = i < 6 ? K->foot_sensor_noise_position : (i < 12 ? K->foot_sensor_noise_velocity : K->foot_height_sensor_noise);
    double ps[6], vs[6];
    for (int i = 0; i < 2; ++i) {
        const double f = contact[i] ? 1.0 : K->high_suspect_number;
        for (int k = 0; k < 3; ++k) {
            qm[(6 + 3 * i + k) * 12 + 6 + 3 * i + k] *= f;
            r[(3 * i + k) * 14 + 3 * i + k] *= f;
            r[(6 + 3 * i + k) * 14 + 6 + 3 * i + k] *= f;
            ps[3 * i + k] = -ee_pos[3 * i + k];
            vs[3 * i + k] = -ee_vel[3 * i + k];
        }
        r[(12 + i) * 14 + 12 + i] *= f;
        ps[3 * i + 2] += K->foot_radius;
    }
    /* :280-304 */
    double accel[3];
    for (int i = 0; i < 3; ++i)
        accel[i] = K->accel_transpose ? (R[i] * accel_local[0] + R[3 + i] * accel_local[1] + R[6 + i] * accel_local[2])
                                      : (R[3 * i] * accel_local[0] + R[3 * i + 1] * accel_local[1] + R[3 * i + 2] * accel_local[2]);
    accel[2] += -9.81;
    double y[14], x[12], at[144], ct[12 * 14], tmp[144], pm[144], ymod[14], ey[14], s[196], t1[14 * 12], t2[12 * 14];
    for (int i = 0; i < 6; ++i) { y[i] = ps[i]; y[6 + i] = vs[i]; }
    y[12] = 0.0; y[13] = 0.0;
    for (int i = 0; i < 12; ++i) {
        double v = 0.0;
        for (int k = 0; k < 12; ++k) v += a[i * 12 + k] * xhat[k];
        for (int k = 0; k < 3; ++k) v += b[i * 3 + k] * accel[k];
        x[i] = v;
    }
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) at[i * 12 + j] = a[j * 12 + i];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 14; ++j) ct[i * 14 + j] = c[j * 12 + i];
    mm(12, 12, 12, a, P, tmp); mm(12, 12, 12, tmp, at, pm);
    for (int i = 0; i < 144; ++i) pm[i] += qm[i];
    mm(14, 1, 12, c, x, ymod);
    for (int i = 0; i < 14; ++i) ey[i] = y[i] - ymod[i];
    mm(14, 12, 12, c, pm, t1); mm(14, 14, 12, t1, ct, s);
    for (int i = 0; i < 196; ++i) s[i] += r[i];
    double sEy[14], sC[14 * 12], pmct[12 * 14], corr[12], gc[144], imgc[144], pn[144];
    lu_solve(14, 1, s, ey, sEy);
    mm(12, 14, 12, pm, ct, pmct);
    mm(12, 1, 14, pmct, sEy, corr);
    for (int i = 0; i < 12; ++i) x[i] += corr[i];
    lu_solve(14, 12, s, c, sC);
    mm(12, 12, 14, pmct, sC, gc);
    for (int i = 0; i < 144; ++i) imgc[i] = ((i / 12 == i % 12) ? 1.0 : 0.0) - gc[i];
    mm(12, 12, 12, imgc, pm, pn);
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) t2[i * 12 + j] = (pn[i * 12 + j] + pn[j * 12 + i]) / 2.0;
    memcpy(pn, t2, sizeof(double) * 144);
    if (pn[0] * pn[13] - pn[1] * pn[12] > 0.000001) {
        for (int i = 0; i < 2; ++i) for (int j = 2; j < 12; ++j) { pn[i * 12 + j] = 0.0; pn[j * 12 + i] = 0.0; }
        pn[0] /= 10.; pn[1] /= 10.; pn[12] /= 10.; pn[13] /= 10.;
    }
    memcpy(xhat, x, sizeof(x));
    memcpy(P, pn, sizeof(pn));
    if (odom) {
        for (int i = 0; i < 3; ++i) odom[i] = x[i];
        for (int i = 0; i < 4; ++i) odom[3 + i] = quat[i];
        for (int i = 0; i < 3; ++i) odom[7 + i] = R[i] * x[3] + R[3 + i] * x[4] + R[6 + i] * x[5];
        for (int i = 0; i < 3; ++i) odom[10 + i] = gyro[i];
    }
}
