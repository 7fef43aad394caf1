/* leg_oracle.c -- see leg_oracle.h.  TEST INFRASTRUCTURE, NOT PRODUCT.  PARITY UNPINNED (external URDF). */
#include "leg_oracle.h"

#include <math.h>
#include <string.h>

#include "mpc_oracle.h"

/* ---- 4x4 homogeneous transforms, row-major ---- */
typedef struct { double m[16]; } T4;

static T4 t4_identity(void) { T4 t; memset(&t, 0, sizeof(t)); t.m[0] = t.m[5] = t.m[10] = t.m[15] = 1.0; return t; }
static T4 t4_mul(const T4 *a, const T4 *b) {
    T4 c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a->m[4 * i + k] * b->m[4 * k + j];
            c.m[4 * i + j] = s;
        }
    return c;
}
static T4 t4_translate(const double v[3]) { T4 t = t4_identity(); t.m[3] = v[0]; t.m[7] = v[1]; t.m[11] = v[2]; return t; }
/* rotation from a unit quaternion (w, x, y, z) */
static T4 t4_from_quat(double w, double x, double y, double z) {
    T4 t = t4_identity();
    t.m[0] = w * w + x * x - y * y - z * z; t.m[1] = 2 * (x * y - w * z);           t.m[2] = 2 * (x * z + w * y);
    t.m[4] = 2 * (x * y + w * z);           t.m[5] = w * w - x * x + y * y - z * z; t.m[6] = 2 * (y * z - w * x);
    t.m[8] = 2 * (x * z - w * y);           t.m[9] = 2 * (y * z + w * x);           t.m[10] = w * w - x * x - y * y + z * z;
    return t;
}
static T4 t4_axis_angle(const double a[3], double q) {
    const double h = 0.5 * q, s = sin(h);
    return t4_from_quat(cos(h), a[0] * s, a[1] * s, a[2] * s);
}
static T4 t4_base(const double pos[3], const double qx[4]) {
    double n = sqrt(qx[0] * qx[0] + qx[1] * qx[1] + qx[2] * qx[2] + qx[3] * qx[3]);
    T4 r = (n > 0.0) ? t4_from_quat(qx[3] / n, qx[0] / n, qx[1] / n, qx[2] / n) : t4_identity();
    r.m[3] = pos[0]; r.m[7] = pos[1]; r.m[11] = pos[2];
    return r;
}

void orc_leg_defaults(orc_leg_model *m, orc_swing_params *p) {
    /* include/MPCParam.h:13-38 */
    const double abad[3] = {0.05556, 0.105, -0.2602}, hip[3] = {-0.077, 0.02050, 0.0}, knee[3] = {-0.1500, -0.02050, -0.25981};
    const double foot[3] = {0.145, 0.0, -0.2598}, contact[3] = {0.0, 0.0, -0.032};
    const double *o[5] = {abad, hip, knee, foot, contact};
    if (m) {
        for (int k = 0; k < 5; ++k)
            for (int i = 0; i < 3; ++i) {
                m->offset[1][k][i] = o[k][i];
                /* left leg: include/MPCParam.h:65 flips the sign of the abad, hip and knee y offsets */
                m->offset[0][k][i] = (i == 1 && k < 3) ? -o[k][i] : o[k][i];
            }
        const double ax[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 1, 0}};
        for (int l = 0; l < 2; ++l) memcpy(m->axis[l], ax, sizeof(ax));
    }
    if (p) {
        memset(p, 0, sizeof(*p));
        p->dt = 0.001f; p->swing_time = 0.5f; p->stance_time = 0.5f; p->gait_height = 0.1f;
        p->p_rel_max = 0.3;
        /* include/MPCParam.h:64-73, summed in the reference's order */
        p->foot_offset_left[0] = abad[0] + hip[0] + knee[0] + foot[0] + contact[0];
        p->foot_offset_left[1] = -abad[1] - hip[1] - knee[1] + foot[1] + contact[1];
        p->foot_offset_left[2] = abad[2] + hip[2] + knee[2] + foot[2] + contact[2];
        p->foot_offset_right[0] = abad[0] + hip[0] + knee[0] + foot[0] + contact[0];
        p->foot_offset_right[1] = abad[1] + hip[1] + knee[1] + foot[1] + contact[1];
        p->foot_offset_right[2] = abad[2] + hip[2] + knee[2] + foot[2] + contact[2];
        p->ik_tol = 1e-3; p->ik_dt = 1e-1; p->ik_damp = 1e-6; p->ik_max_iter = 10;
    }
}

void orc_leg_fk(const orc_leg_model *m, int leg, const double base_pos[3], const double quat[4], const double q[3],
                double p[3], double *J) {
    /* world <- base <- abad joint <- hip joint <- knee joint <- foot <- contact; joint frames keep the origin and axis */
    T4 frames[4];   /* world placement of joint k's frame AFTER its rotation, k = 0..2; [3] = contact */
    T4 cur = t4_base(base_pos, quat);
    double origin[3][3], zaxis[3][3];
    for (int k = 0; k < 3; ++k) {
        T4 tr = t4_translate(m->offset[leg][k]);
        cur = t4_mul(&cur, &tr);
        for (int i = 0; i < 3; ++i) {
            origin[k][i] = cur.m[4 * i + 3];
            zaxis[k][i] = cur.m[4 * i] * m->axis[leg][k][0] + cur.m[4 * i + 1] * m->axis[leg][k][1] + cur.m[4 * i + 2] * m->axis[leg][k][2];
        }
        T4 rot = t4_axis_angle(m->axis[leg][k], q[k]);
        cur = t4_mul(&cur, &rot);
        frames[k] = cur;
    }
    T4 tf = t4_translate(m->offset[leg][3]);
    cur = t4_mul(&cur, &tf);
    T4 tc = t4_translate(m->offset[leg][4]);
    cur = t4_mul(&cur, &tc);
    frames[3] = cur;
    for (int i = 0; i < 3; ++i) p[i] = cur.m[4 * i + 3];
    if (J) {
        for (int k = 0; k < 3; ++k) {
            const double d[3] = {p[0] - origin[k][0], p[1] - origin[k][1], p[2] - origin[k][2]};
            const double *z = zaxis[k];
            J[0 * 3 + k] = z[1] * d[2] - z[2] * d[1];
            J[1 * 3 + k] = z[2] * d[0] - z[0] * d[2];
            J[2 * 3 + k] = z[0] * d[1] - z[1] * d[0];
        }
    }
    (void)frames;
}

void orc_foot_placement(const orc_swing_params *p, const double pos[3], const double des_v[3], double remain,
                        int left_leg_state, double fin[3]) {
    /* include/MPCController.h:106-132 */
    double predicted[3] = {pos[0] + des_v[0] * remain, pos[1] + des_v[1] * remain, pos[2] + des_v[2] * remain};
    double p_rel_max = p->p_rel_max;
    double pfx_rel = des_v[0] * 0.5 * p->stance_time;
    double pfy_rel = des_v[1] * 0.5 * p->stance_time;
    pfx_rel = fmin(fmax(pfx_rel, -p_rel_max), p_rel_max);
    pfy_rel = fmin(fmax(pfy_rel, -p_rel_max), p_rel_max);
    predicted[0] += pfx_rel;
    predicted[1] += pfy_rel;
    predicted[2] = 0;
    if (left_leg_state == 1) { fin[0] = predicted[0] + p->foot_offset_left[0]; fin[1] = predicted[1] + p->foot_offset_left[1]; }
    else { fin[0] = predicted[0] + p->foot_offset_right[0]; fin[1] = predicted[1] + p->foot_offset_right[1]; }
    fin[2] = 0.0;   /* never assigned by the reference; overwritten by the sine profile */
}

void orc_swing_next(const orc_swing_params *p, const double foot[3], const double fin[3], double remain, double nxt[3]) {
    /* include/MPCController.h:155-158; float members promote to double inside the mixed expressions */
    for (int i = 0; i < 3; ++i) nxt[i] = foot[i] + (fin[i] - foot[i]) * (p->swing_time - remain) / p->swing_time;
    nxt[2] = p->gait_height * sin(M_PI * (p->swing_time - remain) / p->swing_time);
}

static void solve3(const double A[9], const double b[3], double x[3]) {
    /* Gaussian elimination with partial pivoting (the reference uses an LDLT, include/pinocchio_kinematics.h:123) */
    double M[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M[i][j] = A[3 * i + j]; M[i][3] = b[i]; }
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r = c + 1; r < 3; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 4; ++j) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        for (int r = c + 1; r < 3; ++r) {
            double f = M[r][c] / M[c][c];
            for (int j = c; j < 4; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 2; i >= 0; --i) {
        double s = M[i][3];
        for (int j = i + 1; j < 3; ++j) s -= M[i][j] * x[j];
        x[i] = s / M[i][i];
    }
}

int orc_leg_ik(const orc_leg_model *m, const orc_swing_params *p, int leg, const double base_pos[3], const double quat[4],
               const double target[3], double q[3], double *err) {
    int it = 0;
    double en = 0.0;
    for (; it < p->ik_max_iter; ++it) {
        double pos[3], J[9], e[3], JJt[9], y[3];
        orc_leg_fk(m, leg, base_pos, quat, q, pos, J);
        for (int i = 0; i < 3; ++i) e[i] = target[i] - pos[i];
        en = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        if (en < p->ik_tol) break;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += J[3 * i + k] * J[3 * j + k];
                JJt[3 * i + j] = s + (i == j ? p->ik_damp : 0.0);
            }
        solve3(JJt, e, y);
        for (int k = 0; k < 3; ++k) q[k] += p->ik_dt * (J[k] * y[0] + J[3 + k] * y[1] + J[6 + k] * y[2]);
    }
    if (err) *err = en;
    return it;
}

/* ---- the reference's 6-D IK task ------------------------------------------------------------------------------------- */
static void gauss_solve(int n, double *A, double *B, int nrhs) {   /* A (n x n) and B (n x nrhs) row-major, B <- A^-1 B */
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
        if (piv != c) {
            for (int j = 0; j < n; ++j) { double t = A[c * n + j]; A[c * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < nrhs; ++j) { double t = B[c * nrhs + j]; B[c * nrhs + j] = B[piv * nrhs + j]; B[piv * nrhs + j] = t; }
        }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = A[r * n + c] / A[c * n + c];
            for (int j = c; j < n; ++j) A[r * n + j] -= f * A[c * n + j];
            for (int j = 0; j < nrhs; ++j) B[r * nrhs + j] -= f * B[c * nrhs + j];
        }
    }
    for (int r = 0; r < n; ++r) for (int j = 0; j < nrhs; ++j) B[r * nrhs + j] /= A[r * n + r];
}

static void skew3(const double w[3], double S[9]) {
    S[0] = 0; S[1] = -w[2]; S[2] = w[1]; S[3] = w[2]; S[4] = 0; S[5] = -w[0]; S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}

void orc_se3_log(const double R[9], const double t[3], double xi[6]) {
    /* rotation vector through the unit quaternion of R (Shepperd's branch on the largest of w, x, y, z) */
    double qw, qx, qy, qz;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) { double s = sqrt(tr + 1.0) * 2.0; qw = 0.25 * s; qx = (R[7] - R[5]) / s; qy = (R[2] - R[6]) / s; qz = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; qw = (R[7] - R[5]) / s; qx = 0.25 * s; qy = (R[1] + R[3]) / s; qz = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; qw = (R[2] - R[6]) / s; qx = (R[1] + R[3]) / s; qy = 0.25 * s; qz = (R[5] + R[7]) / s; }
    else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; qw = (R[3] - R[1]) / s; qx = (R[2] + R[6]) / s; qy = (R[5] + R[7]) / s; qz = 0.25 * s; }
    if (qw < 0.0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
    const double sn = sqrt(qx * qx + qy * qy + qz * qz);
    const double th = 2.0 * atan2(sn, qw);
    const double k = sn > 1e-12 ? th / sn : 2.0;
    const double w[3] = {k * qx, k * qy, k * qz};
    /* t = V(w) v with V = I + a [w]x + b [w]x^2, a = (1 - cos)/th^2, b = (th - sin)/th^3 (series for small angles) */
    double a, b;
    if (th < 1e-4) { a = 0.5 - th * th / 24.0; b = 1.0 / 6.0 - th * th / 120.0; }
    else { a = (1.0 - cos(th)) / (th * th); b = (th - sin(th)) / (th * th * th); }
    double S[9], S2[9], V[9], v[3] = {t[0], t[1], t[2]};
    skew3(w, S);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double q = 0; for (int m2 = 0; m2 < 3; ++m2) q += S[3 * i + m2] * S[3 * m2 + j]; S2[3 * i + j] = q; }
    for (int i = 0; i < 9; ++i) V[i] = a * S[i] + b * S2[i];
    V[0] += 1.0; V[4] += 1.0; V[8] += 1.0;
    gauss_solve(3, V, v, 1);
    for (int i = 0; i < 3; ++i) { xi[i] = v[i]; xi[3 + i] = w[i]; }
}

void orc_se3_jlog(const double R[9], const double t[3], double Jl[36]) {
    /* Jlog6(M) = Jr(xi)^-1, xi = log6(M), Jr(xi) = sum_k (-ad_xi)^k / (k+1)!, ad_xi = [[wx, vx], [0, wx]] ([v; w] ordering) */
    double xi[6], ad[36], term[36], Jr[36], tmp[36];
    orc_se3_log(R, t, xi);
    double Sv[9], Sw[9];
    skew3(xi, Sv); skew3(xi + 3, Sw);
    memset(ad, 0, sizeof(ad));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        ad[6 * i + j] = Sw[3 * i + j]; ad[6 * i + 3 + j] = Sv[3 * i + j]; ad[6 * (3 + i) + 3 + j] = Sw[3 * i + j];
    }
    memset(Jr, 0, sizeof(Jr)); memset(term, 0, sizeof(term));
    for (int i = 0; i < 6; ++i) { Jr[7 * i] = 1.0; term[7 * i] = 1.0; }
    for (int k = 1; k < 60; ++k) {
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
            double q = 0; for (int m2 = 0; m2 < 6; ++m2) q += term[6 * i + m2] * ad[6 * m2 + j];
            tmp[6 * i + j] = -q / (double)(k + 1);
        }
        memcpy(term, tmp, sizeof(term));
        for (int i = 0; i < 36; ++i) Jr[i] += term[i];
    }
    memset(Jl, 0, sizeof(double) * 36);
    for (int i = 0; i < 6; ++i) Jl[7 * i] = 1.0;
    gauss_solve(6, Jr, Jl, 6);
}

int orc_leg_ik6(const orc_leg_model *m, const orc_swing_params *p, int leg, const double base_pos[3], const double quat[4],
                const double target[3], double q[3], double *err) {
    int it = 0;
    double en = 0.0;
    for (; it < p->ik_max_iter; ++it) {
        /* frame placement and LOCAL_WORLD_ALIGNED Jacobian by the 4x4 chain */
        T4 cur = t4_base(base_pos, quat);
        double origin[3][3], zaxis[3][3];
        for (int k = 0; k < 3; ++k) {
            T4 tr = t4_translate(m->offset[leg][k]);
            cur = t4_mul(&cur, &tr);
            for (int i = 0; i < 3; ++i) {
                origin[k][i] = cur.m[4 * i + 3];
                zaxis[k][i] = cur.m[4 * i] * m->axis[leg][k][0] + cur.m[4 * i + 1] * m->axis[leg][k][1] + cur.m[4 * i + 2] * m->axis[leg][k][2];
            }
            T4 rot = t4_axis_angle(m->axis[leg][k], q[k]);
            cur = t4_mul(&cur, &rot);
        }
        T4 tf = t4_translate(m->offset[leg][3]); cur = t4_mul(&cur, &tf);
        T4 tc = t4_translate(m->offset[leg][4]); cur = t4_mul(&cur, &tc);
        double Rf[9], pf[3], J6[18];
        for (int i = 0; i < 3; ++i) { pf[i] = cur.m[4 * i + 3]; for (int j = 0; j < 3; ++j) Rf[3 * i + j] = cur.m[4 * i + j]; }
        for (int k = 0; k < 3; ++k) {
            const double d[3] = {pf[0] - origin[k][0], pf[1] - origin[k][1], pf[2] - origin[k][2]};
            const double *z = zaxis[k];
            J6[0 * 3 + k] = z[1] * d[2] - z[2] * d[1];
            J6[1 * 3 + k] = z[2] * d[0] - z[0] * d[2];
            J6[2 * 3 + k] = z[0] * d[1] - z[1] * d[0];
            for (int i = 0; i < 3; ++i) J6[(3 + i) * 3 + k] = z[i];
        }
        /* iMd = oMf^-1 oMdes */
        double Ri[9], pi[3], e6[6];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Ri[3 * i + j] = Rf[3 * j + i];
            pi[i] = Rf[i] * (target[0] - pf[0]) + Rf[3 + i] * (target[1] - pf[1]) + Rf[6 + i] * (target[2] - pf[2]);
        }
        orc_se3_log(Ri, pi, e6);
        en = 0.0;
        for (int i = 0; i < 6; ++i) en += e6[i] * e6[i];
        en = sqrt(en);
        if (en < p->ik_tol) break;
        /* J = -Jlog6(iMd^-1) J_frame ; iMd^-1 = (Rf, -Rf pi) */
        double pinv[3], Jl[36], Jt[18];
        for (int i = 0; i < 3; ++i) pinv[i] = -(Rf[3 * i] * pi[0] + Rf[3 * i + 1] * pi[1] + Rf[3 * i + 2] * pi[2]);
        orc_se3_jlog(Rf, pinv, Jl);
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 3; ++k) {
            double s = 0; for (int m2 = 0; m2 < 6; ++m2) s += Jl[6 * i + m2] * J6[3 * m2 + k];
            Jt[3 * i + k] = -s;
        }
        double JJt[36], y[6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
            double s = 0; for (int k = 0; k < 3; ++k) s += Jt[3 * i + k] * Jt[3 * j + k];
            JJt[6 * i + j] = s + (i == j ? p->ik_damp : 0.0);
        }
        memcpy(y, e6, sizeof(y));
        gauss_solve(6, JJt, y, 1);
        for (int k = 0; k < 3; ++k) {
            double v = 0; for (int i = 0; i < 6; ++i) v += Jt[3 * i + k] * y[i];
            q[k] += -v * p->ik_dt;
        }
    }
    if (err) *err = en;
    return it;
}

int orc_swing_step(const orc_leg_model *m, const orc_swing_params *p, const double pos[3], const double quat[4],
                   const double q[6], const double des_v[3], int iter, double q_cmd[6], double feet[6], double next_foot[3],
                   double *ik_err, int *ik_iters) {
    orc_gait_params g;
    g.dt = p->dt; g.mpc_step = 5; g.swing_time = p->swing_time; g.stance_time = p->stance_time;
    int ls, rs; double phase, remain;
    orc_calculate_gait(&g, iter, &ls, &rs, &phase, &remain);
    const int leg = (ls == 1) ? 0 : 1;
    double ft[6];
    for (int l = 0; l < 2; ++l) orc_leg_fk(m, l, pos, quat, q + 3 * l, ft + 3 * l, 0);
    double fin[3], nxt[3], qv[3], err;
    orc_foot_placement(p, pos, des_v, remain, ls, fin);
    orc_swing_next(p, ft + 3 * leg, fin, remain, nxt);
    for (int k = 0; k < 3; ++k) qv[k] = q[3 * leg + k];
    int its = p->ik_mode == 1 ? orc_leg_ik6(m, p, leg, pos, quat, nxt, qv, &err) : orc_leg_ik(m, p, leg, pos, quat, nxt, qv, &err);
    for (int k = 0; k < 3; ++k) q_cmd[3 * leg + k] = qv[k];
    if (feet) memcpy(feet, ft, sizeof(ft));
    if (next_foot) memcpy(next_foot, nxt, sizeof(nxt));
    if (ik_err) *ik_err = err;
    if (ik_iters) *ik_iters = its;
    return leg;
}

void orc_grf_to_torque(const orc_leg_model *m, const double quat[4], const double q[6], const double u0[6], double tau[6]) {
    const double zero[3] = {0, 0, 0};
    for (int l = 0; l < 2; ++l) {
        double p[3], J[9];
        orc_leg_fk(m, l, zero, quat, q + 3 * l, p, J);
        for (int k = 0; k < 3; ++k) tau[3 * l + k] = -(J[k] * u0[3 * l] + J[3 + k] * u0[3 * l + 1] + J[6 + k] * u0[3 * l + 2]);
    }
}
