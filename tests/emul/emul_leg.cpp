// Host build of csrc/leg_core.cuh (TEST INFRASTRUCTURE ONLY): the CPU tier checks the kernel mathematics of the
// leg kinematics against the oracle without a GPU.
#include <cstring>

#include "../../include/mpc_b200.h"
#include "../../mpc_limx_control_b200/csrc/leg_core.cuh"

using namespace mpcb200;

static LegModel to_model(const mpc_b200_leg_model* m) {
    LegModel M;
    memcpy(M.offset, m->offset, sizeof(M.offset));
    memcpy(M.axis, m->axis, sizeof(M.axis));
    return M;
}
static SwingParams to_swing(const mpc_b200_swing_params* p) {
    SwingParams S;
    S.dt = p->dt; S.swing_time = p->swing_time; S.stance_time = p->stance_time; S.gait_height = p->gait_height;
    S.p_rel_max = p->p_rel_max;
    for (int i = 0; i < 3; ++i) { S.foot_off_l[i] = p->foot_offset_left[i]; S.foot_off_r[i] = p->foot_offset_right[i]; }
    S.ik_tol = p->ik_tol; S.ik_dt = p->ik_dt; S.ik_damp = p->ik_damp; S.ik_max_iter = p->ik_max_iter;
    S.ik_mode = p->ik_mode;
    return S;
}

extern "C" {

void emul_leg_fk(const mpc_b200_leg_model* m, const double* pos, const double* quat, const double* q, double* feet, double* jac) {
    LegModel M = to_model(m);
    double Rb[9];
    quat_to_rot(quat, Rb);
    for (int leg = 0; leg < 2; ++leg) leg_fk_world(M, leg, pos, Rb, q + 3 * leg, feet + 3 * leg, jac ? jac + 9 * leg : nullptr);
}

int emul_swing_step(const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, const double* pos, const double* quat, const double* q,
                    const double* des_vel, int iter, double* q_cmd, double* feet, double* next_foot, double* ik_err, int* ik_iters) {
    LegModel M = to_model(m);
    SwingParams P = to_swing(p);
    double Rb[9], fin[3], nxt[3], qv[3], phase, remain, err, ft[6];
    int ls, rs;
    quat_to_rot(quat, Rb);
    gait_state(P, iter, ls, rs, phase, remain);
    const int leg = (ls == 1) ? 0 : 1;
    for (int l = 0; l < 2; ++l) leg_fk_world(M, l, pos, Rb, q + 3 * l, ft + 3 * l, nullptr);
    foot_placement(P, pos, des_vel, remain, ls, fin);
    swing_next_position(P, ft + 3 * leg, fin, remain, nxt);
    for (int k = 0; k < 3; ++k) qv[k] = q[3 * leg + k];
    const int its = leg_ik_task(M, P, leg, pos, Rb, nxt, qv, err);
    for (int k = 0; k < 3; ++k) q_cmd[3 * leg + k] = qv[k];
    if (feet) memcpy(feet, ft, sizeof(ft));
    if (next_foot) memcpy(next_foot, nxt, sizeof(nxt));
    if (ik_err) *ik_err = err;
    if (ik_iters) *ik_iters = its;
    return leg;
}

int emul_leg_ik(const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, int leg, const double* pos, const double* quat,
                const double* target, double* q3, double* err) {
    LegModel M = to_model(m);
    SwingParams P = to_swing(p);
    double Rb[9], e = 0.0;
    quat_to_rot(quat, Rb);
    const int its = leg_ik_task(M, P, leg, pos, Rb, target, q3, e);
    if (err) *err = e;
    return its;
}
void emul_se3_log(const double* R, const double* t, double* xi) { se3_log(R, t, xi); }
void emul_se3_jlog(const double* R, const double* t, double* Jl) { se3_jlog(R, t, Jl); }

void emul_grf_to_torque(const mpc_b200_leg_model* m, const double* quat, const double* q, const double* u0, double* tau) {
    LegModel M = to_model(m);
    double Rb[9], p[3], J[9];
    const double zero[3] = {0, 0, 0};
    quat_to_rot(quat, Rb);
    for (int leg = 0; leg < 2; ++leg) {
        leg_fk_world(M, leg, zero, Rb, q + 3 * leg, p, J);
        grf_to_torque(J, u0 + 3 * leg, tau + 3 * leg);
    }
}

}  // extern "C"

// ---- Kalman filter core (csrc/kf_core.cuh) with a one-thread group --------------------------------------------------
#include "../../mpc_limx_control_b200/csrc/kf_core.cuh"
struct GrpSerialKf {
    int tid() const { return 0; }
    int size() const { return 1; }
    void sync() const {}
};
extern "C" void emul_kf_update(const mpc_b200_kf_params* p, const mpc_b200_leg_model* m, double dt, const double* quat, const double* gyro,
                               const double* accel, const double* q, const double* dq, const uint8_t* contact, double* xhat, double* P,
                               double* odom) {
    KfParams K;
    K.foot_radius = p->foot_radius; K.imu_noise_pos = p->imu_process_noise_position; K.imu_noise_vel = p->imu_process_noise_velocity;
    K.foot_noise_pos = p->foot_process_noise_position; K.foot_sensor_pos = p->foot_sensor_noise_position;
    K.foot_sensor_vel = p->foot_sensor_noise_velocity; K.foot_height_noise = p->foot_height_sensor_noise;
    K.suspect = p->high_suspect_number; K.accel_transpose = p->accel_transpose;
    LegModel M = to_model(m);
    KfWork W;
    kf_update(K, M, dt, quat, gyro, accel, q, dq, contact, xhat, P, odom, W, GrpSerialKf());
}
