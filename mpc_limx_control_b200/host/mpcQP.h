// mpcQP.h -- host facade of the reference class mpcQP (reference include/mpcQP.h:8-33,35-182):
// TRON1 problem setup (weights, x0, reference trajectory), model linearisation and the QP solve,
// with the setters/getter BASELINE.json's north_star asks for (set state, reference, gait/contact
// schedule, foot positions; get optimal ground-reaction forces) and a batch entry point.
//
// Bodies call the C ABI only.  The physics is the *intended* single-rigid-body model (yaw rotation,
// world inertia, signed skew, 1/m, two feet -> 6 forces); the reference-as-written placeholders
// (include/mpcQP.h:142-181) are available through QPSolver with mpcQP::literalModel().
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mpc_b200.h"
#include "QPSolver.h"
#include "limxsdk_stub.h"
#include "pinocchio_kinematics.h"

namespace mpcb200 {
namespace host {

class mpcQP {
public:
    // horizon in {10, 20, 50}; params default to the reference constants (include/mpcQP.h:18-22,54-56)
    explicit mpcQP(int horizon = 10, int max_batch = 1, int device = 0, const mpc_b200_tron1_params* params = nullptr)
        : N(horizon), eng(nullptr) {
        if (params) prm = *params; else mpc_b200_tron1_default_params(&prm);
        int rc = mpc_b200_create(&prm, horizon, max_batch, device, &eng);
        if (rc) throw DeviceError(rc, std::string("mpcQP: ") + mpc_b200_strerror(rc));
        x0.assign(13, 0.0); x0[12] = -9.8;
        x_ref.assign((size_t)13 * (N + 1), 0.0);
        feet.assign(6, 0.0);
        contact.assign((size_t)2 * N, 1);
        U_opt.assign((size_t)6 * N, 0.0);
    }
    // THE REFERENCE'S CONSTRUCTOR (include/mpcQP.h:10,35-119), same argument list: joint state, base position / velocity /
    // rpy / angular velocity / quaternion, the kinematics object and `leg` = left_leg_state (0 = the left foot is the
    // support foot, :133-137).  As there: sets the base pose on the kinematics object, runs forward kinematics at state.q,
    // reads contact_L_Link / contact_R_Link (:125-137), builds x0 and the reference trajectory (:66-97, omega_yaw 0.1,
    // velocity_x 0.5), solves; u = U_opt.col(0) (:118) is then available from optimalForce().
    // The quaternion is read as the estimator stores it, [x, y, z, w] (include/state_estimator_fake.h:22).
    mpcQP(const limxsdk::RobotState& state, const Vector3d& currentPosition, const Vector3d& currentVelocity,
          const Vector3d& currentOrientation, const Vector3d& currentAngularVelocity, const Vector4d& currentQuat,
          PinocchioKinematics& kinematics, int leg, int horizon = 10, int device = 0)
        : mpcQP(horizon, 1, device) {
        kinematics.setBaseLinkPose(currentPosition, Quaterniond(currentQuat(3), currentQuat(0), currentQuat(1), currentQuat(2)));
        VectorXd jointPositions((int)state.q.size());
        for (int i = 0; i < (int)state.q.size(); ++i) jointPositions(i) = (double)state.q[i];
        kinematics.forwardKinematics(jointPositions);
        setState(currentPosition, currentVelocity, currentOrientation, currentAngularVelocity);
        setFootPositions(kinematics.getLinkPosition("contact_L_Link"), kinematics.getLinkPosition("contact_R_Link"));
        setReference(0.1, 0.5);
        std::vector<uint8_t> c((size_t)2 * N);
        for (int k = 0; k < N; ++k) { c[2 * k] = (leg == 0); c[2 * k + 1] = (leg != 0); }
        setContactSchedule(c);
        solve();
    }
    // the same with the feet given directly (no kinematics object)
    mpcQP(const Vector3d& currentPosition, const Vector3d& currentVelocity, const Vector3d& currentOrientation,
          const Vector3d& currentAngularVelocity, const Vector3d& leftFoot, const Vector3d& rightFoot, int leg,
          int horizon = 10, int device = 0)
        : mpcQP(horizon, 1, device) {
        setState(currentPosition, currentVelocity, currentOrientation, currentAngularVelocity);
        setFootPositions(leftFoot, rightFoot);
        setReference(0.1, 0.5);                     // omega_yaw, velocity_x (include/mpcQP.h:75-76)
        std::vector<uint8_t> c((size_t)2 * N);
        for (int k = 0; k < N; ++k) { c[2 * k] = (leg == 0); c[2 * k + 1] = (leg != 0); }
        setContactSchedule(c);
        solve();
    }
    ~mpcQP() { if (eng) mpc_b200_destroy(eng); }
    mpcQP(const mpcQP&) = delete;
    mpcQP& operator=(const mpcQP&) = delete;

    // x0 = [rpy, pos, omega, vel, -9.8]  (include/mpcQP.h:66-71)
    void setState(const Vector3d& pos, const Vector3d& vel, const Vector3d& rpy, const Vector3d& omega) {
        for (int i = 0; i < 3; ++i) { x0[i] = rpy(i); x0[3 + i] = pos(i); x0[6 + i] = omega(i); x0[9 + i] = vel(i); }
        x0[12] = -9.8;
    }
    // reference generator of include/mpcQP.h:74-97
    void setReference(double omega_yaw, double velocity_x) {
        const double Ts = prm.Ts;
        for (int i = 0; i <= N; ++i) {
            double t = i * Ts;
            double* c = &x_ref[(size_t)13 * i];
            for (int j = 0; j < 13; ++j) c[j] = x0[j];
            c[2] = x0[2] + t * omega_yaw;
            c[3] = x0[3] + t * velocity_x;
            c[9] = (i == 0) ? x0[9] : velocity_x;
            c[12] = -9.8;
        }
    }
    void setReference(const MatrixXd& xi_ref) {   // 13 x (N+1)
        if (xi_ref.rows() != 13 || xi_ref.cols() != N + 1) throw DeviceError(MPC_B200_EINVAL, "setReference: need 13 x (N+1)");
        x_ref.assign(xi_ref.data(), xi_ref.data() + (size_t)13 * (N + 1));
    }
    void setFootPositions(const Vector3d& left, const Vector3d& right) {
        for (int i = 0; i < 3; ++i) { feet[i] = left(i); feet[3 + i] = right(i); }
    }
    void setContactSchedule(const std::vector<uint8_t>& c) {
        if ((int)c.size() != 2 * N) throw DeviceError(MPC_B200_EINVAL, "setContactSchedule: need N x 2");
        contact = c;
        use_iter = false;   // an explicit schedule replaces a gait clock set earlier
    }
    // schedule from the gait clock (MPC::calculateGait at iter + k*mpcStep); iter < 0 = standing
    void setGaitIteration(int iter) { gait_iter = iter; use_iter = true; }

    bool solve() {
        // this facade holds ONE pair of foot positions (the reference computes one FK result per solve,
        // include/mpcQP.h:125-137); per-step feet are a batch-entry feature (solveBatch, caller-owned [B][N][2][3])
        if (prm.per_step_feet && prm.ltv) throw DeviceError(MPC_B200_EINVAL, "mpcQP::solve: per_step_feet engines take their feet through solveBatch");
        int32_t st = 2, it = 0;
        int rc = mpc_b200_tron1_solve_host(eng, 1, x0.data(), x_ref.data(), feet.data(), use_iter ? nullptr : contact.data(),
                                           use_iter ? &gait_iter : nullptr, U_opt.data(), &st, &it);
        if (rc) throw DeviceError(rc, std::string("mpcQP::solve: ") + mpc_b200_strerror(rc) + " (" + mpc_b200_last_error(eng) + ")");
        status = st; iterations = it;
        return st == 0;
    }
    // u = U_opt.col(0): required ground-reaction forces now (include/mpcQP.h:118) [fLx,fLy,fLz,fRx,fRy,fRz]
    std::array<double, 6> optimalForce() const { return {U_opt[0], U_opt[1], U_opt[2], U_opt[3], U_opt[4], U_opt[5]}; }
    const std::vector<double>& forces() const { return U_opt; }   // 6 x N column-major
    int lastStatus() const { return status; }
    int lastIterations() const { return iterations; }

    // batch entry point (north_star): B independent robots, host buffers laid out as in include/mpc_b200.h
    bool solveBatch(int B, const double* x0s, const double* x_refs, const double* feets, const uint8_t* contacts,
                    const int32_t* iters_in, double* forces_out, int32_t* status_out, int32_t* iters_out) {
        int rc = mpc_b200_tron1_solve_host(eng, B, x0s, x_refs, feets, contacts, iters_in, forces_out, status_out, iters_out);
        if (rc) throw DeviceError(rc, std::string("mpcQP::solveBatch: ") + mpc_b200_strerror(rc) + " (" + mpc_b200_last_error(eng) + ")");
        return true;
    }

    // asynchronous batch entry for queues of independent batches (pinned buffers; up to six batches in flight): queue with
    // solveBatchAsync, collect with waitAll
    void solveBatchAsync(int B, const double* x0s, const double* x_refs, const double* feets, const uint8_t* contacts,
                         const int32_t* iters_in, double* forces_out, int32_t* status_out, int32_t* iters_out) {
        int rc = mpc_b200_tron1_solve_host_async(eng, B, x0s, x_refs, feets, contacts, iters_in, forces_out, status_out, iters_out);
        if (rc) throw DeviceError(rc, std::string("mpcQP::solveBatchAsync: ") + mpc_b200_strerror(rc) + " (" + mpc_b200_last_error(eng) + ")");
    }
    void waitAll() {
        int rc = mpc_b200_wait(eng);
        if (rc) throw DeviceError(rc, std::string("mpcQP::waitAll: ") + mpc_b200_strerror(rc) + " (" + mpc_b200_last_error(eng) + ")");
    }

    // reference-literal continuous model (include/mpcQP.h:139-181): Ac 13x13, Bc 13x3, one support foot
    static void literalModel(const Vector3d& pos, const Vector3d& foot, double mass, MatrixXd& Ac, MatrixXd& Bc) {
        const double dx = foot(0) - pos(0), dy = foot(1) - pos(1), dz = foot(2) - pos(2);
        Ac = MatrixXd::Zero(13, 13); Bc = MatrixXd::Zero(13, 3);
        Ac(0, 7) = dz; Ac(0, 8) = dy; Ac(1, 6) = dz; Ac(1, 8) = dx; Ac(2, 6) = dy; Ac(2, 7) = dx;
        Ac(3, 9) = 1; Ac(4, 10) = 1; Ac(5, 11) = 1; Ac(11, 12) = -1;
        Bc(9, 0) = -mass; Bc(10, 1) = -mass; Bc(11, 2) = -mass;
    }

    mpc_b200_tron1_params prm;

private:
    int N;
    mpc_b200_engine* eng;
    std::vector<double> x0, x_ref, feet, U_opt;
    std::vector<uint8_t> contact;
    int32_t gait_iter = -1;
    bool use_iter = false;
    int status = 0, iterations = 0;
};

}  // namespace host
}  // namespace mpcb200
