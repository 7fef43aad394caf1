// MPCController.h -- headless host shim with the interface of the reference class MPC
// (reference include/MPCController.h:9-39,183-196): run(state, imu, cmd, iter) with POD stand-ins
// for the limxsdk types and an injectable state source instead of the ROS StateEstimatorFake
// (include/state_estimator_fake.h:118-143).
//
// run() performs the reference's four calls (include/MPCController.h:183-196) plus the force MPC the reference
// left as an empty stub:
//   update_odom_state (a1) -> calculateGait (a2) -> computeFootPlacement -> computeSwingFootDesiredPosition
//   (FK + swing profile + IK on the device, csrc/leg_b200.cu; writes the swing leg's cmd.q) ->
//   computeSupportFootForce (include/MPCController.h:178-180: force MPC on the device, then tau = -J'f into the
//   stance leg's cmd.tau).
// The leg model's joint axes are parameters (external URDF, see include/mpc_b200.h).
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <string>

#include "MPCParam.h"
#include "limxsdk_stub.h"
#include "mpcQP.h"

namespace mpcb200 {
namespace host {

struct RobotOdomState {   // reference include/state_estimator_fake.h:19-25
    double pos[3] = {0, 0, 0};
    double ori[3] = {0, 0, 0};
    double quat[4] = {0, 0, 0, 1};   // [x, y, z, w]
    double v_pos[3] = {0, 0, 0};
    double v_ori[3] = {0, 0, 0};
};

class MPC {
public:
    using StateSource = std::function<RobotOdomState()>;
    using FootSource = std::function<void(const limxsdk::RobotState&, const RobotOdomState&, Vector3d& left, Vector3d& right)>;

    explicit MPC(StateSource src = StateSource(), int horizon = 10, int device = 0)
        : estimates(src ? src : StateSource([] { return RobotOdomState(); })), qp(horizon, 1, device), device_(device) {
        mpc_b200_leg_default_model(&leg_model);
        mpc_b200_swing_default_params(&swing_params);
        for (int i = 0; i < 3; ++i) finalPosition(i) = 0.0;
        odom_state = estimates();
    }

    // reference include/MPCController.h:183-196
    void run(limxsdk::RobotState state, limxsdk::ImuData imu, limxsdk::RobotCmd& cmd, int iter) {
        (void)imu;
        update_odom_state();
        calculateGait(iter);
        if (enable_leg_pipeline && iter >= 0) {
            computeFootPlacement(finalPosition);
            computeSwingFootDesiredPosition(state, cmd, iter);
        }
        computeSupportFootForce(state, iter);
        if (enable_leg_pipeline) supportForceToTorque(state, cmd);
    }

    MPCParam param;
    mpc_b200_leg_model leg_model;        // link offsets of include/MPCParam.h:13-38, joint axes (parameters)
    mpc_b200_swing_params swing_params;  // gait / swing / IK constants of the reference
    bool enable_leg_pipeline = true;     // false: force MPC only (BASELINE config 1b latency measurement)
    bool kinematic_feet = false;         // true: the MPC's foot positions come from FK at state.q instead of the nominal offsets
    StateSource estimates;
    FootSource feet_from_kinematics;   // FK hook; default = nominal offsets under the base
    Vector3d desieredV_pos = make3(1.0, 0.0, 0.0);   // reference include/MPCController.h:16-17
    Vector3d desieredV_ori = make3(0.0, 0.0, 0.0);

    // results / introspection
    int leftLegState() const { return left_leg_state; }
    int rightLegState() const { return right_leg_state; }
    double gaitPhase() const { return phase; }
    double remainingSwingTime() const { return remainSwingTime; }
    std::array<double, 6> supportFootForce() const { return qp.optimalForce(); }
    std::array<double, 6> jointTorque() const { return tau_cmd; }            // tau = -J' f, [left 3, right 3]
    std::array<double, 3> swingFootLanding() const { return {finalPosition(0), finalPosition(1), finalPosition(2)}; }
    std::array<double, 3> swingFootNext() const { return next_foot; }
    std::array<double, 6> footPositions() const { return fk_feet; }          // FK of both contact points at state.q
    double ikError() const { return ik_err; }
    int ikIterations() const { return ik_iters; }
    bool lastSolveCertified() const { return qp.lastStatus() == 0; }

    // reference include/MPCController.h:61-75, float members of MPCParam widened exactly as there
    void calculateGait(int iter) {
        double currentTime = iter * param.dt;
        double cycleTime = param.swing_time + param.stance_time;
        phase = std::fmod(currentTime, cycleTime);
        if (phase < param.swing_time) { left_leg_state = 1; right_leg_state = 0; remainSwingTime = param.swing_time - phase; }
        else { left_leg_state = 0; right_leg_state = 1; remainSwingTime = cycleTime - phase; }
    }

    // reference include/MPCController.h:106-132 (host restatement; the device step recomputes it per robot)
    void computeFootPlacement(Vector3d& fin) {
        double px = currentPosition(0) + desieredV_pos(0) * remainSwingTime;
        double py = currentPosition(1) + desieredV_pos(1) * remainSwingTime;
        const double p_rel_max = 0.3;
        double pfx_rel = desieredV_pos(0) * 0.5 * param.stance_time;
        double pfy_rel = desieredV_pos(1) * 0.5 * param.stance_time;
        pfx_rel = std::fmin(std::fmax(pfx_rel, -p_rel_max), p_rel_max);
        pfy_rel = std::fmin(std::fmax(pfy_rel, -p_rel_max), p_rel_max);
        px += pfx_rel; py += pfy_rel;
        const auto& off = (left_leg_state == 1) ? param.static_foot_offset_left : param.static_foot_offset_right;
        fin(0) = px + off[0]; fin(1) = py + off[1]; fin(2) = 0.0;
    }

private:
    static Vector3d make3(double a, double b, double c) { Vector3d v; v(0) = a; v(1) = b; v(2) = c; return v; }

    void joint_angles(const limxsdk::RobotState& state, double q[6]) const {
        for (int i = 0; i < 6; ++i) q[i] = i < (int)state.q.size() ? (double)state.q[i] : 0.0;
    }

    // reference include/MPCController.h:134-175: FK of the swing foot, interpolation towards the landing point,
    // sine height profile, IK, and the write of the swing leg's joint targets into cmd.q
    void computeSwingFootDesiredPosition(const limxsdk::RobotState& state, limxsdk::RobotCmd& cmd, int iter) {
        double q[6], q_cmd[6], des_v[3] = {desieredV_pos(0), desieredV_pos(1), desieredV_pos(2)};
        joint_angles(state, q);
        for (int i = 0; i < 6; ++i) q_cmd[i] = i < (int)cmd.q.size() ? (double)cmd.q[i] : 0.0;
        int32_t it = iter, leg = 0, its = 0;
        const int rc = mpc_b200_swing_step_host(device_, &leg_model, &swing_params, 1, odom_state.pos, odom_state.quat, q, des_v, &it,
                                                q_cmd, fk_feet.data(), next_foot.data(), &leg, &ik_err, &its);
        if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("mpc_b200_swing_step_host: ") + mpc_b200_strerror(rc));
        ik_iters = its;
        have_fk = true;
        if (cmd.q.size() < 6) cmd.q.resize(6, 0.f);
        for (int i = 3 * leg; i < 3 * leg + 3; ++i) cmd.q[i] = (float)q_cmd[i];
    }

    // tau = -J' f of the stance leg(s) into cmd.tau (the force half of the reference's empty stub)
    void supportForceToTorque(const limxsdk::RobotState& state, limxsdk::RobotCmd& cmd) {
        double q[6];
        joint_angles(state, q);
        const auto f = qp.optimalForce();
        const int rc = mpc_b200_grf_to_torque_host(device_, &leg_model, 1, odom_state.quat, q, f.data(), tau_cmd.data());
        if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("mpc_b200_grf_to_torque_host: ") + mpc_b200_strerror(rc));
        if (cmd.tau.size() < 6) cmd.tau.resize(6, 0.f);
        for (int i = 0; i < 6; ++i) cmd.tau[i] = (float)tau_cmd[i];
    }

    // reference include/MPCController.h:45-58
    void update_odom_state() {
        odom_state = estimates();
        for (int i = 0; i < 3; ++i) {
            currentPosition(i) = odom_state.pos[i];
            currentVelocity(i) = odom_state.v_pos[i];
            currentOrientation(i) = odom_state.ori[i];
            currentAngularVelocity(i) = odom_state.v_ori[i];
        }
        for (int i = 0; i < 4; ++i) currentQuat(i) = odom_state.quat[i];
    }

    // body of the reference's empty stub (include/MPCController.h:178-180)
    void computeSupportFootForce(const limxsdk::RobotState& state, int iter) {
        Vector3d left, right;
        if (feet_from_kinematics) feet_from_kinematics(state, odom_state, left, right);
        else if (kinematic_feet) {
            if (!have_fk) {
                double q[6];
                joint_angles(state, q);
                const int rc = mpc_b200_leg_fk_host(device_, &leg_model, 1, odom_state.pos, odom_state.quat, q, fk_feet.data(), nullptr);
                if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("mpc_b200_leg_fk_host: ") + mpc_b200_strerror(rc));
            }
            for (int i = 0; i < 3; ++i) { left(i) = fk_feet[i]; right(i) = fk_feet[3 + i]; }
        } else {
            const double c = std::cos(currentOrientation(2)), s = std::sin(currentOrientation(2));
            const auto& L = param.static_foot_offset_left; const auto& R = param.static_foot_offset_right;
            left(0) = currentPosition(0) + c * L[0] - s * L[1]; left(1) = currentPosition(1) + s * L[0] + c * L[1]; left(2) = currentPosition(2) + L[2];
            right(0) = currentPosition(0) + c * R[0] - s * R[1]; right(1) = currentPosition(1) + s * R[0] + c * R[1]; right(2) = currentPosition(2) + R[2];
        }
        qp.setState(currentPosition, currentVelocity, currentOrientation, currentAngularVelocity);
        qp.setFootPositions(left, right);
        qp.setReference(desieredV_ori(2), desieredV_pos(0));
        qp.setGaitIteration(iter);
        qp.solve();
        have_fk = false;
    }

    RobotOdomState odom_state;
    mpcQP qp;
    int device_ = 0;
    bool have_fk = false;
    Vector3d finalPosition;
    std::array<double, 3> next_foot{{0, 0, 0}};
    std::array<double, 6> fk_feet{{0, 0, 0, 0, 0, 0}}, tau_cmd{{0, 0, 0, 0, 0, 0}};
    double ik_err = 0.0;
    int ik_iters = 0;
    int left_leg_state = 0, right_leg_state = 0;
    double phase = 0, remainSwingTime = 0;
    Vector3d currentPosition, currentVelocity, currentOrientation, currentAngularVelocity;
    Vector4d currentQuat;
};

}  // namespace host
}  // namespace mpcb200
