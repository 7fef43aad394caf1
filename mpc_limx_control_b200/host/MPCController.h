// MPCController.h -- headless host shim with the interface of the reference class MPC
// (reference include/MPCController.h:9-39,183-196): run(state, imu, cmd, iter) with POD stand-ins
// for the limxsdk types and an injectable state source instead of the ROS StateEstimatorFake
// (include/state_estimator_fake.h:118-143).
//
// In scope: update_odom_state (a1), calculateGait (a2) and
// computeSupportFootForce -- the reference's empty stub (include/MPCController.h:178-180) -- which
// here runs the force MPC on the device and keeps the optimal ground-reaction forces.
// Out of scope this round: foot placement and the swing-leg IK (section 8f rank 2).
#pragma once
#include <array>
#include <cmath>
#include <functional>

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
        : estimates(src ? src : StateSource([] { return RobotOdomState(); })), qp(horizon, 1, device) {
        odom_state = estimates();
    }

    // reference include/MPCController.h:183-196
    void run(limxsdk::RobotState state, limxsdk::ImuData imu, limxsdk::RobotCmd& cmd, int iter) {
        (void)imu; (void)cmd;
        update_odom_state();
        calculateGait(iter);
        computeSupportFootForce(state, iter);
    }

    MPCParam param;
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
    bool lastSolveCertified() const { return qp.lastStatus() == 0; }

    // reference include/MPCController.h:61-75, float members of MPCParam widened exactly as there
    void calculateGait(int iter) {
        double currentTime = iter * param.dt;
        double cycleTime = param.swing_time + param.stance_time;
        phase = std::fmod(currentTime, cycleTime);
        if (phase < param.swing_time) { left_leg_state = 1; right_leg_state = 0; remainSwingTime = param.swing_time - phase; }
        else { left_leg_state = 0; right_leg_state = 1; remainSwingTime = cycleTime - phase; }
    }

private:
    static Vector3d make3(double a, double b, double c) { Vector3d v; v(0) = a; v(1) = b; v(2) = c; return v; }

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
        else {
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
    }

    RobotOdomState odom_state;
    mpcQP qp;
    int left_leg_state = 0, right_leg_state = 0;
    double phase = 0, remainSwingTime = 0;
    Vector3d currentPosition, currentVelocity, currentOrientation, currentAngularVelocity;
    Vector4d currentQuat;
};

}  // namespace host
}  // namespace mpcb200
