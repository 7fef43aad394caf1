// stateEstimator.h -- headless host shim with the interface of the reference class stateEstimator
// (reference include/stateEstimator.h:85-182: updateJointStates / updateContact / updateImu / update), without
// OCS2, Pinocchio, ROS or realtime_tools.  The 12-state / 14-measurement linear Kalman filter itself
// (include/stateEstimator.h:217-310) runs on the device through the C ABI (mpc_b200_kf_update_host); the leg
// kinematics come from mpc_b200_leg_model (link offsets of include/MPCParam.h:13-38, joint axes as parameters).
// update() returns the RobotOdomState the reference fills at :319-333, so an instance can serve as the state
// source of the MPC shim (MPC::StateSource) in place of StateEstimatorFake.
#pragma once
#include <array>
#include <cstdint>
#include <string>

#include "../../include/mpc_b200.h"
#include "MPCController.h"

namespace mpcb200 {
namespace host {

class stateEstimator {
public:
    explicit stateEstimator(int device = 0) : device_(device) {
        mpc_b200_kf_default_params(&params);
        mpc_b200_leg_default_model(&leg_model);
        xHat_.fill(0.0);                                   // include/stateEstimator.h:191
        p_.fill(0.0);
        for (int i = 0; i < 12; ++i) p_[i * 12 + i] = 100.0;   // :206-207
    }

    mpc_b200_kf_params params;          // noise constants, include/stateEstimator.h:124-130
    mpc_b200_leg_model leg_model;

    // reference :339-343
    void updateJointStates(const std::array<double, 6>& jointPos, const std::array<double, 6>& jointVel) { q_ = jointPos; dq_ = jointVel; }
    // reference :94 (contact_flag_t)
    void updateContact(bool left, bool right) { contact_[0] = left ? 1 : 0; contact_[1] = right ? 1 : 0; }
    // reference :345-360 (the covariances are carried for the odometry message only and are not used by the filter)
    void updateImu(const std::array<double, 4>& quat_xyzw, const std::array<double, 3>& angularVelLocal,
                   const std::array<double, 3>& linearAccelLocal) {
        quat_ = quat_xyzw; gyro_ = angularVelLocal; accel_ = linearAccelLocal;
    }

    // reference :217-337: one filter update with period dt seconds
    RobotOdomState update(double dt) {
        double odom[13];
        const int rc = mpc_b200_kf_update_host(device_, &params, &leg_model, 1, dt, quat_.data(), gyro_.data(), accel_.data(), q_.data(),
                                               dq_.data(), contact_, xHat_.data(), p_.data(), odom);
        if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("mpc_b200_kf_update_host: ") + mpc_b200_strerror(rc));
        for (int i = 0; i < 3; ++i) { robotOdomState_.pos[i] = odom[i]; robotOdomState_.v_pos[i] = odom[7 + i]; robotOdomState_.v_ori[i] = odom[10 + i]; }
        for (int i = 0; i < 4; ++i) robotOdomState_.quat[i] = odom[3 + i];
        // roll, pitch, yaw for the MPC state (the reference's estimators fill `ori`, include/state_estimator_fake.h:53-67)
        const double x = quat_[0], y = quat_[1], z = quat_[2], w = quat_[3];
        robotOdomState_.ori[0] = std::atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
        double sp = 2.0 * (w * y - z * x);
        sp = sp > 1.0 ? 1.0 : (sp < -1.0 ? -1.0 : sp);
        robotOdomState_.ori[1] = std::asin(sp);
        robotOdomState_.ori[2] = std::atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
        return robotOdomState_;
    }

    const std::array<double, 12>& state() const { return xHat_; }         // base position, velocity, foot positions (world)
    const std::array<double, 144>& covariance() const { return p_; }
    RobotOdomState robotOdomState_;

private:
    int device_;
    std::array<double, 12> xHat_;
    std::array<double, 144> p_;
    std::array<double, 6> q_{{0, 0, 0, 0, 0, 0}}, dq_{{0, 0, 0, 0, 0, 0}};
    std::array<double, 4> quat_{{0, 0, 0, 1}};
    std::array<double, 3> gyro_{{0, 0, 0}}, accel_{{0, 0, 9.81}};
    uint8_t contact_[2] = {1, 1};
};

}  // namespace host
}  // namespace mpcb200
