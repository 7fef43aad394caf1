// pinocchio_kinematics.h -- host shim with the interface of the reference class PinocchioKinematics
// (reference include/pinocchio_kinematics.h:14-157): forwardKinematics / getLinkPosition / inverseKinematics /
// setBaseLinkPose, so that the reference's call sites (include/mpcQP.h:125-137, include/MPCController.h:134-175) keep
// their shape.  Bodies call the C ABI only (batched kernels of csrc/leg_b200.cu with B = 1).
//
// Pinocchio and the URDF the reference loads (:24, external repository) are absent: the kinematic model is the
// mpc_b200_leg_model parameter struct (link offsets of include/MPCParam.h:13-38; joint axes are parameters).
// Differences from the reference class, by necessity: no `model` / `data` members, no inverseDynamics /
// printJointPositions (not on the hot path), nothing is printed (the reference prints per iteration, :96-99,135-146).
#pragma once
#include <array>
#include <cstdio>
#include <string>

#include "../../include/mpc_b200.h"
#include "QPSolver.h"

namespace mpcb200 {
namespace host {

struct Vector3d {
    double v[3] = {0, 0, 0};
    Vector3d() {}
    Vector3d(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
    static Vector3d Zero() { return Vector3d(); }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
struct Vector4d {
    double v[4] = {0, 0, 0, 1};
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
// Eigen::Quaterniond stand-in: constructor order (w, x, y, z) as in Eigen (include/pinocchio_kinematics.h:153)
struct Quaterniond {
    double w_ = 1, x_ = 0, y_ = 0, z_ = 0;
    Quaterniond() {}
    Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
    double w() const { return w_; }
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
};

class PinocchioKinematics {
public:
    mpc_b200_leg_model model;          // stands in for pinocchio::Model (:17)
    mpc_b200_swing_params ik_params;   // IK constants of :74-77; ik_params.ik_mode = 1 selects the 6-D task as written

    explicit PinocchioKinematics(int device = 0) : device_(device) {
        mpc_b200_leg_default_model(&model);
        mpc_b200_swing_default_params(&ik_params);
    }

    // :30-33  joint configuration -> placements of contact_L_Link / contact_R_Link (at the base pose set last)
    void forwardKinematics(const VectorXd& q) {
        double q6[6];
        for (int i = 0; i < 6; ++i) q6[i] = i < q.size() ? q(i) : 0.0;
        const int rc = mpc_b200_leg_fk_host(device_, &model, 1, base_pos_, base_quat_, q6, feet_.data(), nullptr);
        if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("PinocchioKinematics::forwardKinematics: ") + mpc_b200_strerror(rc));
    }

    // :36-43  unknown frame -> message on stderr and the zero vector, as the reference does
    Vector3d getLinkPosition(const std::string& link_name) const {
        const int leg = leg_of(link_name);
        if (leg < 0) {
            std::fprintf(stderr, "Cannot find frame: %s in the model.\n", link_name.c_str());
            return Vector3d::Zero();
        }
        return Vector3d(feet_[3 * leg], feet_[3 * leg + 1], feet_[3 * leg + 2]);
    }

    // :61-149  damped least squares from the initial guess; tolerance / iteration cap are arguments as in the reference
    VectorXd inverseKinematics(const std::string& link_name, const Vector3d& target_position, const VectorXd& initial_guess,
                               double tolerance = 1e-3, int max_iterations = 10) {
        VectorXd q = initial_guess;
        const int leg = leg_of(link_name);
        if (leg < 0) {
            std::fprintf(stderr, "Cannot find frame: %s in the model.\n", link_name.c_str());
            return q;                                                          // :68-71
        }
        mpc_b200_swing_params p = ik_params;
        p.ik_tol = tolerance; p.ik_max_iter = max_iterations;
        double q_in[6], q_out[6];
        for (int i = 0; i < 6; ++i) q_in[i] = i < q.size() ? q(i) : 0.0;
        const int32_t l = leg;
        int32_t its = 0;
        const int rc = mpc_b200_leg_ik_host(device_, &model, &p, 1, base_pos_, base_quat_, &l, target_position.v, q_in, q_out, &last_err_, &its);
        if (rc != MPC_B200_OK) throw DeviceError(rc, std::string("PinocchioKinematics::inverseKinematics: ") + mpc_b200_strerror(rc));
        last_iters_ = its;
        if (q.size() < 6) q.resize(6);
        for (int i = 0; i < 6; ++i) q(i) = q_out[i];
        return q;
    }

    // :153-157  (the reference writes data.oMi[1]; its fixed-base model then ignores it in forwardKinematics -- here the
    // base pose is honoured, which is what the call sites assume)
    void setBaseLinkPose(const Vector3d& position, const Quaterniond& orientation) {
        for (int i = 0; i < 3; ++i) base_pos_[i] = position(i);
        base_quat_[0] = orientation.x(); base_quat_[1] = orientation.y(); base_quat_[2] = orientation.z(); base_quat_[3] = orientation.w();
    }

    double lastIkError() const { return last_err_; }
    int lastIkIterations() const { return last_iters_; }

private:
    static int leg_of(const std::string& n) { return n == "contact_L_Link" ? 0 : (n == "contact_R_Link" ? 1 : -1); }
    int device_ = 0;
    double base_pos_[3] = {0, 0, 0}, base_quat_[4] = {0, 0, 0, 1};   // [x, y, z, w]
    std::array<double, 6> feet_{{0, 0, 0, 0, 0, 0}};
    double last_err_ = 0.0;
    int last_iters_ = 0;
};

}  // namespace host
}  // namespace mpcb200
