// tron1_single.cpp -- BASELINE config 1b through the controller shim: one TRON1 instance, standing
// (iter < 0), nominal feet, hold reference; prints the first-step ground-reaction forces and the
// p50/p99 latency of MPC::run (host call -> forces on host).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>

#include "../MPCController.h"
#include "../stateEstimator.h"

using namespace mpcb200::host;

int main(int argc, char** argv) {
    const int calls = argc > 1 ? atoi(argv[1]) : 2000;
    try {
        RobotOdomState s;
        s.pos[2] = 0.81181;
        MPC mpc([&] { return s; });
        mpc.desieredV_pos(0) = 0.0;
        limxsdk::RobotState st; limxsdk::ImuData imu; limxsdk::RobotCmd cmd;
        mpc.run(st, imu, cmd, -1);
        auto f = mpc.supportFootForce();
        printf("forces %.17g %.17g %.17g %.17g %.17g %.17g certified %d\n", f[0], f[1], f[2], f[3], f[4], f[5],
               (int)mpc.lastSolveCertified());
        st.q = {0.0f, 0.4f, -0.8f, 0.0f, 0.4f, -0.8f};   // bent knees
        mpc.run(st, imu, cmd, 250);   // left swing / right stance per calculateGait
        f = mpc.supportFootForce();
        printf("gait250 left_state %d right_state %d forces %.17g %.17g %.17g %.17g %.17g %.17g\n", mpc.leftLegState(),
               mpc.rightLegState(), f[0], f[1], f[2], f[3], f[4], f[5]);
        auto tq = mpc.jointTorque(); auto nf = mpc.swingFootNext(); auto fk = mpc.footPositions();
        printf("gait250 tau %.17g %.17g %.17g %.17g %.17g %.17g\n", tq[0], tq[1], tq[2], tq[3], tq[4], tq[5]);
        printf("gait250 fk_feet %.17g %.17g %.17g %.17g %.17g %.17g\n", fk[0], fk[1], fk[2], fk[3], fk[4], fk[5]);
        printf("gait250 swing_next %.17g %.17g %.17g cmd_q %.9g %.9g %.9g %.9g %.9g %.9g ik_err %.6g ik_iters %d\n", nf[0], nf[1], nf[2],
               cmd.q[0], cmd.q[1], cmd.q[2], cmd.q[3], cmd.q[4], cmd.q[5], mpc.ikError(), mpc.ikIterations());
        {   // the reference's own call shape: PinocchioKinematics + mpcQP(state, pos, vel, rpy, omega, quat, kin, leg)
            PinocchioKinematics kin;
            Vector4d quat;   // [x, y, z, w] identity
            Vector3d pos(0.0, 0.0, 0.655), vel(0.2, 0.0, 0.0), rpy(0.0, 0.0, 0.0), om(0.0, 0.0, 0.0);
            mpcQP qp(st, pos, vel, rpy, om, quat, kin, /*left_leg_state=*/0);
            auto u = qp.optimalForce();
            Vector3d fl = kin.getLinkPosition("contact_L_Link"), fr = kin.getLinkPosition("contact_R_Link");
            printf("mpcQP_ctor feet %.12g %.12g %.12g %.12g %.12g %.12g u %.12g %.12g %.12g %.12g %.12g %.12g certified %d\n",
                   fl(0), fl(1), fl(2), fr(0), fr(1), fr(2), u[0], u[1], u[2], u[3], u[4], u[5], (int)(qp.lastStatus() == 0));
            VectorXd q0(6);
            for (int i = 0; i < 6; ++i) q0(i) = st.q[i];
            Vector3d target(fl(0) + 0.02, fl(1) - 0.01, fl(2) + 0.03);
            for (int mode = 0; mode < 2; ++mode) {
                kin.ik_params.ik_mode = mode;
                VectorXd qi = kin.inverseKinematics("contact_L_Link", target, q0);
                printf("ik_mode%d q %.12g %.12g %.12g err %.9g iters %d\n", mode, qi(0), qi(1), qi(2), kin.lastIkError(), kin.lastIkIterations());
            }
            Vector3d none = kin.getLinkPosition("no_such_link");
            printf("missing_frame %.1f %.1f %.1f\n", none(0), none(1), none(2));
        }
        st.q = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        mpc.enable_leg_pipeline = false;   // the latency figure below is the force MPC alone (BASELINE config 1b)
        // the Kalman state estimator as the controller's state source: a robot standing still on both feet
        {
            stateEstimator est;
            est.updateJointStates({0.0, 0.4, -0.8, 0.0, 0.4, -0.8}, {0, 0, 0, 0, 0, 0});
            est.updateContact(true, true);
            est.updateImu({0, 0, 0, 1}, {0, 0, 0}, {0, 0, 9.81});
            RobotOdomState o;
            for (int i = 0; i < 400; ++i) o = est.update(0.002);
            printf("estimator pos %.9g %.9g %.9g vel %.3g %.3g %.3g foot_L %.9g %.9g %.9g\n", o.pos[0], o.pos[1], o.pos[2], o.v_pos[0], o.v_pos[1],
                   o.v_pos[2], est.state()[6], est.state()[7], est.state()[8]);
            MPC mpc2([&] { return est.update(0.002); });
            mpc2.desieredV_pos(0) = 0.0;
            mpc2.enable_leg_pipeline = false;
            mpc2.run(st, imu, cmd, -1);
            auto f2 = mpc2.supportFootForce();
            printf("estimator-driven forces %.9g %.9g %.9g %.9g %.9g %.9g certified %d\n", f2[0], f2[1], f2[2], f2[3], f2[4], f2[5],
                   (int)mpc2.lastSolveCertified());
        }
        std::vector<double> us;
        for (int i = 0; i < calls + 100; ++i) {
            auto t0 = std::chrono::steady_clock::now();
            mpc.run(st, imu, cmd, -1);
            if (i >= 100) us.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
        }
        std::sort(us.begin(), us.end());
        if (!us.empty()) printf("latency_us p50 %.2f p99 %.2f calls %zu\n", us[us.size() / 2], us[(size_t)(us.size() * 0.99)], us.size());
        return 0;
    } catch (const DeviceError& e) {
        fprintf(stderr, "tron1_single: %s\n", e.what());
        return 1;
    }
}
