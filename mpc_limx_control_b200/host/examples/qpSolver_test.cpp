// qpSolver_test.cpp -- the reference's only runnable hot-path program (reference
// src/qpSolver_test.cpp: 4-state cart, N = 15, Ts = 0.01, circular reference, 500 closed-loop steps)
// driven through the facade, i.e. every discretise / build / solve / update runs on the B200.
// Deviations from the reference driver, both deliberate: the spurious A_eq
// block is not stacked, and the solver state starts at the demo's (2,0,0,0) instead of zero.
// Output: one line per step  "k u0 u1 x0 x1 x2 x3 err"  (machine readable), summary on stderr.
#include <chrono>
#include <cmath>
#include <cstdio>

#include "../QPSolver.h"

using namespace mpcb200::host;

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 500;
    const double Ts = 0.01;
    const int N = 15;
    MatrixXd Ac = MatrixXd::FromRows(4, 4, {0, 1, 0, 0, 0, -0.1, 0, 0, 0, 0, 0, 1, 0, 0, 0, -0.1});
    MatrixXd Bc = MatrixXd::FromRows(4, 2, {0, 0, 5, 0, 0, 0, 0, 5});
    MatrixXd Q = VectorXd({50, 5, 50, 5}).asDiagonal();
    MatrixXd R = 0.1 * MatrixXd::Identity(2, 2);
    MatrixXd P = 20 * Q;
    VectorXd x_min({-5, -3, -5, -3});
    VectorXd x_max = -x_min;
    try {
        QPSolver qpSolver(Ts, N, Ac, Bc, Q, R, P, x_min, x_max, -8.0, 8.0);
        VectorXd xi({2, 0, 0, 0});
        qpSolver.setState(xi);
        const double radius = 2.0, angular_vel = 0.5;
        int uncertified = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < steps; ++k) {
            MatrixXd xi_ref(4, N + 1);
            for (int i = 0; i <= N; ++i) {
                double theta = angular_vel * (k * Ts + i * Ts);
                xi_ref(0, i) = radius * cos(theta);
                xi_ref(2, i) = radius * sin(theta);
                xi_ref(1, i) = -radius * angular_vel * sin(theta);
                xi_ref(3, i) = radius * angular_vel * cos(theta);
            }
            MatrixXd H, A_eq, A_ineq, U_opt;
            VectorXd f, b_eq, lb, ub, lbA_ineq, ubA_ineq;
            qpSolver.buildQPParams(xi, xi_ref, H, f, A_eq, b_eq, lb, ub, A_ineq, lbA_ineq, ubA_ineq);
            if (!qpSolver.solveQP(H, f, A_ineq, lb, ub, lbA_ineq, ubA_ineq, U_opt)) ++uncertified;
            VectorXd u({U_opt(0, 0), U_opt(1, 0)});
            qpSolver.updateState(u);
            xi = qpSolver.getState();
            double err = std::hypot(xi(0) - xi_ref(0, 0), xi(2) - xi_ref(2, 0));
            printf("%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k, u(0), u(1), xi(0), xi(1), xi(2), xi(3), err);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "qp_test: %d steps, %d not certified, %.3f ms/step (build + solve + update on device)\n", steps,
                uncertified, ms / steps);
        return uncertified ? 2 : 0;
    } catch (const DeviceError& e) {
        fprintf(stderr, "qp_test: %s\n", e.what());
        return 1;
    }
}
