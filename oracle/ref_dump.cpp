// ref_dump -- drives the reference's OWN QPSolver (compiled unmodified from /root/reference/src/QPSolver.cpp against
// oracle/ref_shim/) and writes what it computes, as raw doubles, for tests/golden/make_ref_golden.py.
// TEST INFRASTRUCTURE: built into oracle/_ref/ (git-ignored), only ever run by the golden generator and by
// tests/test_ref_pin.py when /root/reference is present.  Built with -fno-access-control so that the private members
// Ad, Bd and xi (include/QPSolver.h:39-56) can be read and seeded without touching the reference header.
//
//   ref_dump cases <in.bin> <out.bin>   generic: one QPSolver per case (ctor -> discretizeSystem -> buildQPParams -> updateState)
//   ref_dump demo  <out.bin>            the src/qpSolver_test.cpp:6-50 scenario, 500 closed-loop steps
#include "QPSolver.h"
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

using Eigen::MatrixXd;
using Eigen::VectorXd;

static FILE *g_out;
static void put(const Eigen::Dense &m) { fwrite(m.data(), sizeof(double), (size_t)m.size(), g_out); }
static void put1(double x) { fwrite(&x, sizeof(double), 1, g_out); }

static MatrixXd get(FILE *f, int r, int c) {
    MatrixXd m(r, c);
    if (fread(m.data(), sizeof(double), (size_t)r * c, f) != (size_t)r * c) { fprintf(stderr, "ref_dump: short read\n"); exit(2); }
    return m;
}

static int run_cases(const char *in, const char *out) {
    FILE *f = fopen(in, "rb");
    g_out = fopen(out, "wb");
    if (!f || !g_out) return 2;
    int32_t ncases;
    if (fread(&ncases, 4, 1, f) != 1) return 2;
    for (int c = 0; c < ncases; ++c) {
        int32_t hd[3];
        double sc[3];
        if (fread(hd, 4, 3, f) != 3 || fread(sc, 8, 3, f) != 3) return 2;
        int NX = hd[0], NU = hd[1], N = hd[2];
        double Ts = sc[0], u_min = sc[1], u_max = sc[2];
        // the header types the state Vector4d / the input Vector2d (include/QPSolver.h:22,34,55); other sizes need the
        // shim's relaxed fixed-size mode (oracle/ref_shim/Eigen/Dense header comment)
        Eigen::shim_relax_fixed = !(NX == 4 && NU == 2);
        MatrixXd Ac = get(f, NX, NX), Bc = get(f, NX, NU), Q = get(f, NX, NX), R = get(f, NU, NU), P = get(f, NX, NX);
        VectorXd x_min = get(f, NX, 1), x_max = get(f, NX, 1), x0v = get(f, NX, 1);
        MatrixXd xi_ref = get(f, NX, N + 1);
        VectorXd uv = get(f, NU, 1);

        QPSolver qp(Ts, N, Ac, Bc, Q, R, P, x_min, x_max, u_min, u_max);   // runs discretizeSystem (QPSolver.cpp:18)
        Eigen::Vector4d xi0 = x0v;
        MatrixXd H, A_eq, A_ineq;
        VectorXd fv, b_eq, lb, ub, lbA, ubA;
        qp.buildQPParams(xi0, xi_ref, H, fv, A_eq, b_eq, lb, ub, A_ineq, lbA, ubA);
        qp.xi = xi0;                       // QPSolver.cpp:12 starts at zero; seed the state, then one updateState
        Eigen::Vector2d u = uv;
        qp.updateState(u);
        put(qp.Ad); put(qp.Bd); put(H); put(fv); put(A_eq); put(b_eq); put(lb); put(ub); put(A_ineq); put(lbA); put(ubA);
        put(qp.xi);
    }
    fclose(f); fclose(g_out);
    return 0;
}

// src/qpSolver_test.cpp:6-50 scenario.  The QP handed to the solver drops the equality block (qpSolver_test.cpp:58-63,
// SURVEY.md appendix B.1) and is read column-major (B.2) so that the loop closes; step 0 is additionally solved exactly
// as written (stacked equality block, row-major read) and its status recorded.
static int run_demo(const char *out) {
    g_out = fopen(out, "wb");
    if (!g_out) return 2;
    double Ts = 0.01;
    int N = 15;
    Eigen::Matrix4d Ac;
    Eigen::Matrix<double, 4, 2> Bc;
    Ac << 0, 1, 0, 0,
          0, -0.1, 0, 0,
          0, 0, 0, 1,
          0, 0, 0, -0.1;
    Bc << 0, 0,
          5, 0,
          0, 0,
          0, 5;
    Eigen::Matrix4d Q = (Eigen::Vector4d() << 50, 5, 50, 5).finished().asDiagonal();
    Eigen::Matrix2d R = 0.1 * Eigen::Matrix2d::Identity();
    Eigen::Matrix4d P = 20 * Q;
    Eigen::Vector4d x_min = (Eigen::Vector4d() << -5, -3, -5, -3).finished();
    Eigen::Vector4d x_max = -x_min;
    double u_min = -8.0, u_max = 8.0;
    QPSolver qpSolver(Ts, N, Ac, Bc, Q, R, P, x_min, x_max, u_min, u_max);
    Eigen::Vector4d xi(2, 0, 0, 0);
    qpSolver.xi = xi;
    double trajectory_radius = 2.0, angular_vel = 0.5;
    put(qpSolver.Ad); put(qpSolver.Bd);
    put(xi);
    for (int k = 0; k < 500; ++k) {
        Eigen::Matrix<double, 4, 16> xi_ref;
        for (int i = 0; i <= N; ++i) {
            double t = k * Ts + i * Ts;
            double theta = angular_vel * t;
            xi_ref(0, i) = trajectory_radius * cos(theta);
            xi_ref(2, i) = trajectory_radius * sin(theta);
            xi_ref(1, i) = -trajectory_radius * angular_vel * sin(theta);
            xi_ref(3, i) = trajectory_radius * angular_vel * cos(theta);
        }
        MatrixXd H, A_eq, A_ineq;
        VectorXd f, b_eq, lb, ub, lbA_ineq, ubA_ineq;
        qpSolver.buildQPParams(xi, xi_ref, H, f, A_eq, b_eq, lb, ub, A_ineq, lbA_ineq, ubA_ineq);
        double as_written_status = -1;
        if (k == 0) {
            MatrixXd A_total(A_eq.rows() + A_ineq.rows(), 2 * N);
            A_total << A_eq, A_ineq;
            VectorXd lbA_total(b_eq.size() + lbA_ineq.size());
            lbA_total << b_eq, lbA_ineq;
            VectorXd ubA_total(b_eq.size() + ubA_ineq.size());
            ubA_total << b_eq, ubA_ineq;
            Eigen::Matrix<double, 2, 15> U_bad;
            qpOASES::shim_A_is_colmajor = false;
            qpSolver.solveQP(H, f, A_total, lb, ub, lbA_total, ubA_total, U_bad);
            as_written_status = qpOASES::shim_last_status;
        }
        Eigen::Matrix<double, 2, 15> U_opt;
        qpOASES::shim_A_is_colmajor = true;
        qpSolver.solveQP(H, f, A_ineq, lb, ub, lbA_ineq, ubA_ineq, U_opt);
        Eigen::Vector2d u = U_opt.col(0);
        qpSolver.updateState(u);
        xi = qpSolver.getState();
        // per step: u (2), x (4), solver status, [step 0 only: as-written status]
        put(u); put(xi); put1(qpOASES::shim_last_status);
        if (k == 0) { put1(as_written_status); put(H); put(f); put(A_ineq); put(lbA_ineq); put(ubA_ineq); put(U_opt); }
    }
    fclose(g_out);
    return 0;
}

int main(int argc, char **argv) {
    // QPSolver::updateState/getState print the state on every call (QPSolver.cpp:110,114), solveQP reports failures on
    // stderr (:99): keep stdout out of the way, leave stderr alone.
    if (!freopen("/dev/null", "w", stdout)) return 2;
    if (argc == 4 && !strcmp(argv[1], "cases")) return run_cases(argv[2], argv[3]);
    if (argc == 3 && !strcmp(argv[1], "demo")) return run_demo(argv[2]);
    fprintf(stderr, "usage: ref_dump cases <in.bin> <out.bin> | ref_dump demo <out.bin>\n");
    return 1;
}
