// emul_lti.cpp -- TEST INFRASTRUCTURE ONLY: host build (one-thread group) of the product's generic
// LTI device source (csrc/lti_core.cuh) so its mathematics is checked in the CPU test tier.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mpc_limx_control_b200/csrc/lti_core.cuh"

using namespace mpcb200;

struct GrpSerial {
    int tid() const { return 0; }
    int size() const { return 1; }
    void sync() const {}
};

extern "C" {

void emul_lti_discretize(int NX, int NU, double Ts, const double* Ac, const double* Bc, double* Ad, double* Bd) {
    int m = NX + NU;
    std::vector<double> work(3 * m * m);
    lti_discretize(NX, NU, Ts, Ac, Bc, Ad, Bd, work.data(), GrpSerial());
}

void emul_lti_build(int NX, int NU, int N, const double* Ad, const double* Bd, const double* Q, const double* R,
                    const double* P, const double* x_min, const double* x_max, double u_min, double u_max,
                    const double* xi0, const double* xi_ref, double* H, double* f, double* A_eq, double* b_eq, double* lb,
                    double* ub, double* A_ineq, double* lbA, double* ubA, double* A_aug, double* B_aug) {
    LtiDims d{NX, NU, N};
    std::vector<double> work((size_t)d.p() * d.n() + d.p());
    lti_build(d, Ad, Bd, Q, R, P, x_min, x_max, u_min, u_max, xi0, xi_ref, A_aug, B_aug, H, f, A_eq, b_eq, lb, ub, A_ineq,
              lbA, ubA, work.data(), GrpSerial());
}

int emul_qp_dense(int n, int m, const double* H, const double* f, const double* A, const double* lb, const double* ub,
                  const double* lbA, const double* ubA, double* U, int* iters, int max_newton, int max_admm) {
    std::vector<double> work(qp_dense_work_doubles(n, m));
    QpWork W = qp_dense_carve(work.data(), n, m);
    return qp_dense_solve(n, H, f, m, A, lb, ub, lbA, ubA, U, iters, W, max_newton, max_admm, 1e-9, GrpSerial());
}

void emul_lti_update(int NX, int NU, const double* Ad, const double* Bd, double* xi, const double* u) {
    std::vector<double> work(NX);
    lti_update(NX, NU, Ad, Bd, xi, u, work.data(), GrpSerial());
}
}
