// QPSolver.cpp -- facade bodies: each reference method (src/QPSolver.cpp) becomes one C-ABI call.
#include "QPSolver.h"

namespace mpcb200 {
namespace host {

void QPSolver::check(int rc, const char* where) const {
    if (rc == MPC_B200_OK) return;
    std::string msg = std::string(where) + ": " + mpc_b200_strerror(rc);
    if (ctx) { msg += " ("; msg += mpc_b200_lti_last_error(ctx); msg += ")"; }
    throw DeviceError(rc, msg);
}

// reference src/QPSolver.cpp:3-19
QPSolver::QPSolver(double Ts_, int N_, const MatrixXd& Ac_, const MatrixXd& Bc_, const MatrixXd& Q_, const MatrixXd& R_,
                   const MatrixXd& P_, const VectorXd& x_min_, const VectorXd& x_max_, double u_min_, double u_max_,
                   int device)
    : Ts(Ts_), N(N_), Ac(Ac_), Bc(Bc_), Q(Q_), R(R_), P(P_), x_min(x_min_), x_max(x_max_), u_min(u_min_), u_max(u_max_),
      ctx(nullptr), last_status(0), last_iters(0) {
    NX = Ac.rows();
    NU = Bc.cols();
    if (Ac.cols() != NX || Bc.rows() != NX || Q.rows() != NX || Q.cols() != NX || P.rows() != NX || P.cols() != NX ||
        R.rows() != NU || R.cols() != NU || x_min.size() != NX || x_max.size() != NX || N < 1)
        throw DeviceError(MPC_B200_EINVAL, "QPSolver: inconsistent dimensions");
    if (NX + NU > 64) throw DeviceError(MPC_B200_EINVAL, "QPSolver: NX + NU > 64 is not supported by the discretisation kernel");
    xi = VectorXd::Zero(NX);
    check(mpc_b200_lti_create(device, &ctx), "QPSolver");
    try {
        discretizeSystem();
    } catch (...) {   // the destructor does not run for a throwing constructor: release the context here
        mpc_b200_lti_destroy(ctx);
        ctx = nullptr;
        throw;
    }
}

QPSolver::~QPSolver() {
    if (ctx) mpc_b200_lti_destroy(ctx);
}

// reference src/QPSolver.cpp:21-29
void QPSolver::discretizeSystem() {
    Ad.resize(NX, NX);
    Bd.resize(NX, NU);
    check(mpc_b200_lti_discretize(ctx, 1, NX, NU, Ts, Ac.data(), Bc.data(), Ad.data(), Bd.data()), "discretizeSystem");
}

// reference src/QPSolver.cpp:31-81
void QPSolver::buildQPParams(const VectorXd& xi0, const MatrixXd& xi_ref, MatrixXd& H, VectorXd& f, MatrixXd& A_eq,
                             VectorXd& b_eq, VectorXd& lb, VectorXd& ub, MatrixXd& A_ineq, VectorXd& lbA_ineq,
                             VectorXd& ubA_ineq) {
    if (xi0.size() != NX || xi_ref.rows() != NX || xi_ref.cols() != N + 1)
        throw DeviceError(MPC_B200_EINVAL, "buildQPParams: xi0 must be NX, xi_ref NX x (N+1)");
    const int n = NU * N;
    H.resize(n, n); f.resize(n);
    A_eq.resize(NX * N, n); b_eq.resize(NX * N);
    lb.resize(n); ub.resize(n);
    A_ineq.resize(2 * NX * N, n); lbA_ineq.resize(2 * NX * N); ubA_ineq.resize(2 * NX * N);
    check(mpc_b200_lti_build_qp(ctx, 1, NX, NU, N, Ad.data(), Bd.data(), Q.data(), R.data(), P.data(), x_min.data(),
                                x_max.data(), u_min, u_max, xi0.data(), xi_ref.data(), H.data(), f.data(), A_eq.data(),
                                b_eq.data(), lb.data(), ub.data(), A_ineq.data(), lbA_ineq.data(), ubA_ineq.data(),
                                nullptr, nullptr),
          "buildQPParams");
}

// reference src/QPSolver.cpp:83-106
bool QPSolver::solveQP(const MatrixXd& H, const VectorXd& f, const MatrixXd& A_total, const VectorXd& lb,
                       const VectorXd& ub, const VectorXd& lbA_total, const VectorXd& ubA_total, MatrixXd& U_opt) {
    const int n = NU * N, m = A_total.rows();
    if (H.rows() != n || H.cols() != n || f.size() != n || lb.size() != n || ub.size() != n ||
        (m > 0 && (A_total.cols() != n || lbA_total.size() != m || ubA_total.size() != m)))
        throw DeviceError(MPC_B200_EINVAL, "solveQP: inconsistent dimensions");
    U_opt.resize(NU, N);   // column k = u_k (src/QPSolver.cpp:104)
    int32_t st = 2, it = 0;
    check(mpc_b200_qp_solve_dense(ctx, 1, n, m, H.data(), f.data(), m ? A_total.data() : nullptr, lb.data(), ub.data(),
                                  m ? lbA_total.data() : nullptr, m ? ubA_total.data() : nullptr, U_opt.data(), &st, &it),
          "solveQP");
    last_status = st;
    last_iters = it;
    return st == 0;
}

// reference src/QPSolver.cpp:108-111
void QPSolver::updateState(const VectorXd& u) {
    if (u.size() != NU) throw DeviceError(MPC_B200_EINVAL, "updateState: u must be NU");
    check(mpc_b200_lti_update_state(ctx, 1, NX, NU, Ad.data(), Bd.data(), xi.data(), u.data()), "updateState");
}

// reference src/QPSolver.cpp:113-116
VectorXd QPSolver::getState() const { return xi; }
void QPSolver::setState(const VectorXd& x) {
    if (x.size() != NX) throw DeviceError(MPC_B200_EINVAL, "setState: x must be NX");
    xi = x;
}

}  // namespace host
}  // namespace mpcb200
