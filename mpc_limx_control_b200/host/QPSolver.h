// QPSolver.h -- host facade with the interface of the reference class QPSolver
// (reference include/QPSolver.h:10-56): same method names, argument order and meaning.  Every
// method body is a call into the C ABI (include/mpc_b200.h); there is no CPU implementation here.
//
// Differences from the reference, all deliberate:
//   - sizes are generic: the header's Vector4d / Matrix<2,15> / Vector2d hard-typing of the demo
//     (include/QPSolver.h:22,31,34,37) becomes VectorXd / MatrixXd; the .cpp was already generic
//   - solveQP returns false when the device solver does not certify optimality (the reference
//     ignores qpOASES' status and always returns true, src/QPSolver.cpp:98-105)
//   - rows of A_total whose bounds are equal and finite are honoured as equalities, but the
//     reference's spurious A_eq/b_eq block should not be stacked (the examples do not)
//   - updateState / getState do not print (src/QPSolver.cpp:110,114)
#pragma once
#include <stdexcept>
#include <string>

#include "../../include/mpc_b200.h"
#include "mat.h"

namespace mpcb200 {
namespace host {

class DeviceError : public std::runtime_error {
public:
    DeviceError(int code, const std::string& what) : std::runtime_error(what), code(code) {}
    int code;
};

class QPSolver {
public:
    QPSolver(double Ts, int N, const MatrixXd& Ac, const MatrixXd& Bc, const MatrixXd& Q, const MatrixXd& R,
             const MatrixXd& P, const VectorXd& x_min, const VectorXd& x_max, double u_min, double u_max,
             int device = 0);
    ~QPSolver();
    QPSolver(const QPSolver&) = delete;
    QPSolver& operator=(const QPSolver&) = delete;

    void discretizeSystem();

    void buildQPParams(const VectorXd& xi0, const MatrixXd& xi_ref, MatrixXd& H, VectorXd& f, MatrixXd& A_eq,
                       VectorXd& b_eq, VectorXd& lb, VectorXd& ub, MatrixXd& A_ineq, VectorXd& lbA_ineq,
                       VectorXd& ubA_ineq);

    bool solveQP(const MatrixXd& H, const VectorXd& f, const MatrixXd& A_total, const VectorXd& lb, const VectorXd& ub,
                 const VectorXd& lbA_total, const VectorXd& ubA_total, MatrixXd& U_opt);

    void updateState(const VectorXd& u);
    VectorXd getState() const;
    void setState(const VectorXd& xi);

    // extras for parity tests
    const MatrixXd& getAd() const { return Ad; }
    const MatrixXd& getBd() const { return Bd; }
    int lastStatus() const { return last_status; }
    int lastIterations() const { return last_iters; }

private:
    void check(int rc, const char* where) const;
    double Ts;
    int N, NX, NU;
    MatrixXd Ac, Bc, Ad, Bd, Q, R, P;
    VectorXd x_min, x_max;
    double u_min, u_max;
    VectorXd xi;
    mpc_b200_lti* ctx;
    int last_status, last_iters;
};

}  // namespace host
}  // namespace mpcb200
