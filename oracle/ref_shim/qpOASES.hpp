// Minimal <qpOASES.hpp> stand-in -- TEST INFRASTRUCTURE, NOT qpOASES.  See Eigen/Dense in this directory.
//
// qpOASES (unpinned third-party dependency of the reference, CMakeLists.txt:30,52) is absent from the image.  This
// header only lets src/QPSolver.cpp:83-106 compile unmodified and routes QProblem::init to the oracle's cold-start
// dense active-set solver (oracle/mpc_oracle.c: orc_qp_solve), i.e. it pins the reference's DATA PATH into and out of
// the solver (argument order, sizes, U_opt layout) -- not qpOASES's own arithmetic, which stays unpinned.
//
// qpOASES reads the constraint matrix ROW-major (nC x nV).  The reference hands it a column-major Eigen buffer
// (src/QPSolver.cpp:93-96).  qpOASES::shim_A_is_colmajor selects what the shim does with it:
//   false (default) = what qpOASES would do: read the buffer row-major (the reference-as-written defect);
//   true            = read it column-major, i.e. solve the problem the caller meant.
#ifndef MPC_B200_REF_SHIM_QPOASES
#define MPC_B200_REF_SHIM_QPOASES
#include <vector>

extern "C" int orc_qp_solve(int n, const double *H, const double *f, int mA, const double *A, const double *lbA,
                            const double *ubA, const double *lb, const double *ub, double *u, double *y_bnd,
                            double *y_row, int *iters);

namespace qpOASES {

typedef double real_t;
typedef int int_t;
const real_t INFTY = 1.0e20;
enum PrintLevel { PL_NONE = 0, PL_LOW, PL_MEDIUM, PL_HIGH };
enum returnValue { SUCCESSFUL_RETURN = 0, RET_MAX_NWSR_REACHED = 64, RET_INIT_FAILED_INFEASIBILITY = 37 };
inline bool shim_A_is_colmajor = false;
inline int shim_last_status = 0;     // orc_qp_solve status of the last init (0 solved)
inline int shim_last_iters = 0;

struct Options {
    PrintLevel printLevel;
    Options() : printLevel(PL_MEDIUM) {}
};

class QProblem {
    int nV_, nC_;
    std::vector<double> x_;
    bool solved_;
public:
    QProblem(int_t nV, int_t nC) : nV_(nV), nC_(nC), x_(nV, 0.0), solved_(false) {}
    void setOptions(const Options &) {}
    returnValue init(const real_t *H, const real_t *g, const real_t *A, const real_t *lb, const real_t *ub,
                     const real_t *lbA, const real_t *ubA, int_t &nWSR) {
        std::vector<double> Acm((size_t)nC_ * nV_);
        for (int i = 0; i < nC_; ++i)
            for (int j = 0; j < nV_; ++j)
                Acm[(size_t)j * nC_ + i] = shim_A_is_colmajor ? A[(size_t)j * nC_ + i] : A[(size_t)i * nV_ + j];
        int iters = 0;
        int st = orc_qp_solve(nV_, H, g, nC_, nC_ ? Acm.data() : nullptr, lbA, ubA, lb, ub, x_.data(), nullptr, nullptr, &iters);
        shim_last_status = st; shim_last_iters = iters; nWSR = iters;
        solved_ = (st == 0);
        return st == 0 ? SUCCESSFUL_RETURN : (st == 1 ? RET_MAX_NWSR_REACHED : RET_INIT_FAILED_INFEASIBILITY);
    }
    // qpOASES leaves the output untouched when the problem is not solved (RET_QP_NOT_SOLVED)
    returnValue getPrimalSolution(real_t *x) const {
        if (!solved_) return RET_INIT_FAILED_INFEASIBILITY;
        for (int i = 0; i < nV_; ++i) x[i] = x_[i];
        return SUCCESSFUL_RETURN;
    }
};

}  // namespace qpOASES
#endif
