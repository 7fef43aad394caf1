// lti_core.cuh -- generic condensed-MPC path behind the reference's QPSolver class, any (NX, NU, N).
//
// Group-cooperative code like tron1_core.cuh (one group = one CTA per problem instance on the GPU,
// one serial thread in the test-only host build).  All matrices column-major (Eigen layout), all
// workspaces in global memory: this path serves the single-robot facade and the 4-state demo
// (reference src/qpSolver_test.cpp), not the batched TRON1 throughput path.
//
//   lti_discretize   src/QPSolver.cpp:21-29   exact ZOH: exp([[Ac,Bc],[0,0]] Ts) by scaling and
//                                             squaring of a degree-18 Taylor polynomial (Horner)
//   lti_build        src/QPSolver.cpp:31-81   A_aug, B_aug (recursion B(i,j) = Ad B(i-1,j), equal to
//                                             Ad^(i-j-1) Bd), H, f, box and state-bound rows
//   qp_dense_solve   src/QPSolver.cpp:83-106  min 1/2 u'Hu + f'u, lb<=u<=ub, lbA<=Au<=ubA:
//                                             primal-dual active-set iterations on a Cholesky of H with
//                                             a Schur complement on the active rows, ADMM fallback
//   lti_update       src/QPSolver.cpp:108-111 xi <- Ad xi + Bd u
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

namespace mpcb200 {

#define LTI_CM(M, ld, i, j) ((M)[(size_t)(i) + (size_t)(ld) * (size_t)(j)])
#define LTI_INF_HALF 0.5e20

// C(m x n) = alpha * A(m x k) B(k x n) + beta * D(m x n)   (D may be null when beta == 0)
template <class G>
MPC_HD void lti_mm(int m, int n, int k, double alpha, const double* A, const double* B, double beta, const double* D,
                   double* C, const G& g) {
    for (int idx = g.tid(); idx < m * n; idx += g.size()) {
        int i = idx % m, j = idx / m;
        double s = 0.0;
        for (int l = 0; l < k; ++l) s += LTI_CM(A, m, i, l) * LTI_CM(B, k, l, j);
        C[idx] = alpha * s + (beta != 0.0 ? beta * D[idx] : 0.0);
    }
    g.sync();
}

// work: 3 * m * m doubles, m = NX + NU
template <class G>
MPC_HD void lti_discretize(int NX, int NU, double Ts, const double* Ac, const double* Bc, double* Ad, double* Bd,
                           double* work, const G& g) {
    const int m = NX + NU, mm = m * m;
    double *A = work, *E = work + mm, *T = work + 2 * mm;
    for (int idx = g.tid(); idx < mm; idx += g.size()) {
        int i = idx % m, j = idx / m;
        double v = 0.0;
        if (i < NX) v = (j < NX ? LTI_CM(Ac, NX, i, j) : LTI_CM(Bc, NX, i, j - NX)) * Ts;
        A[idx] = v;
    }
    g.sync();
    // 1-norm -> number of squarings so that |A / 2^s|_1 <= 1/2
    double nrm = 0.0;
    for (int j = 0; j < m; ++j) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += fabs(LTI_CM(A, m, i, j));
        nrm = s > nrm ? s : nrm;
    }
    int sq = 0;
    while (nrm > 0.5 && sq < 60) { nrm *= 0.5; ++sq; }
    const double sc = ldexp(1.0, -sq);
    g.sync();
    for (int idx = g.tid(); idx < mm; idx += g.size()) A[idx] *= sc;
    g.sync();
    // Horner: E = I + A/1 (I + A/2 (I + ... (I + A/18)))
    const int deg = 18;
    for (int idx = g.tid(); idx < mm; idx += g.size()) E[idx] = (idx % m == idx / m) ? 1.0 : 0.0;
    g.sync();
    for (int d = deg; d >= 1; --d) {
        lti_mm(m, m, m, 1.0 / (double)d, A, E, 0.0, (const double*)0, T, g);
        for (int idx = g.tid(); idx < mm; idx += g.size()) E[idx] = T[idx] + ((idx % m == idx / m) ? 1.0 : 0.0);
        g.sync();
    }
    for (int s = 0; s < sq; ++s) {
        lti_mm(m, m, m, 1.0, E, E, 0.0, (const double*)0, T, g);
        for (int idx = g.tid(); idx < mm; idx += g.size()) E[idx] = T[idx];
        g.sync();
    }
    for (int idx = g.tid(); idx < NX * NX; idx += g.size()) Ad[idx] = LTI_CM(E, m, idx % NX, idx / NX);
    for (int idx = g.tid(); idx < NX * NU; idx += g.size()) Bd[idx] = LTI_CM(E, m, idx % NX, NX + idx / NX);
    g.sync();
}

struct LtiDims {
    int NX, NU, N;
    MPC_HD int p() const { return NX * (N + 1); }
    MPC_HD int n() const { return NU * N; }
};

// work: p*n (QB) + p (e) doubles.  Any output pointer may be null except A_aug/B_aug (needed internally).
template <class G>
MPC_HD void lti_build(LtiDims d, const double* Ad, const double* Bd, const double* Q, const double* R, const double* P,
                      const double* x_min, const double* x_max, double u_min, double u_max, const double* xi0,
                      const double* xi_ref, double* A_aug, double* B_aug, double* H, double* f, double* A_eq, double* b_eq,
                      double* lb, double* ub, double* A_ineq, double* lbA, double* ubA, double* work, const G& g) {
    const int NX = d.NX, NU = d.NU, N = d.N, p = d.p(), n = d.n();
    double* QB = work;
    double* e = work + (size_t)p * n;
    // A_aug (src/QPSolver.cpp:36-40)
    for (int idx = g.tid(); idx < p * NX; idx += g.size()) A_aug[idx] = 0.0;
    for (int idx = g.tid(); idx < p * n; idx += g.size()) B_aug[idx] = 0.0;
    g.sync();
    for (int i = g.tid(); i < NX; i += g.size()) LTI_CM(A_aug, p, i, i) = 1.0;
    g.sync();
    for (int i = 1; i <= N; ++i) {
        for (int idx = g.tid(); idx < NX * NX; idx += g.size()) {
            int r = idx % NX, c = idx / NX;
            double s = 0.0;
            for (int l = 0; l < NX; ++l) s += LTI_CM(Ad, NX, r, l) * LTI_CM(A_aug, p, (i - 1) * NX + l, c);
            LTI_CM(A_aug, p, i * NX + r, c) = s;
        }
        // B_aug block row i (src/QPSolver.cpp:42-47): block(i,i-1) = Bd, block(i,j) = Ad * block(i-1,j)
        for (int idx = g.tid(); idx < NX * NU * i; idx += g.size()) {
            int r = idx % NX, cc = idx / NX;   // cc in [0, NU*i)
            int j = cc / NU;
            double s;
            if (j == i - 1) s = LTI_CM(Bd, NX, r, cc % NU);
            else {
                s = 0.0;
                for (int l = 0; l < NX; ++l) s += LTI_CM(Ad, NX, r, l) * LTI_CM(B_aug, p, (i - 1) * NX + l, cc);
            }
            LTI_CM(B_aug, p, i * NX + r, cc) = s;
        }
        g.sync();
    }
    // QB = Q_bar B_aug (block diagonal Q_bar, src/QPSolver.cpp:50-56) ; e = A_aug xi0 - vec(xi_ref)
    for (int idx = g.tid(); idx < p * n; idx += g.size()) {
        int row = idx % p, c = idx / p;
        int i = row / NX, r = row % NX;
        const double* W = (i == N) ? P : Q;
        double s = 0.0;
        for (int l = 0; l < NX; ++l) s += LTI_CM(W, NX, r, l) * LTI_CM(B_aug, p, i * NX + l, c);
        QB[idx] = s;
    }
    for (int row = g.tid(); row < p; row += g.size()) {
        double s = 0.0;
        for (int l = 0; l < NX; ++l) s += LTI_CM(A_aug, p, row, l) * xi0[l];
        e[row] = s - xi_ref[row];
    }
    g.sync();
    // H = 2 (B' QB + R_bar) ; f = 2 QB' e   (src/QPSolver.cpp:58-60; Q_bar symmetric)
    if (H)
        for (int idx = g.tid(); idx < n * n; idx += g.size()) {
            int i = idx % n, j = idx / n;
            double s = 0.0;
            for (int l = 0; l < p; ++l) s += LTI_CM(B_aug, p, l, i) * LTI_CM(QB, p, l, j);
            if (i / NU == j / NU) s += LTI_CM(R, NU, i % NU, j % NU);
            H[idx] = 2.0 * s;
        }
    if (f)
        for (int j = g.tid(); j < n; j += g.size()) {
            double s = 0.0;
            for (int l = 0; l < p; ++l) s += LTI_CM(QB, p, l, j) * e[l];
            f[j] = 2.0 * s;
        }
    // src/QPSolver.cpp:62-64 (kept for interface completeness; spurious block)
    if (A_eq)
        for (int idx = g.tid(); idx < NX * N * n; idx += g.size()) {
            int r = idx % (NX * N), c = idx / (NX * N);
            A_eq[idx] = LTI_CM(B_aug, p, NX + r, c);
        }
    if (b_eq)
        for (int r = g.tid(); r < NX * N; r += g.size()) {
            double s = 0.0;
            for (int l = 0; l < NX; ++l) s += LTI_CM(A_aug, p, NX + r, l) * xi0[l];
            b_eq[r] = s;
        }
    // :66-68
    for (int i = g.tid(); i < n; i += g.size()) { if (lb) lb[i] = u_min; if (ub) ub[i] = u_max; }
    // :70-80   rows [2i NX, 2i NX + NX) = B_pred(i+1); the other NX rows stay zero with +-INFTY
    const int mi = 2 * NX * N;
    if (A_ineq)
        for (int idx = g.tid(); idx < mi * n; idx += g.size()) {
            int row = idx % mi, c = idx / mi;
            int i = row / (2 * NX), r = row % (2 * NX);
            A_ineq[idx] = r < NX ? LTI_CM(B_aug, p, (i + 1) * NX + r, c) : 0.0;
        }
    for (int row = g.tid(); row < mi; row += g.size()) {
        int i = row / (2 * NX), r = row % (2 * NX);
        double lo = -1.0e20, hi = 1.0e20;
        if (r < NX) {
            double s = 0.0;   // Ad^(i+1) xi0 == block row i+1 of A_aug times xi0
            for (int l = 0; l < NX; ++l) s += LTI_CM(A_aug, p, (i + 1) * NX + r, l) * xi0[l];
            lo = x_min[r] - s; hi = x_max[r] - s;
        }
        if (lbA) lbA[row] = lo;
        if (ubA) ubA[row] = hi;
    }
    g.sync();
}

template <class G>
MPC_HD void lti_update(int NX, int NU, const double* Ad, const double* Bd, double* xi, const double* u, double* work, const G& g) {
    for (int r = g.tid(); r < NX; r += g.size()) {
        double s = 0.0;
        for (int j = 0; j < NX; ++j) s += LTI_CM(Ad, NX, r, j) * xi[j];
        for (int j = 0; j < NU; ++j) s += LTI_CM(Bd, NX, r, j) * u[j];
        work[r] = s;
    }
    g.sync();
    for (int r = g.tid(); r < NX; r += g.size()) xi[r] = work[r];
    g.sync();
}

// ------------------------------------------------------------------------------------------------
// dense QP.  Row index r in [0, n+m): r < n is the bound on u_r, r >= n is row r-n of A.
struct QpWork {
    // all sized by the caller: see qp_dense_work_doubles()
    double *L, *Y, *S, *K;          // n*n each
    double *v0, *lam, *rhsl, *t;    // n, n+m, n, n+m
    double *z, *y, *w, *ut;         // n+m, n+m, n, n
    int *act, *list;                // n+m, n
    int* flag;                      // 4 ints of group-uniform scratch
};
MPC_HD size_t qp_dense_work_doubles(int n, int m) {
    return (size_t)4 * n * n + (size_t)4 * n + (size_t)4 * (n + m) + (size_t)(n + m) / 2 + n / 2 + 8;
}
MPC_HD QpWork qp_dense_carve(double* base, int n, int m) {
    QpWork W;
    size_t nn = (size_t)n * n, mt = (size_t)n + m;
    double* p = base;
    W.L = p; p += nn; W.Y = p; p += nn; W.S = p; p += nn; W.K = p; p += nn;
    W.v0 = p; p += n; W.rhsl = p; p += n; W.w = p; p += n; W.ut = p; p += n;
    W.lam = p; p += mt; W.t = p; p += mt; W.z = p; p += mt; W.y = p; p += mt;
    W.act = (int*)p; p += (mt + 1) / 2;
    W.list = (int*)p; p += (n + 1) / 2;
    W.flag = (int*)p;
    return W;
}

// in-place lower Cholesky of the column-major n x n matrix M (ld n); returns false if not PD
template <class G>
MPC_HD bool dense_cholesky(int n, double* M, int* flag, const G& g) {
    if (g.tid() == 0) flag[0] = 0;
    g.sync();
    for (int k = 0; k < n; ++k) {
        // left-looking column k
        double d = LTI_CM(M, n, k, k);
        for (int j = 0; j < k; ++j) d -= LTI_CM(M, n, k, j) * LTI_CM(M, n, k, j);
        if (!(d > 0.0)) { flag[0] = 1; d = 1.0; }
        double di = 1.0 / sqrt(d);
        g.sync();   // everyone has read the old diagonal
        for (int i = k + g.tid(); i < n; i += g.size()) {
            if (i == k) { LTI_CM(M, n, k, k) = d * di; continue; }
            double s = LTI_CM(M, n, i, k);
            for (int j = 0; j < k; ++j) s -= LTI_CM(M, n, i, j) * LTI_CM(M, n, k, j);
            LTI_CM(M, n, i, k) = s * di;
        }
        g.sync();
    }
    return flag[0] == 0;
}
// x <- L^-1 x (serial, one thread); x <- L^-T x
MPC_HD void tri_fwd(int n, const double* L, double* x) {
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int j = 0; j < i; ++j) s -= LTI_CM(L, n, i, j) * x[j];
        x[i] = s / LTI_CM(L, n, i, i);
    }
}
MPC_HD void tri_bwd(int n, const double* L, double* x) {
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int j = i + 1; j < n; ++j) s -= LTI_CM(L, n, j, i) * x[j];
        x[i] = s / LTI_CM(L, n, i, i);
    }
}

MPC_HD double qp_row_dot(int n, int m, const double* A, int r, const double* u) {
    if (r < n) return u[r];
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += LTI_CM(A, m, r - n, j) * u[j];
    return s;
}
MPC_HD double qp_lo(int n, const double* lb, const double* lbA, int r) { return r < n ? lb[r] : lbA[r - n]; }
MPC_HD double qp_hi(int n, const double* ub, const double* ubA, int r) { return r < n ? ub[r] : ubA[r - n]; }

// equality-constrained solve on the active rows (act != 0); returns false when the Schur complement is
// not positive definite (dependent active rows)
template <class G>
MPC_HD bool qp_active_solve(int n, int m, const double* A, const double* lb, const double* ub, const double* lbA,
                            const double* ubA, QpWork& W, double* u, const G& g) {
    const int mt = n + m;
    if (g.tid() == 0) {
        int k = 0;
        for (int r = 0; r < mt && k < n; ++r) if (W.act[r]) W.list[k++] = r;
        int extra = 0;
        for (int r = 0; r < mt; ++r) extra += W.act[r] != 0;
        W.flag[1] = k;
        W.flag[2] = extra > n;   // more active rows than variables: dependent by counting
    }
    g.sync();
    const int k = W.flag[1];
    if (W.flag[2]) return false;
    // Y = L^-1 C_R'  (one column per thread)
    for (int c = g.tid(); c < k; c += g.size()) {
        double* y = W.Y + (size_t)n * c;
        int r = W.list[c];
        for (int j = 0; j < n; ++j) y[j] = r < n ? (j == r ? 1.0 : 0.0) : LTI_CM(A, m, r - n, j);
        tri_fwd(n, W.L, y);
    }
    g.sync();
    // S = Y'Y ; rhs = b_R + C_R v0   with v0 = H^-1 f  (W.v0 holds L^-1 f, so C_R H^-1 f = Y' (L^-1 f))
    for (int idx = g.tid(); idx < k * k; idx += g.size()) {
        int i = idx % k, j = idx / k;
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += W.Y[(size_t)n * i + l] * W.Y[(size_t)n * j + l];
        W.S[(size_t)i + (size_t)k * j] = s;
    }
    for (int c = g.tid(); c < k; c += g.size()) {
        int r = W.list[c];
        double b = W.act[r] < 0 ? qp_lo(n, lb, lbA, r) : qp_hi(n, ub, ubA, r);
        if (W.act[r] == 2) b = qp_lo(n, lb, lbA, r);
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += W.Y[(size_t)n * c + l] * W.v0[l];
        W.rhsl[c] = b + s;
    }
    g.sync();
    if (k > 0) {
        if (!dense_cholesky(k, W.S, W.flag, g)) return false;
        if (g.tid() == 0) { tri_fwd(k, W.S, W.rhsl); tri_bwd(k, W.S, W.rhsl); }
        g.sync();
    }
    // u = L^-T (Y lam - L^-1 f)
    for (int i = g.tid(); i < n; i += g.size()) {
        double s = -W.v0[i];
        for (int c = 0; c < k; ++c) s += W.Y[(size_t)n * c + i] * W.rhsl[c];
        W.w[i] = s;
    }
    g.sync();
    if (g.tid() == 0) {
        tri_bwd(n, W.L, W.w);
        for (int i = 0; i < n; ++i) u[i] = W.w[i];
        for (int r = 0; r < mt; ++r) W.lam[r] = 0.0;
        for (int c = 0; c < k; ++c) W.lam[W.list[c]] = W.rhsl[c];
    }
    g.sync();
    return true;
}

// KKT check of (u, lam) + primal-dual active-set update.  Returns true when (u, lam) is optimal.
template <class G>
MPC_HD bool qp_check_update(int n, int m, const double* A, const double* lb, const double* ub, const double* lbA,
                            const double* ubA, QpWork& W, const double* u, double tol, bool update, const G& g) {
    const int mt = n + m;
    if (g.tid() == 0) { W.flag[0] = 0; W.flag[3] = 0; }
    g.sync();
    double umax = 1.0;
    for (int i = 0; i < n; ++i) umax = fabs(u[i]) > umax ? fabs(u[i]) : umax;
    for (int r = g.tid(); r < mt; r += g.size()) {
        double lo = qp_lo(n, lb, lbA, r), hi = qp_hi(n, ub, ubA, r);
        bool has_lo = lo > -LTI_INF_HALF, has_hi = hi < LTI_INF_HALF;
        if (!has_lo && !has_hi) { W.act[r] = 0; continue; }
        double t = qp_row_dot(n, m, A, r, u);
        double scale = 1.0;
        if (r >= n) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += LTI_CM(A, m, r - n, j) * LTI_CM(A, m, r - n, j);
            if (s == 0.0) { W.act[r] = 0; continue; }
            scale = sqrt(s);
        }
        const double ptol = tol * umax * scale;
        int a = W.act[r], na = a;
        bool bad = false;
        if (has_lo && has_hi && hi - lo <= 0.0) {
            na = 2;
            if (a != 2) bad = true;
        } else if (a == 0) {
            if (has_lo && t < lo - ptol) { na = -1; bad = true; }
            else if (has_hi && t > hi + ptol) { na = 1; bad = true; }
        } else {
            double l = W.lam[r];
            // H u + f = C' lam : lam >= 0 at a lower bound, <= 0 at an upper bound
            if (a < 0 && l < -tol * umax) { na = 0; bad = true; }
            if (a == 1 && l > tol * umax) { na = 0; bad = true; }
        }
        if (bad) W.flag[0] = 1;
        if (update && na != a) { W.act[r] = na; W.flag[3] = 1; }
    }
    g.sync();
    return W.flag[0] == 0;
}

// returns status (0 solved, 1 iteration limit, 2 failed).  H, A column-major; u out (n).
template <class G>
MPC_HD int qp_dense_solve(int n, const double* H, const double* f, int m, const double* A, const double* lb,
                          const double* ub, const double* lbA, const double* ubA, double* u, int* iters_out,
                          QpWork W, int max_newton, int max_admm, double tol, const G& g) {
    const int mt = n + m;
    int iters = 0;
    for (int idx = g.tid(); idx < n * n; idx += g.size()) W.L[idx] = H[idx];
    for (int r = g.tid(); r < mt; r += g.size()) { W.act[r] = 0; W.lam[r] = 0.0; }
    for (int i = g.tid(); i < n; i += g.size()) W.v0[i] = f[i];
    g.sync();
    if (!dense_cholesky(n, W.L, W.flag, g)) { if (iters_out) *iters_out = 0; return 2; }
    if (g.tid() == 0) tri_fwd(n, W.L, W.v0);   // v0 = L^-1 f
    g.sync();
    // equalities enter the first active set
    for (int r = g.tid(); r < mt; r += g.size()) {
        double lo = qp_lo(n, lb, lbA, r), hi = qp_hi(n, ub, ubA, r);
        if (lo > -LTI_INF_HALF && hi < LTI_INF_HALF && hi - lo <= 0.0) W.act[r] = 2;
    }
    g.sync();
    bool solved = false;
    for (int it = 0; it < max_newton && !solved; ++it) {
        ++iters;
        if (!qp_active_solve(n, m, A, lb, ub, lbA, ubA, W, u, g)) break;
        solved = qp_check_update(n, m, A, lb, ub, lbA, ubA, W, u, tol, true, g);
        if (!solved && W.flag[3] == 0) break;   // nothing to change yet not optimal
    }
    if (solved) { if (iters_out) *iters_out = iters; return 0; }

    // ---- ADMM fallback (OSQP splitting): K = H + sigma I + rho C'C --------------------------------
    const double sigma = 1e-6;
    double hmax = 0.0;
    for (int i = 0; i < n; ++i) hmax = LTI_CM(H, n, i, i) > hmax ? LTI_CM(H, n, i, i) : hmax;
    const double rho = sqrt(hmax) * 0.5 + 1e-3;
    for (int idx = g.tid(); idx < n * n; idx += g.size()) {
        int i = idx % n, j = idx / n;
        double s = LTI_CM(H, n, i, j) + (i == j ? sigma + rho : 0.0);
        for (int r = 0; r < m; ++r) s += rho * LTI_CM(A, m, r, i) * LTI_CM(A, m, r, j);
        W.K[idx] = s;
    }
    for (int i = g.tid(); i < n; i += g.size()) W.ut[i] = 0.0;
    for (int r = g.tid(); r < mt; r += g.size()) { W.z[r] = 0.0; W.y[r] = 0.0; }
    g.sync();
    if (!dense_cholesky(n, W.K, W.flag, g)) { if (iters_out) *iters_out = iters; return 2; }
    const double alpha = 1.6;
    int status = 1;
    for (int it = 0; it < max_admm; ++it) {
        ++iters;
        for (int i = g.tid(); i < n; i += g.size()) {
            double s = sigma * W.ut[i] - f[i] + (rho * W.z[i] - W.y[i]);
            for (int r = 0; r < m; ++r) s += LTI_CM(A, m, r, i) * (rho * W.z[n + r] - W.y[n + r]);
            W.w[i] = s;
        }
        g.sync();
        if (g.tid() == 0) { tri_fwd(n, W.K, W.w); tri_bwd(n, W.K, W.w); }
        g.sync();
        for (int r = g.tid(); r < mt; r += g.size()) {
            double zt = qp_row_dot(n, m, A, r, W.w);
            double zr = alpha * zt + (1.0 - alpha) * W.z[r];
            double v = zr + W.y[r] / rho;
            double lo = qp_lo(n, lb, lbA, r), hi = qp_hi(n, ub, ubA, r);
            double zn = v < lo ? lo : (v > hi ? hi : v);
            W.y[r] += rho * (zr - zn);
            W.z[r] = zn;
            W.t[r] = v;
        }
        g.sync();
        for (int i = g.tid(); i < n; i += g.size()) W.ut[i] = alpha * W.w[i] + (1.0 - alpha) * W.ut[i];
        g.sync();
        if ((it + 1) % 25 == 0) {
            // polish: active set from the clipped rows, exact equality solve, KKT verification
            for (int r = g.tid(); r < mt; r += g.size()) {
                double lo = qp_lo(n, lb, lbA, r), hi = qp_hi(n, ub, ubA, r);
                int a = 0;
                if (lo > -LTI_INF_HALF && hi < LTI_INF_HALF && hi - lo <= 0.0) a = 2;
                else if (W.t[r] < lo) a = -1;
                else if (W.t[r] > hi) a = 1;
                if (r >= n && a != 0) {
                    double s = 0.0;
                    for (int j = 0; j < n; ++j) s += fabs(LTI_CM(A, m, r - n, j));
                    if (s == 0.0) a = 0;
                }
                W.act[r] = a;
            }
            g.sync();
            if (qp_active_solve(n, m, A, lb, ub, lbA, ubA, W, u, g) &&
                qp_check_update(n, m, A, lb, ub, lbA, ubA, W, u, tol, false, g)) { status = 0; break; }
        }
    }
    if (status != 0) {
        for (int i = g.tid(); i < n; i += g.size()) u[i] = W.ut[i];
        g.sync();
    }
    if (iters_out) *iters_out = iters;
    return status;
}

}  // namespace mpcb200
