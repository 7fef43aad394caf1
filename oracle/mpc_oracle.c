/* mpc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see mpc_oracle.h for scope and citations).
 * Plain C11, FP64, column-major matrices.  No product code may link this file. */
#include "mpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#define CM(M, ld, i, j) ((M)[(size_t)(i) + (size_t)(ld) * (size_t)(j)])

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* C(m x n) = A(m x k) * B(k x n), all column-major, C must not alias A/B */
static void gemm(int m, int n, int k, const double *A, const double *B, double *C) {
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) CM(C, m, i, j) = 0.0;
        for (int l = 0; l < k; ++l) {
            double b = CM(B, k, l, j);
            if (b == 0.0) continue;
            for (int i = 0; i < m; ++i) CM(C, m, i, j) += CM(A, m, i, l) * b;
        }
    }
}
/* dense gemm without the zero skip: used where the reference multiplies through zero-filled
 * matrices (src/QPSolver.cpp:58-60) so the CPU baseline pays the same work */
static void gemm_dense(int m, int n, int k, const double *A, const double *B, double *C) {
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) CM(C, m, i, j) = 0.0;
        for (int l = 0; l < k; ++l) {
            double b = CM(B, k, l, j);
            for (int i = 0; i < m; ++i) CM(C, m, i, j) += CM(A, m, i, l) * b;
        }
    }
}
/* C(n x m) = A(k x n)^T * B(k x m) */
static void gemm_tn(int n, int m, int k, const double *A, const double *B, double *C) {
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += CM(A, k, l, i) * CM(B, k, l, j);
            CM(C, n, i, j) = s;
        }
}

/* solve A X = B in place (LU, partial pivoting); A n x n destroyed, B n x m overwritten */
static int lu_solve(int n, int m, double *A, double *B) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(CM(A, n, k, k));
        for (int i = k + 1; i < n; ++i)
            if (fabs(CM(A, n, i, k)) > best) { best = fabs(CM(A, n, i, k)); p = i; }
        if (best == 0.0) return 1;
        if (p != k) {
            for (int j = 0; j < n; ++j) { double t = CM(A, n, k, j); CM(A, n, k, j) = CM(A, n, p, j); CM(A, n, p, j) = t; }
            for (int j = 0; j < m; ++j) { double t = CM(B, n, k, j); CM(B, n, k, j) = CM(B, n, p, j); CM(B, n, p, j) = t; }
        }
        double inv = 1.0 / CM(A, n, k, k);
        for (int i = k + 1; i < n; ++i) {
            double l = CM(A, n, i, k) * inv;
            if (l == 0.0) continue;
            CM(A, n, i, k) = l;
            for (int j = k + 1; j < n; ++j) CM(A, n, i, j) -= l * CM(A, n, k, j);
            for (int j = 0; j < m; ++j) CM(B, n, i, j) -= l * CM(B, n, k, j);
        }
    }
    for (int j = 0; j < m; ++j)
        for (int i = n - 1; i >= 0; --i) {
            double s = CM(B, n, i, j);
            for (int l = i + 1; l < n; ++l) s -= CM(A, n, i, l) * CM(B, n, l, j);
            CM(B, n, i, j) = s / CM(A, n, i, i);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ expm */

static void axpy_mat(int nn, double a, const double *X, double *Y) {
    for (int i = 0; i < nn; ++i) Y[i] += a * X[i];
}
static void add_identity(int n, double a, double *Y) {
    for (int i = 0; i < n; ++i) CM(Y, n, i, i) += a;
}

void orc_expm(int n, const double *Ain, double *E) {
    static const double b3[] = {120., 60., 12., 1.};
    static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
    static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                                2162160., 110880., 3960., 90., 1.};
    static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600.,
                                 1187353796428800., 129060195264000., 10559470521600.,
                                 670442572800., 33522128640., 1323241920., 40840800., 960960.,
                                 16380., 182., 1.};
    const int nn = n * n;
    double *A = dalloc(nn), *A2 = dalloc(nn), *A4 = dalloc(nn), *A6 = dalloc(nn), *A8 = dalloc(nn);
    double *U = dalloc(nn), *V = dalloc(nn), *T = dalloc(nn), *W = dalloc(nn);
    memcpy(A, Ain, sizeof(double) * nn);
    double l1 = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fabs(CM(A, n, i, j));
        if (s > l1) l1 = s;
    }
    int squarings = 0;
    if (l1 < 1.495585217958292e-002) {
        gemm(n, n, n, A, A, A2);
        memset(T, 0, sizeof(double) * nn); axpy_mat(nn, b3[3], A2, T); add_identity(n, b3[1], T);
        gemm(n, n, n, A, T, U);
        memset(V, 0, sizeof(double) * nn); axpy_mat(nn, b3[2], A2, V); add_identity(n, b3[0], V);
    } else if (l1 < 2.539398330063230e-001) {
        gemm(n, n, n, A, A, A2); gemm(n, n, n, A2, A2, A4);
        memset(T, 0, sizeof(double) * nn);
        axpy_mat(nn, b5[5], A4, T); axpy_mat(nn, b5[3], A2, T); add_identity(n, b5[1], T);
        gemm(n, n, n, A, T, U);
        memset(V, 0, sizeof(double) * nn);
        axpy_mat(nn, b5[4], A4, V); axpy_mat(nn, b5[2], A2, V); add_identity(n, b5[0], V);
    } else if (l1 < 9.504178996162932e-001) {
        gemm(n, n, n, A, A, A2); gemm(n, n, n, A2, A2, A4); gemm(n, n, n, A4, A2, A6);
        memset(T, 0, sizeof(double) * nn);
        axpy_mat(nn, b7[7], A6, T); axpy_mat(nn, b7[5], A4, T); axpy_mat(nn, b7[3], A2, T); add_identity(n, b7[1], T);
        gemm(n, n, n, A, T, U);
        memset(V, 0, sizeof(double) * nn);
        axpy_mat(nn, b7[6], A6, V); axpy_mat(nn, b7[4], A4, V); axpy_mat(nn, b7[2], A2, V); add_identity(n, b7[0], V);
    } else if (l1 < 2.097847961257068e+000) {
        gemm(n, n, n, A, A, A2); gemm(n, n, n, A2, A2, A4); gemm(n, n, n, A4, A2, A6); gemm(n, n, n, A6, A2, A8);
        memset(T, 0, sizeof(double) * nn);
        axpy_mat(nn, b9[9], A8, T); axpy_mat(nn, b9[7], A6, T); axpy_mat(nn, b9[5], A4, T);
        axpy_mat(nn, b9[3], A2, T); add_identity(n, b9[1], T);
        gemm(n, n, n, A, T, U);
        memset(V, 0, sizeof(double) * nn);
        axpy_mat(nn, b9[8], A8, V); axpy_mat(nn, b9[6], A6, V); axpy_mat(nn, b9[4], A4, V);
        axpy_mat(nn, b9[2], A2, V); add_identity(n, b9[0], V);
    } else {
        const double maxnorm = 5.371920351148152;
        frexp(l1 / maxnorm, &squarings);
        if (squarings < 0) squarings = 0;
        double sc = ldexp(1.0, -squarings);
        for (int i = 0; i < nn; ++i) A[i] *= sc;
        gemm(n, n, n, A, A, A2); gemm(n, n, n, A2, A2, A4); gemm(n, n, n, A4, A2, A6);
        /* U = A * (A6*(b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I) */
        memset(W, 0, sizeof(double) * nn);
        axpy_mat(nn, b13[13], A6, W); axpy_mat(nn, b13[11], A4, W); axpy_mat(nn, b13[9], A2, W);
        gemm(n, n, n, A6, W, T);
        axpy_mat(nn, b13[7], A6, T); axpy_mat(nn, b13[5], A4, T); axpy_mat(nn, b13[3], A2, T); add_identity(n, b13[1], T);
        gemm(n, n, n, A, T, U);
        /* V = A6*(b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I */
        memset(W, 0, sizeof(double) * nn);
        axpy_mat(nn, b13[12], A6, W); axpy_mat(nn, b13[10], A4, W); axpy_mat(nn, b13[8], A2, W);
        gemm(n, n, n, A6, W, V);
        axpy_mat(nn, b13[6], A6, V); axpy_mat(nn, b13[4], A4, V); axpy_mat(nn, b13[2], A2, V); add_identity(n, b13[0], V);
    }
    /* E = (V - U)^-1 (V + U) */
    for (int i = 0; i < nn; ++i) { T[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    lu_solve(n, n, T, E);
    for (int s = 0; s < squarings; ++s) { gemm(n, n, n, E, E, T); memcpy(E, T, sizeof(double) * nn); }
    free(A); free(A2); free(A4); free(A6); free(A8); free(U); free(V); free(T); free(W);
}

void orc_matpow(int n, const double *A, int k, double *P) {
    const int nn = n * n;
    double *base = dalloc(nn), *tmp = dalloc(nn);
    memcpy(base, A, sizeof(double) * nn);
    memset(P, 0, sizeof(double) * nn);
    for (int i = 0; i < n; ++i) CM(P, n, i, i) = 1.0;
    while (k > 0) {
        if (k & 1) { gemm(n, n, n, P, base, tmp); memcpy(P, tmp, sizeof(double) * nn); }
        k >>= 1;
        if (k) { gemm(n, n, n, base, base, tmp); memcpy(base, tmp, sizeof(double) * nn); }
    }
    free(base); free(tmp);
}

/* ------------------------------------------------------------------------------ QPSolver restated */

void orc_discretize(int NX, int NU, const double *Ac, const double *Bc, double Ts, double *Ad, double *Bd) {
    /* src/QPSolver.cpp:21-29 */
    const int m = NX + NU;
    double *M = dalloc((size_t)m * m), *E = dalloc((size_t)m * m);
    for (int j = 0; j < NX; ++j)
        for (int i = 0; i < NX; ++i) CM(M, m, i, j) = CM(Ac, NX, i, j) * Ts;
    for (int j = 0; j < NU; ++j)
        for (int i = 0; i < NX; ++i) CM(M, m, i, NX + j) = CM(Bc, NX, i, j) * Ts;
    orc_expm(m, M, E);
    for (int j = 0; j < NX; ++j)
        for (int i = 0; i < NX; ++i) CM(Ad, NX, i, j) = CM(E, m, i, j);
    for (int j = 0; j < NU; ++j)
        for (int i = 0; i < NX; ++i) CM(Bd, NX, i, j) = CM(E, m, i, NX + j);
    free(M); free(E);
}

/* shared tail of buildQPParams: cost from prediction matrices, dense product order of
 * src/QPSolver.cpp:50-60 */
static void cost_from_prediction(int NX, int NU, int N, const double *A_aug, const double *B_aug,
                                 const double *Q, const double *R, const double *P,
                                 const double *xi0, const double *xi_ref, double *H, double *f) {
    const int p = NX * (N + 1), n = NU * N;
    double *Qb = dalloc((size_t)p * p), *QB = dalloc((size_t)p * n);
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < NX; ++c)
            for (int r = 0; r < NX; ++r) CM(Qb, p, i * NX + r, i * NX + c) = CM(Q, NX, r, c);
    for (int c = 0; c < NX; ++c)
        for (int r = 0; r < NX; ++r) CM(Qb, p, N * NX + r, N * NX + c) = CM(P, NX, r, c);
    gemm_dense(p, n, p, Qb, B_aug, QB); /* Q_bar * B_aug */
    if (H) {
        gemm_tn(n, n, p, B_aug, QB, H); /* B_aug' (Q_bar B_aug) */
        for (int i = 0; i < N; ++i)
            for (int c = 0; c < NU; ++c)
                for (int r = 0; r < NU; ++r) CM(H, n, i * NU + r, i * NU + c) += CM(R, NU, r, c);
        for (size_t i = 0; i < (size_t)n * n; ++i) H[i] *= 2.0;
    }
    if (f) {
        double *e = dalloc(p);
        for (int i = 0; i < p; ++i) {
            double s = 0.0;
            for (int j = 0; j < NX; ++j) s += CM(A_aug, p, i, j) * xi0[j];
            e[i] = s - xi_ref[i]; /* column-major NX x (N+1) flatten == step-major stacking */
        }
        /* f = 2 B' Q_bar e  ==  2 (Q_bar B)' e  (Q_bar symmetric) */
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int i = 0; i < p; ++i) s += CM(QB, p, i, j) * e[i];
            f[j] = 2.0 * s;
        }
        free(e);
    }
    free(Qb); free(QB);
}

void orc_build_qp_params(int NX, int NU, int N, const double *Ad, const double *Bd,
                         const double *Q, const double *R, const double *P,
                         const double *x_min, const double *x_max, double u_min, double u_max,
                         const double *xi0, const double *xi_ref,
                         double *A_aug_o, double *B_aug_o, double *H, double *f,
                         double *A_eq, double *b_eq, double *lb, double *ub,
                         double *A_ineq, double *lbA, double *ubA) {
    const int p = NX * (N + 1), n = NU * N;
    double *A_aug = dalloc((size_t)p * NX), *B_aug = dalloc((size_t)p * n);
    double *Pw = dalloc((size_t)NX * NX), *blk = dalloc((size_t)NX * (NX > NU ? NX : NU));
    /* src/QPSolver.cpp:36-40 */
    for (int i = 0; i < NX; ++i) CM(A_aug, p, i, i) = 1.0;
    for (int i = 1; i <= N; ++i) {
        for (int c = 0; c < NX; ++c)
            for (int r = 0; r < NX; ++r) {
                double s = 0.0;
                for (int l = 0; l < NX; ++l) s += CM(Ad, NX, r, l) * CM(A_aug, p, (i - 1) * NX + l, c);
                CM(A_aug, p, i * NX + r, c) = s;
            }
    }
    /* src/QPSolver.cpp:42-47: one Ad.pow() per block, as the reference does */
    for (int i = 1; i <= N; ++i)
        for (int j = 0; j < i; ++j) {
            orc_matpow(NX, Ad, i - j - 1, Pw);
            gemm(NX, NU, NX, Pw, Bd, blk);
            for (int c = 0; c < NU; ++c)
                for (int r = 0; r < NX; ++r) CM(B_aug, p, i * NX + r, j * NU + c) = CM(blk, NX, r, c);
        }
    cost_from_prediction(NX, NU, N, A_aug, B_aug, Q, R, P, xi0, xi_ref, H, f);
    /* src/QPSolver.cpp:62-64 (kept for completeness; known-spurious block) */
    if (A_eq)
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < NX * N; ++r) CM(A_eq, NX * N, r, c) = CM(B_aug, p, NX + r, c);
    if (b_eq)
        for (int r = 0; r < NX * N; ++r) {
            double s = 0.0;
            for (int j = 0; j < NX; ++j) s += CM(A_aug, p, NX + r, j) * xi0[j];
            b_eq[r] = s;
        }
    /* :66-68 */
    if (lb) for (int i = 0; i < n; ++i) lb[i] = u_min;
    if (ub) for (int i = 0; i < n; ++i) ub[i] = u_max;
    /* :70-80 */
    const int mi = 2 * NX * N;
    if (A_ineq) memset(A_ineq, 0, sizeof(double) * (size_t)mi * n);
    if (lbA) for (int i = 0; i < mi; ++i) lbA[i] = -ORC_INFTY;
    if (ubA) for (int i = 0; i < mi; ++i) ubA[i] = ORC_INFTY;
    for (int i = 0; i < N; ++i) {
        orc_matpow(NX, Ad, i + 1, Pw);
        if (A_ineq)
            for (int c = 0; c < n; ++c)
                for (int r = 0; r < NX; ++r) CM(A_ineq, mi, 2 * i * NX + r, c) = CM(B_aug, p, (i + 1) * NX + r, c);
        for (int r = 0; r < NX; ++r) {
            double s = 0.0;
            for (int j = 0; j < NX; ++j) s += CM(Pw, NX, r, j) * xi0[j];
            if (lbA) lbA[2 * i * NX + r] = x_min[r] - s;
            if (ubA) ubA[2 * i * NX + r] = x_max[r] - s;
        }
    }
    if (A_aug_o) memcpy(A_aug_o, A_aug, sizeof(double) * (size_t)p * NX);
    if (B_aug_o) memcpy(B_aug_o, B_aug, sizeof(double) * (size_t)p * n);
    free(A_aug); free(B_aug); free(Pw); free(blk);
}

void orc_update_state(int NX, int NU, const double *Ad, const double *Bd, double *xi, const double *u) {
    double tmp[64];
    double *t = NX <= 64 ? tmp : dalloc(NX);
    for (int r = 0; r < NX; ++r) {
        double s = 0.0;
        for (int j = 0; j < NX; ++j) s += CM(Ad, NX, r, j) * xi[j];
        for (int j = 0; j < NU; ++j) s += CM(Bd, NX, r, j) * u[j];
        t[r] = s;
    }
    memcpy(xi, t, sizeof(double) * NX);
    if (t != tmp) free(t);
}

/* --------------------------------------------------------------- dual active-set QP (Goldfarb-Idnani) */

typedef struct {
    int kind;   /* 0: bound on variable idx, 1: row idx of A */
    int idx;
    double sgn; /* constraint is  sgn * (c' u) >= sgn*b  with c = e_idx or A[idx,:] */
    double b;   /* right-hand side in the ORIGINAL orientation (lower or upper value) */
    int eq;
} gi_con;

typedef struct {
    int n, mA;
    const double *A;
} gi_ctx;

static double con_eval(const gi_ctx *c, const gi_con *k, const double *x) {
    double v;
    if (k->kind == 0) v = x[k->idx];
    else {
        v = 0.0;
        for (int j = 0; j < c->n; ++j) v += CM(c->A, c->mA, k->idx, j) * x[j];
    }
    return k->sgn * (v - k->b);
}
static void con_normal(const gi_ctx *c, const gi_con *k, double *np) {
    if (k->kind == 0) {
        memset(np, 0, sizeof(double) * c->n);
        np[k->idx] = k->sgn;
    } else
        for (int j = 0; j < c->n; ++j) np[j] = k->sgn * CM(c->A, c->mA, k->idx, j);
}

static void givens(double a, double b, double *c, double *s) {
    if (b == 0.0) { *c = 1.0; *s = 0.0; return; }
    double h = hypot(a, b);
    *c = a / h; *s = b / h;
}

int orc_qp_solve(int n, const double *H, const double *f, int mA, const double *A,
                 const double *lbA, const double *ubA, const double *lb, const double *ub,
                 double *x, double *y_bnd, double *y_row, int *iters_out) {
    const double BIG = ORC_INFTY * 0.5;
    gi_ctx ctx = {n, mA, A};
    int maxc = 2 * (n + mA);
    gi_con *cons = (gi_con *)calloc(maxc ? maxc : 1, sizeof(gi_con));
    int nc = 0;
    /* equalities first (lb==ub), then inequalities */
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < n; ++i) {
            double lo = lb ? lb[i] : -ORC_INFTY, hi = ub ? ub[i] : ORC_INFTY;
            int is_eq = (lo > -BIG && hi < BIG && hi - lo <= 0.0);
            if (pass == 0) { if (is_eq) cons[nc++] = (gi_con){0, i, 1.0, lo, 1}; }
            else if (!is_eq) {
                if (lo > -BIG) cons[nc++] = (gi_con){0, i, 1.0, lo, 0};
                if (hi < BIG) cons[nc++] = (gi_con){0, i, -1.0, hi, 0};
            }
        }
        for (int r = 0; r < mA; ++r) {
            int nz = 0;
            for (int j = 0; j < n && !nz; ++j) nz = CM(A, mA, r, j) != 0.0;
            if (!nz) continue;
            double lo = lbA[r], hi = ubA[r];
            int is_eq = (lo > -BIG && hi < BIG && hi - lo <= 0.0);
            if (pass == 0) { if (is_eq) cons[nc++] = (gi_con){1, r, 1.0, lo, 1}; }
            else if (!is_eq) {
                if (lo > -BIG) cons[nc++] = (gi_con){1, r, 1.0, lo, 0};
                if (hi < BIG) cons[nc++] = (gi_con){1, r, -1.0, hi, 0};
            }
        }
    }
    int status = 0, iters = 0;
    double *L = dalloc((size_t)n * n), *J = dalloc((size_t)n * n), *R = dalloc((size_t)n * n);
    double *d = dalloc(n), *z = dalloc(n), *r = dalloc(n), *np = dalloc(n), *u = dalloc(n + 1);
    int *act = (int *)calloc(n + 1, sizeof(int));
    char *is_act = (char *)calloc(nc ? nc : 1, 1);
    int q = 0;

    /* Cholesky H = L L' */
    for (int j = 0; j < n; ++j) {
        double s = CM(H, n, j, j);
        for (int k = 0; k < j; ++k) s -= CM(L, n, j, k) * CM(L, n, j, k);
        if (!(s > 0.0)) { status = 2; goto done; }
        double dj = sqrt(s);
        CM(L, n, j, j) = dj;
        for (int i = j + 1; i < n; ++i) {
            double t = CM(H, n, i, j);
            for (int k = 0; k < j; ++k) t -= CM(L, n, i, k) * CM(L, n, j, k);
            CM(L, n, i, j) = t / dj;
        }
    }
    /* J = L^-T  (column j of J solves L' J[:,j] = e_j) */
    for (int j = 0; j < n; ++j) {
        for (int i = n - 1; i >= 0; --i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = i + 1; k < n; ++k) s -= CM(L, n, k, i) * CM(J, n, k, j);
            CM(J, n, i, j) = s / CM(L, n, i, i);
        }
    }
    /* x = -H^-1 f */
    for (int i = 0; i < n; ++i) {
        double s = -f[i];
        for (int k = 0; k < i; ++k) s -= CM(L, n, i, k) * x[k];
        x[i] = s / CM(L, n, i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= CM(L, n, k, i) * x[k];
        x[i] = s / CM(L, n, i, i);
    }
    double scale = 1.0;
    for (int i = 0; i < n; ++i) if (fabs(x[i]) > scale) scale = fabs(x[i]);
    const double feas_tol = 1e-11;
    const int max_iters = 50 * (n + nc) + 100;
    int next_eq = 0;

    for (;;) {
        /* choose the constraint to add: pending equalities first, else most violated inequality */
        int ip = -1;
        double sp = 0.0;
        while (next_eq < nc && cons[next_eq].eq) {
            double s = con_eval(&ctx, &cons[next_eq], x);
            ip = next_eq++;
            sp = s;
            break;
        }
        if (ip < 0) {
            double worst = -feas_tol * scale;
            for (int k = 0; k < nc; ++k) {
                if (cons[k].eq || is_act[k]) continue;
                double s = con_eval(&ctx, &cons[k], x);
                double nrm = 1.0;
                if (cons[k].kind == 1) {
                    nrm = 0.0;
                    for (int j = 0; j < n; ++j) nrm += CM(A, mA, cons[k].idx, j) * CM(A, mA, cons[k].idx, j);
                    nrm = sqrt(nrm);
                }
                if (s < worst * nrm) { worst = s / nrm; ip = k; sp = s; }
            }
            if (ip < 0) break; /* optimal */
        }
        const int eq = cons[ip].eq;
        con_normal(&ctx, &cons[ip], np);
        if (eq && sp > 0.0) { /* orient the equality so that the step increases c'x */
            for (int j = 0; j < n; ++j) np[j] = -np[j];
            cons[ip].sgn = -cons[ip].sgn;
            sp = -sp;
        }
        double uplus = 0.0;
        for (;;) {
            if (++iters > max_iters) { status = 1; goto done; }
            /* d = J' n+ ; z = J2 d2 ; r = R^-1 d1 */
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
                for (int i = 0; i < n; ++i) s += CM(J, n, i, j) * np[i];
                d[j] = s;
            }
            for (int i = 0; i < n; ++i) {
                double s = 0.0;
                for (int j = q; j < n; ++j) s += CM(J, n, i, j) * d[j];
                z[i] = s;
            }
            for (int i = q - 1; i >= 0; --i) {
                double s = d[i];
                for (int j = i + 1; j < q; ++j) s -= CM(R, n, i, j) * r[j];
                r[i] = s / CM(R, n, i, i);
            }
            double znorm2 = 0.0, ztn = 0.0, npn = 0.0;
            for (int i = 0; i < n; ++i) { znorm2 += z[i] * z[i]; ztn += z[i] * np[i]; npn += np[i] * np[i]; }
            /* partial step length t1 and the blocking constraint l */
            double t1 = INFINITY; int l = -1;
            for (int k = 0; k < q; ++k) {
                if (cons[act[k]].eq) continue;
                if (r[k] > 0.0) {
                    double t = u[k] / r[k];
                    if (t < t1) { t1 = t; l = k; }
                }
            }
            int z_is_zero = !(znorm2 > 1e-24 * npn * (1.0 + 0.0) && fabs(ztn) > 1e-14 * npn);
            double t2 = z_is_zero ? INFINITY : -sp / ztn;
            if (eq && z_is_zero) {
                /* dependent equality: redundant if consistent, else infeasible */
                if (fabs(sp) <= 1e-9 * scale) break;
                if (l < 0) { status = 2; goto done; }
            }
            double t = t1 < t2 ? t1 : t2;
            if (!isfinite(t)) { status = 2; goto done; }
            if (!isfinite(t2)) {
                /* dual step only */
                for (int k = 0; k < q; ++k) u[k] -= t * r[k];
                uplus += t;
            } else {
                for (int i = 0; i < n; ++i) x[i] += t * z[i];
                for (int k = 0; k < q; ++k) u[k] -= t * r[k];
                uplus += t;
                if (t == t2) {
                    /* full step: add the constraint. Rotate d so that d[q+1..] = 0 */
                    for (int j = n - 1; j > q; --j) {
                        double c, s;
                        givens(d[j - 1], d[j], &c, &s);
                        double dn = c * d[j - 1] + s * d[j];
                        d[j - 1] = dn; d[j] = 0.0;
                        for (int i = 0; i < n; ++i) {
                            double a = CM(J, n, i, j - 1), b = CM(J, n, i, j);
                            CM(J, n, i, j - 1) = c * a + s * b;
                            CM(J, n, i, j) = -s * a + c * b;
                        }
                    }
                    for (int i = 0; i <= q; ++i) CM(R, n, i, q) = d[i];
                    act[q] = ip; u[q] = uplus; is_act[ip] = 1; ++q;
                    break;
                }
                sp = con_eval(&ctx, &cons[ip], x);
            }
            /* drop blocking constraint l */
            is_act[act[l]] = 0;
            for (int k = l; k < q - 1; ++k) {
                act[k] = act[k + 1]; u[k] = u[k + 1];
                for (int i = 0; i <= k + 1; ++i) CM(R, n, i, k) = CM(R, n, i, k + 1);
            }
            --q;
            for (int k = l; k < q; ++k) {
                double c, s;
                givens(CM(R, n, k, k), CM(R, n, k + 1, k), &c, &s);
                for (int j = k; j < q; ++j) {
                    double a = CM(R, n, k, j), b = CM(R, n, k + 1, j);
                    CM(R, n, k, j) = c * a + s * b;
                    CM(R, n, k + 1, j) = -s * a + c * b;
                }
                for (int i = 0; i < n; ++i) {
                    double a = CM(J, n, i, k), b = CM(J, n, i, k + 1);
                    CM(J, n, i, k) = c * a + s * b;
                    CM(J, n, i, k + 1) = -s * a + c * b;
                }
            }
        }
        for (int i = 0; i < n; ++i) if (fabs(x[i]) > scale) scale = fabs(x[i]);
    }
done:
    /* fixed variables (lb == ub) are returned exactly at their bound, as qpOASES reports them */
    if (status == 0 && lb && ub)
        for (int i = 0; i < n; ++i)
            if (lb[i] > -BIG && ub[i] < BIG && ub[i] - lb[i] <= 0.0) x[i] = lb[i];
    if (y_bnd) memset(y_bnd, 0, sizeof(double) * n);
    if (y_row && mA) memset(y_row, 0, sizeof(double) * mA);
    for (int k = 0; k < q; ++k) {
        const gi_con *c = &cons[act[k]];
        double yv = c->sgn * u[k];
        if (c->kind == 0) { if (y_bnd) y_bnd[c->idx] += yv; }
        else if (y_row) y_row[c->idx] += yv;
    }
    if (iters_out) *iters_out = iters;
    free(cons); free(L); free(J); free(R); free(d); free(z); free(r); free(np); free(u); free(act); free(is_act);
    return status;
}

void orc_kkt_residual(int n, const double *H, const double *f, int mA, const double *A,
                      const double *lbA, const double *ubA, const double *lb, const double *ub,
                      const double *u, const double *y_bnd, const double *y_row, double res[4]) {
    const double BIG = ORC_INFTY * 0.5;
    double stat = 0.0, prim = 0.0, dual = 0.0, comp = 0.0;
    for (int i = 0; i < n; ++i) {
        double g = f[i];
        for (int j = 0; j < n; ++j) g += CM(H, n, i, j) * u[j];
        g -= y_bnd ? y_bnd[i] : 0.0;
        for (int r = 0; r < mA; ++r) g -= CM(A, mA, r, i) * (y_row ? y_row[r] : 0.0);
        if (fabs(g) > stat) stat = fabs(g);
    }
    for (int pass = 0; pass < 2; ++pass) {
        int cnt = pass == 0 ? n : mA;
        for (int k = 0; k < cnt; ++k) {
            double v, lo, hi, y;
            if (pass == 0) { v = u[k]; lo = lb ? lb[k] : -ORC_INFTY; hi = ub ? ub[k] : ORC_INFTY; y = y_bnd ? y_bnd[k] : 0.0; }
            else {
                v = 0.0;
                for (int j = 0; j < n; ++j) v += CM(A, mA, k, j) * u[j];
                lo = lbA[k]; hi = ubA[k]; y = y_row ? y_row[k] : 0.0;
            }
            int has_lo = lo > -BIG, has_hi = hi < BIG;
            if (has_lo && lo - v > prim) prim = lo - v;
            if (has_hi && v - hi > prim) prim = v - hi;
            int eq = has_lo && has_hi && hi - lo <= 0.0;
            if (!eq) {
                if (y > 0.0) { /* lower active */
                    if (!has_lo) { if (y > dual) dual = y; }
                    else if (fabs(y * (v - lo)) > comp) comp = fabs(y * (v - lo));
                } else if (y < 0.0) {
                    if (!has_hi) { if (-y > dual) dual = -y; }
                    else if (fabs(y * (hi - v)) > comp) comp = fabs(y * (hi - v));
                }
            }
        }
    }
    res[0] = stat; res[1] = prim; res[2] = dual; res[3] = comp;
}

/* ----------------------------------------------------------------------------------------- gait */

void orc_gait_defaults(orc_gait_params *g) {
    g->dt = 0.001f;       /* include/MPCParam.h:44 */
    g->mpc_step = 5;      /* :46 */
    g->swing_time = 0.5f; /* :48 */
    g->stance_time = 0.5f;/* :49 */
}

void orc_calculate_gait(const orc_gait_params *g, int iter, int *left_leg_state, int *right_leg_state,
                        double *phase_o, double *remain_o) {
    /* include/MPCController.h:61-75.  `iter * param.dt` is int*float evaluated in float and
     * widened on assignment; `swing_time + stance_time` is a float add widened likewise.
     * volatile blocks FMA contraction / constant folding surprises. */
    volatile float ct = (float)iter * g->dt;
    volatile float cy = g->swing_time + g->stance_time;
    double currentTime = (double)ct;
    double cycleTime = (double)cy;
    double phase = fmod(currentTime, cycleTime);
    int l, r; double rem;
    if (phase < (double)g->swing_time) { l = 1; r = 0; rem = (double)g->swing_time - phase; }
    else { l = 0; r = 1; rem = cycleTime - phase; }
    if (left_leg_state) *left_leg_state = l;
    if (right_leg_state) *right_leg_state = r;
    if (phase_o) *phase_o = phase;
    if (remain_o) *remain_o = rem;
}

void orc_contact_schedule(const orc_gait_params *g, int iter, int N, uint8_t *contact) {
    for (int k = 0; k < N; ++k) {
        if (iter < 0) { contact[2 * k] = 1; contact[2 * k + 1] = 1; continue; }
        int l, r;
        orc_calculate_gait(g, iter + k * g->mpc_step, &l, &r, 0, 0);
        contact[2 * k] = (uint8_t)(l == 0);
        contact[2 * k + 1] = (uint8_t)(r == 0);
    }
}

/* ---------------------------------------------------------------------------------------- TRON1 */

void orc_tron1_defaults(orc_tron1_params *p) {
    static const double q[13] = {1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1};
    static const double I[9] = {140110.479E-06, 534.939E-06, 28184.116E-06,
                                534.939E-06, 110641.449E-06, -27.278E-06,
                                28184.116E-06, -27.278E-06, 98944.542E-06};
    p->Ts = 0.005;      /* dtMPC = dt*mpcStep, include/MPCParam.h:47 */
    p->mass = 9.585;
    memcpy(p->inertia, I, sizeof(I));
    memcpy(p->q, q, sizeof(q));
    p->r = 0.1;
    p->p_scale = 20.0;
    p->mu = 0.5;
    p->f_max = 2.0 * 9.585 * 9.8;
    p->ltv = 1;
    p->per_step_feet = 0;
}

static void inv3(const double *M, double *Mi) { /* column-major 3x3 */
    double a = M[0], b = M[3], c = M[6], d = M[1], e = M[4], f = M[7], g = M[2], h = M[5], i = M[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    double id = 1.0 / det;
    Mi[0] = (e * i - f * h) * id; Mi[3] = (c * h - b * i) * id; Mi[6] = (b * f - c * e) * id;
    Mi[1] = (f * g - d * i) * id; Mi[4] = (a * i - c * g) * id; Mi[7] = (c * d - a * f) * id;
    Mi[2] = (d * h - e * g) * id; Mi[5] = (b * g - a * h) * id; Mi[8] = (a * e - b * d) * id;
}

void orc_tron1_model(const orc_tron1_params *p, double yaw, const double pos[3], const double feet[6],
                     double *Ac, double *Bc) {
    memset(Ac, 0, sizeof(double) * 169);
    memset(Bc, 0, sizeof(double) * 78);
    double c = cos(yaw), s = sin(yaw);
    double Rz[9] = {c, s, 0, -s, c, 0, 0, 0, 1}; /* column-major [[c,-s,0],[s,c,0],[0,0,1]] */
    /* Theta_dot = Rz' omega */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) CM(Ac, 13, i, 6 + j) = CM(Rz, 3, j, i);
    for (int i = 0; i < 3; ++i) CM(Ac, 13, 3 + i, 9 + i) = 1.0;
    CM(Ac, 13, 11, 12) = 1.0; /* v_z_dot = ... + g, with the state g = -9.8 (include/mpcQP.h:71) */
    double Ii[9], T[9], Iwi[9];
    inv3(p->inertia, Ii);
    gemm(3, 3, 3, Rz, Ii, T);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += CM(T, 3, i, k) * CM(Rz, 3, j, k);
            CM(Iwi, 3, i, j) = v;
        }
    for (int ft = 0; ft < 2; ++ft) {
        double r[3] = {feet[3 * ft] - pos[0], feet[3 * ft + 1] - pos[1], feet[3 * ft + 2] - pos[2]};
        double S[9] = {0, r[2], -r[1], -r[2], 0, r[0], r[1], -r[0], 0}; /* column-major skew */
        double W[9];
        gemm(3, 3, 3, Iwi, S, W);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) CM(Bc, 13, 6 + i, 3 * ft + j) = CM(W, 3, i, j);
        for (int i = 0; i < 3; ++i) CM(Bc, 13, 9 + i, 3 * ft + i) = 1.0 / p->mass;
    }
}

void orc_tron1_model_literal(const orc_tron1_params *p, const double pos[3], const double foot[3],
                             double *Ac, double *Bc) {
    /* include/mpcQP.h:139-181 as written */
    double dx = foot[0] - pos[0], dy = foot[1] - pos[1], dz = foot[2] - pos[2];
    memset(Ac, 0, sizeof(double) * 169);
    memset(Bc, 0, sizeof(double) * 39);
    CM(Ac, 13, 0, 7) = dz; CM(Ac, 13, 0, 8) = dy;
    CM(Ac, 13, 1, 6) = dz; CM(Ac, 13, 1, 8) = dx;
    CM(Ac, 13, 2, 6) = dy; CM(Ac, 13, 2, 7) = dx;
    CM(Ac, 13, 3, 9) = 1; CM(Ac, 13, 4, 10) = 1; CM(Ac, 13, 5, 11) = 1;
    CM(Ac, 13, 11, 12) = -1;
    CM(Bc, 13, 9, 0) = -p->mass; CM(Bc, 13, 10, 1) = -p->mass; CM(Bc, 13, 11, 2) = -p->mass;
}

void orc_tron1_reference(const double x0[13], int N, double Ts, double omega_yaw, double velocity_x,
                         double *x_ref) {
    /* include/mpcQP.h:74-97 */
    for (int i = 0; i <= N; ++i) {
        double t = i * Ts;
        double *c = x_ref + 13 * i;
        memcpy(c, x0, sizeof(double) * 13);
        c[2] = x0[2] + t * omega_yaw;
        c[3] = x0[3] + t * velocity_x;
        c[9] = (i == 0) ? x0[9] : velocity_x;
        c[12] = -9.8;
    }
}

void orc_tron1_condense(const orc_tron1_params *p, int N, const double *x0, const double *x_ref,
                        const double *feet, double *A_aug_o, double *B_aug_o, double *H, double *f) {
    const int NX = 13, NU = 6, pp = NX * (N + 1), n = NU * N;
    double *A_aug = dalloc((size_t)pp * NX), *B_aug = dalloc((size_t)pp * n);
    double *Ads = dalloc((size_t)169 * N), *Bds = dalloc((size_t)78 * N);
    double Ac[169], Bc[78], Q[169], R[36], P[169], T[169], T2[169];
    const int nmodel = p->ltv ? N : 1;
    for (int k = 0; k < nmodel; ++k) {
        const double *lin = (k == 0) ? x0 : x_ref + 13 * k;
        const double *fk = (p->per_step_feet && p->ltv) ? feet + 6 * k : feet;
        orc_tron1_model(p, lin[2], lin + 3, fk, Ac, Bc);
        orc_discretize(NX, NU, Ac, Bc, p->Ts, Ads + 169 * k, Bds + 78 * k);
    }
    for (int k = nmodel; k < N; ++k) {
        memcpy(Ads + 169 * k, Ads, sizeof(double) * 169);
        memcpy(Bds + 78 * k, Bds, sizeof(double) * 78);
    }
    for (int i = 0; i < NX; ++i) CM(A_aug, pp, i, i) = 1.0;
    for (int i = 1; i <= N; ++i) {
        const double *Ad = Ads + 169 * (i - 1);
        for (int c = 0; c < NX; ++c)
            for (int r = 0; r < NX; ++r) {
                double s = 0.0;
                for (int l = 0; l < NX; ++l) s += CM(Ad, NX, r, l) * CM(A_aug, pp, (i - 1) * NX + l, c);
                CM(A_aug, pp, i * NX + r, c) = s;
            }
    }
    for (int i = 1; i <= N; ++i)
        for (int j = 0; j < i; ++j) {
            if (!p->ltv) {
                orc_matpow(NX, Ads, i - j - 1, T); /* src/QPSolver.cpp:45 */
            } else {
                memset(T, 0, sizeof(T));
                for (int d = 0; d < NX; ++d) CM(T, NX, d, d) = 1.0;
                for (int k = j + 1; k < i; ++k) { gemm(NX, NX, NX, Ads + 169 * k, T, T2); memcpy(T, T2, sizeof(T)); }
            }
            double blk[78];
            gemm(NX, NU, NX, T, Bds + 78 * j, blk);
            for (int c = 0; c < NU; ++c)
                for (int r = 0; r < NX; ++r) CM(B_aug, pp, i * NX + r, j * NU + c) = CM(blk, NX, r, c);
        }
    memset(Q, 0, sizeof(Q)); memset(P, 0, sizeof(P)); memset(R, 0, sizeof(R));
    for (int i = 0; i < NX; ++i) { CM(Q, NX, i, i) = p->q[i]; CM(P, NX, i, i) = p->p_scale * p->q[i]; }
    for (int i = 0; i < NU; ++i) CM(R, NU, i, i) = p->r;
    if (H || f) cost_from_prediction(NX, NU, N, A_aug, B_aug, Q, R, P, x0, x_ref, H, f);
    if (A_aug_o) memcpy(A_aug_o, A_aug, sizeof(double) * (size_t)pp * NX);
    if (B_aug_o) memcpy(B_aug_o, B_aug, sizeof(double) * (size_t)pp * n);
    free(A_aug); free(B_aug); free(Ads); free(Bds);
}

void orc_tron1_constraints(const orc_tron1_params *p, int N, const uint8_t *contact,
                           double *A, double *lbA, double *ubA, double *lb, double *ub) {
    const int n = 6 * N, m = 8 * N;
    memset(A, 0, sizeof(double) * (size_t)m * n);
    for (int k = 0; k < N; ++k)
        for (int ft = 0; ft < 2; ++ft) {
            int b = 6 * k + 3 * ft, r = 8 * k + 4 * ft;
            int c = contact[2 * k + ft] != 0;
            lb[b] = lb[b + 1] = c ? -ORC_INFTY : 0.0;
            ub[b] = ub[b + 1] = c ? ORC_INFTY : 0.0;
            lb[b + 2] = 0.0;
            ub[b + 2] = c ? p->f_max : 0.0;
            CM(A, m, r, b + 2) = p->mu; CM(A, m, r, b) = -1.0;
            CM(A, m, r + 1, b + 2) = p->mu; CM(A, m, r + 1, b) = 1.0;
            CM(A, m, r + 2, b + 2) = p->mu; CM(A, m, r + 2, b + 1) = -1.0;
            CM(A, m, r + 3, b + 2) = p->mu; CM(A, m, r + 3, b + 1) = 1.0;
            for (int q = 0; q < 4; ++q) { lbA[r + q] = 0.0; ubA[r + q] = ORC_INFTY; }
        }
}

/* exact Euclidean projection of v onto {|x|<=mu z, |y|<=mu z, 0<=z<=fmax} */
static void project_pyramid(double mu, double fmax, const double v[3], double out[3]) {
    double ax = fabs(v[0]), ay = fabs(v[1]);
    double a = ax > ay ? ax : ay, b = ax > ay ? ay : ax, w = v[2], t;
    if (mu * w >= a) t = w;
    else {
        t = (w + mu * a) / (1.0 + mu * mu);
        if (mu * t < b) t = (w + mu * (a + b)) / (1.0 + 2.0 * mu * mu);
    }
    if (t > fmax) t = fmax;
    if (t < 0.0) t = 0.0;
    double lim = mu * t;
    out[0] = v[0] > lim ? lim : (v[0] < -lim ? -lim : v[0]);
    out[1] = v[1] > lim ? lim : (v[1] < -lim ? -lim : v[1]);
    out[2] = t;
}

double orc_tron1_natural_residual(const orc_tron1_params *p, int N, const double *H, const double *f,
                                  const uint8_t *contact, const double *u) {
    const int n = 6 * N;
    double worst = 0.0;
    double *g = dalloc(n);
    for (int i = 0; i < n; ++i) {
        double s = f[i];
        for (int j = 0; j < n; ++j) s += CM(H, n, i, j) * u[j];
        g[i] = s;
    }
    for (int s = 0; s < 2 * N; ++s) {
        double v[3] = {u[3 * s] - g[3 * s], u[3 * s + 1] - g[3 * s + 1], u[3 * s + 2] - g[3 * s + 2]};
        double o[3] = {0, 0, 0};
        if (contact[s]) project_pyramid(p->mu, p->f_max, v, o);
        for (int c = 0; c < 3; ++c) {
            double e = fabs(u[3 * s + c] - o[c]);
            if (e > worst) worst = e;
        }
    }
    free(g);
    return worst;
}

int orc_tron1_solve(const orc_tron1_params *p, int N, const double *x0, const double *x_ref,
                    const double *feet, const uint8_t *contact, double *forces, int *iters) {
    const int n = 6 * N, m = 8 * N;
    double *H = dalloc((size_t)n * n), *f = dalloc(n), *A = dalloc((size_t)m * n);
    double *lbA = dalloc(m), *ubA = dalloc(m), *lb = dalloc(n), *ub = dalloc(n);
    orc_tron1_condense(p, N, x0, x_ref, feet, 0, 0, H, f);
    orc_tron1_constraints(p, N, contact, A, lbA, ubA, lb, ub);
    int st = orc_qp_solve(n, H, f, m, A, lbA, ubA, lb, ub, forces, 0, 0, iters);
    free(H); free(f); free(A); free(lbA); free(ubA); free(lb); free(ub);
    return st;
}

int orc_tron1_rollout(const orc_tron1_params *p, const orc_gait_params *g, int N, int steps, double *x,
                      double omega_yaw, double velocity_x, int iter0, const double off_l[3],
                      const double off_r[3], double *u_traj) {
    int bad = 0;
    double *xr = dalloc((size_t)13 * (N + 1)), *U = dalloc((size_t)6 * N);
    uint8_t *contact = (uint8_t *)calloc((size_t)2 * N, 1);
    orc_tron1_params q = *p;
    q.per_step_feet = 0;
    for (int s = 0; s < steps; ++s) {
        double c = cos(x[2]), sn = sin(x[2]), feet[6], Ac[169], Bc[78], Ad[169], Bd[78];
        feet[0] = x[3] + c * off_l[0] - sn * off_l[1]; feet[1] = x[4] + sn * off_l[0] + c * off_l[1]; feet[2] = 0.0;
        feet[3] = x[3] + c * off_r[0] - sn * off_r[1]; feet[4] = x[4] + sn * off_r[0] + c * off_r[1]; feet[5] = 0.0;
        orc_tron1_reference(x, N, q.Ts, omega_yaw, velocity_x, xr);
        orc_contact_schedule(g, iter0 < 0 ? iter0 : iter0 + s * g->mpc_step, N, contact);
        int it = 0;
        if (orc_tron1_solve(&q, N, x, xr, feet, contact, U, &it)) ++bad;
        if (u_traj) memcpy(u_traj + (size_t)6 * s, U, sizeof(double) * 6);
        orc_tron1_model(&q, x[2], x + 3, feet, Ac, Bc);
        orc_discretize(13, 6, Ac, Bc, q.Ts, Ad, Bd);
        orc_update_state(13, 6, Ad, Bd, x, U);
    }
    free(xr); free(U); free(contact);
    return bad;
}

typedef struct {
    const orc_tron1_params *p;
    int N, B, tid, nthreads;
    const double *x0, *x_ref, *feet;
    const uint8_t *contact;
    double *forces;
    int32_t *status, *iters;
    int bad;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    const int N = j->N;
    /* one solve per host core, the worker pinned to its core (SURVEY.md section 8d / BASELINE.md section 3); on a box with
     * fewer allowed CPUs than threads the call fails harmlessly and the scheduler places the thread */
    {
        cpu_set_t allowed, one;
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            int n = CPU_COUNT(&allowed), want = n > 0 ? j->tid % n : 0, seen = 0;
            for (int c = 0; c < CPU_SETSIZE; ++c) {
                if (!CPU_ISSET(c, &allowed)) continue;
                if (seen++ == want) { CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof(one), &one); break; }
            }
        }
    }
    const size_t fstride = (j->p->per_step_feet ? (size_t)6 * N : 6);
    for (int b = j->tid; b < j->B; b += j->nthreads) {
        int it = 0;
        int st = orc_tron1_solve(j->p, N, j->x0 + (size_t)13 * b, j->x_ref + (size_t)13 * (N + 1) * b,
                                 j->feet + fstride * b, j->contact + (size_t)2 * N * b,
                                 j->forces + (size_t)6 * N * b, &it);
        if (j->status) j->status[b] = st;
        if (j->iters) j->iters[b] = it;
        if (st) j->bad++;
    }
    return 0;
}

int orc_tron1_solve_batch(const orc_tron1_params *p, int N, int B, const double *x0, const double *x_ref,
                          const double *feet, const uint8_t *contact, double *forces,
                          int32_t *status, int32_t *iters, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (batch_job){p, N, B, t, nthreads, x0, x_ref, feet, contact, forces, status, iters, 0};
        pthread_create(&th[t], 0, batch_worker, &jobs[t]);
    }
    int bad = 0;
    for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], 0); bad += jobs[t].bad; }
    return bad;
}
