// tron1_core.cuh -- per-instance TRON1 convex-MPC algorithm (linearise -> discretise -> condense ->
// QP solve), written once as group-cooperative code.
//
// A "group" is the set of threads that cooperates on ONE robot instance: a warp or a few warps
// on the GPU (GrpCuda in mpc_b200.cu), a single serial thread in the test-only host build
// (tests/emul/).  Every phase is a strided loop `for (w = g.tid(); w < n; w += g.size())`
// followed by g.sync(), so the same source is valid for any group size.
//
// What it replaces in the reference (paths relative to /root/reference):
//   include/mpcQP.h:121-182   buildSystemModel      -> model_step()        (intended SRBD physics)
//   src/QPSolver.cpp:21-29    discretizeSystem      -> closed-form ZOH of the nilpotent SRBD model
//   src/QPSolver.cpp:36-60    buildQPParams H, f    -> build_hessian(), adjoint()  (B_aug never formed)
//   src/QPSolver.cpp:83-106   solveQP (qpOASES)     -> active-face Newton iterations on a packed
//                                                      Cholesky + ADMM fallback, exact projection
//   include/MPCController.h:61-75 calculateGait     -> gait_contact()
//
// Structure that is exploited: with x = [Theta, p, omega, v, g],
//   A_c^3 = 0 and A_c^2 B_c = 0, so  A_d = I + Ts A_c + Ts^2/2 A_c^2,  B_d = Ts B_c + Ts^2/2 A_c B_c
// exactly, and block (i,j) of B_aug (i > j) is
//   Theta rows: Ts^2 S(i,j) W_j     S(i,j) = 1/2 Rz_j' + sum_{j<k<i} Rz_k'
//   p rows    : Ts^2/m (i-j-1/2) I
//   omega rows: Ts W_j              W_j = Iw_j^-1 [r_j]x  (per foot)
//   v rows    : Ts/m I
// so H = 2(B'QB + R) and f = 2B'Q(A x0 - x_ref) reduce to suffix sums over the horizon.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

namespace mpcb200 {

struct Tron1Const {
    double Ts, inv_m;
    double Iinv[9];   // inverse body inertia, row-major (symmetric)
    double q[13];     // diag(Q)  (reference include/mpcQP.h:54)
    double r;         // R = r I  (include/mpcQP.h:55)
    double p_scale;   // P = p_scale Q (include/mpcQP.h:56)
    double mu, f_max; // friction pyramid, normal-force cap
    double tol;       // natural-residual tolerance (relative to max(1,|u|_inf))
    double gamma;     // step of the natural map used for face prediction
    double admm_alpha;
    int ltv, per_step_feet;
    int max_newton, max_admm;
    float gait_dt, gait_swing, gait_stance;
    int gait_mpc_step;
    double gait_cycle, gait_inv_cycle;     // (double)(swing + stance) with the float add of the reference, and its reciprocal
    double foot_off_l[3], foot_off_r[3];   // nominal base->foot offsets (include/MPCParam.h:64-73)
};

enum { ST_SOLVED = 0, ST_MAXITER = 1, ST_FAILED = 2, ST_DEFER = 3 };   // ST_DEFER: internal (Riccati work type -> dense class), never reported

// ------------------------------------------------------------------------------------------------
// gait: bit-exact restatement of MPC::calculateGait (include/MPCController.h:61-75) with the float
// members of MPCParam (include/MPCParam.h:44-49).  `iter * dt` is an int*float product rounded to
// float, then widened; `swing + stance` is a float add.  No FMA contraction can occur (single ops).
MPC_HD void gait_contact(const Tron1Const& P, int iter, int& left_stance, int& right_stance) {
    if (iter < 0) { left_stance = 1; right_stance = 1; return; }   // standing
#if defined(__CUDA_ARCH__)
    float ct = __fmul_rn((float)iter, P.gait_dt);
#else
    volatile float ctv = (float)iter * P.gait_dt;
    float ct = ctv;
#endif
    // phase = fmod((double)ct, (double)cycle), evaluated EXACTLY without the library's iterative fmod:
    // x = (double)ct and y = (double)cycle carry 24-bit significands and q = trunc(x / y) < 2^29, so q y is
    // exact in double, r = fma(-q, y, x) is exact, and a quotient that is off by one (x * (1/y) is rounded)
    // is repaired by one correction step.  The result is the mathematically exact remainder, which is
    // what fmod returns (checked against the C library on the reference's disagreement cases in the tests).
    const double x = (double)ct, y = P.gait_cycle;
    double q = trunc(x * P.gait_inv_cycle);
    double phase = fma(-q, y, x);
    if (phase < 0.0) phase += y;
    else if (phase >= y) phase -= y;
    if (!(y > 0.0) || !(x >= 0.0)) phase = fmod(x, y);   // degenerate parameters: defer to the library
    int left_swing = phase < (double)P.gait_swing;
    left_stance = !left_swing;
    right_stance = left_swing;
}

// ------------------------------------------------------------------------------------------------
// N_ = horizon, NC_ = capacity in compact decision variables (3 per stance foot-step).
// NC_ = 3N serves gaits with one stance foot per step (the reference's alternating gait),
// NC_ = 6N serves double support; the kernel wrapper routes instances by their actual size.
// AINL_ = the packed factor lives inside the struct (shared memory); false = the wrapper points S.A at
// external storage (global memory) -- needed when (NC+1)(NC+2)/2 doubles exceed shared memory (N = 50
// double support: 364 KB).
// TILED_ = the matrix is stored as 8x8 tiles of the lower triangle (tile (I,J) at (I(I+1)/2+J)*64, inside a tile two
// 8x4 halves [half][row][4], so that the fragments of an FP64 tensor-core MMA (DMMA.8x8x4) are 32 consecutive doubles)
// and factorised by the tiled right-looking Cholesky with DMMA trailing updates (chol_tiled, horizon 50).  The
// right-hand side is then row NC (fixed), rows nc..NC-1 are identity padding, and after the factorisation the diagonal
// tiles hold inv(L_KK) instead of L_KK.
// RIC_ = the active-face solves run as a Riccati recursion over the horizon (riccati_face_solve) instead of factorising the
// condensed Hessian: no matrix storage at all (the struct shrinks to the per-step gains), O(N) work per solve.  Instances
// whose active-face iteration does not certify are handed to a dense-class work type (ST_DEFER).
// storage that only the Riccati work type carries (an empty base for the others: their layout, tuned to the shared-memory
// budget of four CTAs per SM, stays exactly as it was)
template <bool RIC_, int RC_TOTAL_>
struct RicStore {
    alignas(16) double rc[RC_TOTAL_];   // gains (when inside the struct), exchange buffers of one step
    double* adjx;                       // 18 (N + 1) doubles of adjoint scratch provided by the caller
};
template <int RC_TOTAL_>
struct RicStore<false, RC_TOTAL_> {};
template <int N, int NC, bool AINL>
struct RicLayout {
    // force-space gain rows [Ku (12) | ku0] (6 per step, when inside the struct) and the exchange buffers of one step
    // (riccati_backward_step / riccati_forward_step)
    static constexpr int RC_K = 0, RC_XM = RC_K + (AINL ? NC * 13 : 0), RC_U6 = RC_XM + 120,
                         RC_GH = RC_U6 + 48, RC_KT = RC_GH + 48, RC_WQ = RC_KT + 78, RC_D = RC_WQ + 12,
                         RC_UK = RC_D + 24, RC_KF = RC_UK + 6, RC_TOTAL = RC_KF + (AINL ? 0 : 4 * 78);
};

template <int N_, int NC_, bool AINL_ = true, bool TILED_ = false, bool RIC_ = false>
struct Tron1Work : RicStore<RIC_, RicLayout<N_, NC_, AINL_>::RC_TOTAL> {
    static constexpr int N = N_;
    static constexpr int NC = NC_;
    static constexpr bool AINL = AINL_;
    static constexpr bool TILED = TILED_;
    static constexpr bool RICCATI = RIC_;
    static_assert(!(RIC_ && TILED_), "the Riccati work type stores no matrix");
    static_assert(!RIC_ || NC_ == 6 * N_, "the Riccati work type keeps six gain rows per step");
    static constexpr int NS = 2 * N;     // foot-steps
    static constexpr int NV = 6 * N;     // decision variables (full layout)
    static constexpr int PKN = (NC + 1) * (NC + 2) / 2;  // packed lower triangle incl. rhs row
    static constexpr int NT = (NC + 8) / 8;              // tile rows covering rows 0..NC (row NC = right-hand side)
    static constexpr int TSZ = NT * (NT + 1) / 2 * 64;   // tiled lower triangle
    static constexpr bool UEXT = RIC_ && !AINL_;         // Riccati work type with external gains: the forces live behind the gains in the same slab
    static constexpr int ASZ = RIC_ ? NC * 13 + (UEXT ? 6 * N : 0) : (TILED ? TSZ : PKN);   // doubles of matrix (Riccati: gain + force) storage per instance
    double* Aext;           // external factor storage (only used when !AINL)
    alignas(16) double Astore[(AINL && !RIC_) ? ASZ : 2];   // reduced Hessian / Cholesky factor; the rhs is row nc (packed) or row NC (tiled)
    alignas(16) double pbuf[TILED ? NT * 64 + 128 : 2];  // tiled: current panel column + two inverse diagonal tiles (double buffer)
    double xt[TILED ? 8 * NT : 2], st[TILED ? 8 * NT : 2];   // tiled triangular solves: solution blocks, running sums
    double dinv[RIC_ ? 2 : NC];        // 1 / L_kk
    static constexpr int CBS = (NC + 4) & ~1;   // stride of one broadcast buffer (even: 16-byte aligned halves)
    alignas(16) double colbuf[RIC_ ? 2 : 2 * CBS];   // double-buffered broadcast copy of the current pivot column
    double w[RIC_ ? 2 : NC], z[RIC_ ? 2 : NC], y[RIC_ ? 2 : NC];   // compact solve vector, ADMM iterates
    using RL = RicLayout<N_, NC_, AINL_>;
    static constexpr int RC_K = RL::RC_K, RC_XM = RL::RC_XM, RC_U6 = RL::RC_U6, RC_GH = RL::RC_GH, RC_KT = RL::RC_KT,
                         RC_WQ = RL::RC_WQ, RC_D = RL::RC_D, RC_UK = RL::RC_UK, RC_KF = RL::RC_KF, RC_TOTAL = RL::RC_TOTAL;
    double W[N * 18];       // W[k][foot] 3x3 row-major:  Iw_k^-1 [r]x
    double cs[N * 2];       // cos, sin of yaw_k
    double cc[N + 1], ss[N + 1];   // prefix sums  sum_{k<i} cos / sin
    double dc[RIC_ ? 1 : N], ds[RIC_ ? 1 : N];    // D_j = 1/2 Rz_j' - C_{j+1}  (cos-like / sin-like entries; the Riccati work type forms them on the fly)
    double SW[RIC_ ? 8 : N * 8];       // suffix sums over i>j of w_i * {1, cc, ss, cc^2, ss^2, cc ss, i, i^2}
    // Riccati work type: f is never formed (the gradient is the adjoint of the full tracking error), adj lives in the
    // caller's dead input staging area (adjx), g and res overlay the exchange buffers of the sweeps, and ee holds the
    // free-response error e0 until the forward sweep adds the input response in place (a further active-face iteration
    // rebuilds e0 from the re-staged inputs)
    static constexpr int RC_X = RC_TOTAL - RC_XM;                 // doubles of exchange buffers
    static constexpr bool GALIAS = RIC_ && (8 * N <= RC_X);
    double f[RIC_ ? 2 : NV];             // full layout: linear term (live for the whole solve)
    // ee | adj | g are contiguous and dead between the right-hand side of a face solve and the optimality check:
    // the 3x3-block elimination uses them as its broadcast arrays (gj3_solve_regs); u must survive (swing feet stay 0)
    double ee[(N + 1) * 12];  // tracking error: free response during setup, input response later
    double adj[RIC_ ? 2 : (N + 1) * 18]; // adjoint terms / suffix sums; first 6(N+1) doubles double as `tau`
    double g[GALIAS ? 2 : NV], u[UEXT ? 2 : NV];      // full layout: gradient, solution
    double res[GALIAS ? 2 : NS];
    const double* x0;       // 13 doubles (staged by the caller)
    const double* feet;     // 6 or 6N doubles
    int8_t contact[NS], ax[NS], ay[NS], zt[NS], nax[NS], nay[NS], nzt[NS];
    int16_t cidx[NS];       // foot-step -> compact stance index
    int8_t cinv[NS];        // compact stance index -> foot-step (2 step + foot)
    int nc;         // compact variable count (3 * stance foot-steps)
    int interior;   // group-uniform: the face just solved was the interior face of every stance foot-step (Z = I, nothing fixed)
    int flag;       // group-uniform scratch flag
#if defined(MPC_PHASE_TIMING)
    long long prof[16];
    long long t_last;
#endif
    MPC_HD double* adjp() { if constexpr (RIC_) return this->adjx; else return adj; }
    MPC_HD double* gp() { if constexpr (GALIAS) return this->rc + RC_XM; else return g; }
    MPC_HD double* resp() { if constexpr (GALIAS) return this->rc + RC_XM + NV; else return res; }
    MPC_HD double* up() { if constexpr (UEXT) return Aext + NC * 13; else return u; }   // forces (full layout)
    MPC_HD double* tau() { return adjp(); }
    MPC_HD double* Kp() { if constexpr (AINL) return this->rc + RC_K; else return Aext; }   // Riccati gains: inside the struct or external (global memory)
    // address of the packed factor: a compile-time offset for the shared-memory case, a pointer otherwise
    MPC_HD double* Ap() { if constexpr (AINL) return Astore; else return Aext; }
    // index of entry (i, j), i >= j, of the lower triangle
    static MPC_HD int pk(int i, int j) {
        if constexpr (TILED) {
            const int I = i >> 3, J = j >> 3;
            return ((I * (I + 1) / 2 + J) << 6) + ((j & 4) << 3) + ((i & 7) << 2) + (j & 3);
        } else {
            return i * (i + 1) / 2 + j;
        }
    }
    MPC_HD int rhs_row() const { return TILED ? NC : nc; }
};

// finer ticks inside the Riccati sweeps (profiling builds: -DMPC_RIC_TICKS=1 backward phases, =2 forward phases)
#if defined(MPC_RIC_TICKS)
#define MPC_RTICK(which, S, g, id) do { if (MPC_RIC_TICKS == (which)) MPC_TICK(S, g, id); } while (0)
#else
#define MPC_RTICK(which, S, g, id) do { } while (0)
#endif
#define MPC_PK(i, j) ((i) * ((i) + 1) / 2 + (j))

// optional per-phase cycle accounting (profiling build only: -DMPC_PHASE_TIMING, tools/phase_timing.py)
#if defined(MPC_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define MPC_TICK(S, g, id)                                         \
    do {                                                           \
        long long t_ = clock64();                                  \
        if ((g).tid() == 0) { (S).prof[id] += t_ - (S).t_last; (S).t_last = t_; } \
    } while (0)
#else
#define MPC_TICK(S, g, id) do { } while (0)
#endif

// ---- phase: model of horizon step k  (include/mpcQP.h:121-182, intended physics) -----------------
template <class WK>
MPC_HD void model_step(const Tron1Const& P, WK& S, const double* xref, int k) {
    [[maybe_unused]] constexpr int N = WK::N;
    const double* lin = (k == 0 || !P.ltv) ? S.x0 : (xref + 13 * k);
    double s, c;
    sincos(lin[2], &s, &c);
    S.cs[2 * k] = c; S.cs[2 * k + 1] = s;
    // Iw^-1 = Rz Ib^-1 Rz'
    const double* I = P.Iinv;
    double T[9];  // T = Rz * Iinv
    for (int j = 0; j < 3; ++j) {
        T[0 * 3 + j] = c * I[0 * 3 + j] - s * I[1 * 3 + j];
        T[1 * 3 + j] = s * I[0 * 3 + j] + c * I[1 * 3 + j];
        T[2 * 3 + j] = I[2 * 3 + j];
    }
    double Iw[9];  // Iw = T * Rz'
    for (int i = 0; i < 3; ++i) {
        Iw[i * 3 + 0] = T[i * 3 + 0] * c - T[i * 3 + 1] * s;
        Iw[i * 3 + 1] = T[i * 3 + 0] * s + T[i * 3 + 1] * c;
        Iw[i * 3 + 2] = T[i * 3 + 2];
    }
    const double* ft = S.feet + ((P.per_step_feet && P.ltv) ? 6 * k : 0);
    for (int a = 0; a < 2; ++a) {
        double rx = ft[3 * a] - lin[3], ry = ft[3 * a + 1] - lin[4], rz = ft[3 * a + 2] - lin[5];
        double* Wk = S.W + 18 * k + 9 * a;
        // [r]x = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]]
        for (int i = 0; i < 3; ++i) {
            double i0 = Iw[i * 3], i1 = Iw[i * 3 + 1], i2 = Iw[i * 3 + 2];
            Wk[i * 3 + 0] = i1 * rz - i2 * ry;
            Wk[i * 3 + 1] = -i0 * rz + i2 * rx;
            Wk[i * 3 + 2] = i0 * ry - i1 * rx;
        }
    }
}

// weights of prediction step i (1..N): Q for i<N, P = p_scale Q for i==N (src/QPSolver.cpp:50-56)
template <int N>
MPC_HD double step_weight(const Tron1Const& P, int i) { return i == N ? P.p_scale : 1.0; }

// ---- serial scans over the horizon, one lane per component ------------------------------------------------
// The values are first pulled into registers (independent loads, pipelined), scanned there (one dependent add
// per step) and written back: an in-place `acc += a[j]; a[j] = acc` loop serialises a shared-memory load, an
// add and a store per step (~50 cycles each).  Same order of additions as the plain loop, so bit-identical.
template <int N, int STRIDE>
MPC_HD void suffix_scan_inplace(double* a) {          // a[j * STRIDE] <- sum_{i >= j} a[i * STRIDE]
    constexpr int CH = (N % 10 == 0) ? 10 : N;
    double acc = 0.0;
    for (int c0 = N - CH; c0 >= 0; c0 -= CH) {
        double v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = a[(c0 + j) * STRIDE];
#pragma unroll
        for (int j = CH - 1; j >= 0; --j) { acc += v[j]; v[j] = acc; }
#pragma unroll
        for (int j = 0; j < CH; ++j) a[(c0 + j) * STRIDE] = v[j];
    }
}
// out[(k + 1) * OSTRIDE] = sum_{m <= k} scale * in[m * ISTRIDE], out[0] = 0   (exclusive prefix, N + 1 outputs)
template <int N, int ISTRIDE, int OSTRIDE>
MPC_HD void prefix_scan(const double* in, double* out, double scale, bool scaled) {
    constexpr int CH = (N % 10 == 0) ? 10 : N;
    double acc = 0.0;
    out[0] = 0.0;
    for (int c0 = 0; c0 < N; c0 += CH) {
        double v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = in[(c0 + j) * ISTRIDE];
#pragma unroll
        for (int j = 0; j < CH; ++j) { acc += scaled ? scale * v[j] : v[j]; v[j] = acc; }
#pragma unroll
        for (int j = 0; j < CH; ++j) out[(c0 + j + 1) * OSTRIDE] = v[j];
    }
}

// ---- phase: prefix / suffix sums that depend on the yaw sequence only ---------------------------
template <class WK, class G>
MPC_HD void horizon_sums(const Tron1Const& P, WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    // prefix cc_i, ss_i (two serial scans, one per thread)
    for (int c = g.tid(); c < 2; c += g.size()) prefix_scan<N, 2, 1>(S.cs + c, c ? S.ss : S.cc, 1.0, false);
    g.sync();
    for (int j = g.tid(); j < N; j += g.size()) {
        if constexpr (!WK::RICCATI) {
            S.dc[j] = 0.5 * S.cs[2 * j] - S.cc[j + 1];
            S.ds[j] = 0.5 * S.cs[2 * j + 1] - S.ss[j + 1];
        }
    }
    if constexpr (WK::RICCATI) { g.sync(); return; }   // the suffix sums below feed build_hessian only
    // per-step terms w_i * {1, cc_i, ss_i, cc_i^2, ss_i^2, cc_i ss_i, i, i^2} (parallel over i = 1..N) ...
    for (int i = 1 + g.tid(); i <= N; i += g.size()) {
        double w = step_weight<N>(P, i), cc = S.cc[i], ss = S.ss[i], di = (double)i;
        double* t = S.SW + 8 * (i - 1);
        t[0] = w; t[1] = w * cc; t[2] = w * ss; t[3] = w * cc * cc; t[4] = w * ss * ss; t[5] = w * cc * ss;
        t[6] = w * di; t[7] = w * di * di;
    }
    g.sync();
    // ... then suffix sums over i > j, one component per thread (uniform code, no divergence)
    for (int c = g.tid(); c < 8; c += g.size()) suffix_scan_inplace<N, 8>(S.SW + c);
    g.sync();
}

// ---- phase: free-response error  e0_i = (A_aug x0 - x_ref)_i  for i = 0..N ----------------------
template <class WK, class G>
MPC_HD void free_response(const Tron1Const& P, WK& S, const double* xref, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const double Ts = P.Ts;
    for (int i = g.tid(); i <= N; i += g.size()) {
        const double* x0 = S.x0;
        const double* xr = xref + 13 * i;
        double* e = S.ee + 12 * i;
        double cc = S.cc[i], ss = S.ss[i], di = (double)i;
        double gz = x0[12];
        e[0] = x0[0] + Ts * (cc * x0[6] + ss * x0[7]) - xr[0];
        e[1] = x0[1] + Ts * (-ss * x0[6] + cc * x0[7]) - xr[1];
        e[2] = x0[2] + Ts * (di * x0[8]) - xr[2];
        e[3] = x0[3] + di * Ts * x0[9] - xr[3];
        e[4] = x0[4] + di * Ts * x0[10] - xr[4];
        e[5] = x0[5] + di * Ts * x0[11] + 0.5 * di * di * Ts * Ts * gz - xr[5];
        e[6] = x0[6] - xr[6];
        e[7] = x0[7] - xr[7];
        e[8] = x0[8] - xr[8];
        e[9] = x0[9] - xr[9];
        e[10] = x0[10] - xr[10];
        e[11] = x0[11] + di * Ts * gz - xr[11];
    }
    g.sync();
}

// ---- phase: input response  ee_i = (B_aug u)_i  for i = 0..N (u in full layout) -----------------
template <class WK, class G>
MPC_HD void input_response(const Tron1Const& P, WK& S, const double* u, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const double Ts = P.Ts;
    // per step: tau_k = sum_a W_ka u_ka ; phi_k = sum_a u_ka / m
    for (int k = g.tid(); k < N; k += g.size()) {
        const double* W0 = S.W + 18 * k;
        const double* uk = u + 6 * k;
        double* t = S.tau() + 6 * k;
        for (int i = 0; i < 3; ++i) {
            t[i] = W0[i * 3] * uk[0] + W0[i * 3 + 1] * uk[1] + W0[i * 3 + 2] * uk[2]
                 + W0[9 + i * 3] * uk[3] + W0[9 + i * 3 + 1] * uk[4] + W0[9 + i * 3 + 2] * uk[5];
            t[3 + i] = (uk[i] + uk[3 + i]) * P.inv_m;
        }
    }
    g.sync();
    // level 1: omega_k = Ts sum_{m<k} tau_m, v_k = Ts sum_{m<k} phi_m  (one component per thread, uniform code)
    for (int c = g.tid(); c < 6; c += g.size()) prefix_scan<N, 6, 12>(S.tau() + c, S.ee + 6 + c, Ts, true);   // omega -> 6..8, v -> 9..11
    g.sync();
    // level 2: per-step increments  dTheta_k = Ts Rz_k'(omega_k + Ts/2 tau_k),  dp_k = Ts (v_k + Ts/2 phi_k)
    for (int k = g.tid(); k < N; k += g.size()) {
        double* t = S.tau() + 6 * k;
        const double* e = S.ee + 12 * k;
        double cz = S.cs[2 * k], sz = S.cs[2 * k + 1];
        double a0 = e[6] + 0.5 * Ts * t[0], a1 = e[7] + 0.5 * Ts * t[1], a2 = e[8] + 0.5 * Ts * t[2];
        double b0 = e[9] + 0.5 * Ts * t[3], b1 = e[10] + 0.5 * Ts * t[4], b2 = e[11] + 0.5 * Ts * t[5];
        t[0] = Ts * (cz * a0 + sz * a1);
        t[1] = Ts * (-sz * a0 + cz * a1);
        t[2] = Ts * a2;
        t[3] = Ts * b0; t[4] = Ts * b1; t[5] = Ts * b2;
    }
    g.sync();
    // level 3: Theta_k, p_k = prefix sums of the increments
    for (int c = g.tid(); c < 6; c += g.size()) prefix_scan<N, 6, 12>(S.tau() + c, S.ee + c, 1.0, false);       // Theta -> 0..2, p -> 3..5
    g.sync();
}

// ---- phase: adjoint  out = 2 B_aug' Qbar e   (e: (N+1) x 12, out: full layout 6N) ----------------
template <class WK, class G>
MPC_HD void adjoint(const Tron1Const& P, WK& S, const double* e, double* out, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const double Ts = P.Ts;
    const double* q = P.q;
    // weighted terms of step i (row i-1 of adj holds step i, i = 1..N)
    for (int i = 1 + g.tid(); i <= N; i += g.size()) {
        const double* ei = e + 12 * i;
        double w = step_weight<N>(P, i), cc = S.cc[i], ss = S.ss[i], di = (double)i;
        double* a = S.adjp() + 18 * (i - 1);
        double t0 = w * q[0] * ei[0], t1 = w * q[1] * ei[1], t2 = w * q[2] * ei[2];
        a[0] = cc * t0 - ss * t1;   // C_i' Q e_Theta, x
        a[1] = ss * t0 + cc * t1;   //               y
        a[2] = di * t2;             //               z
        a[3] = t0; a[4] = t1; a[5] = t2;
        for (int c = 0; c < 3; ++c) {
            a[6 + c] = w * q[6 + c] * ei[6 + c];            // omega
            double tp = w * q[3 + c] * ei[3 + c];
            a[9 + c] = di * tp;                             // i * Qp e_p
            a[12 + c] = tp;                                 // Qp e_p
            a[15 + c] = w * q[9 + c] * ei[9 + c];           // Qv e_v
        }
    }
    g.sync();
    // suffix sums over i > j, stored at row j  (row j currently holds step j+1)
    for (int c = g.tid(); c < 18; c += g.size()) suffix_scan_inplace<N, 18>(S.adjp() + c);
    g.sync();
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        int j = s >> 1;
        const double* a = S.adjp() + 18 * j;
        const double* Wj = S.W + 9 * s;
        double dc, ds;
        if constexpr (WK::RICCATI) { dc = 0.5 * S.cs[2 * j] - S.cc[j + 1]; ds = 0.5 * S.cs[2 * j + 1] - S.ss[j + 1]; }
        else { dc = S.dc[j]; ds = S.ds[j]; }
        const double jh = (double)j + 0.5;
        // sum_i S(i,j)' Q e_Theta
        double vt0 = a[0] + dc * a[3] - ds * a[4];
        double vt1 = a[1] + ds * a[3] + dc * a[4];
        double vt2 = a[2] - jh * a[5];
        double m0 = Ts * (Ts * vt0 + a[6]), m1 = Ts * (Ts * vt1 + a[7]), m2 = Ts * (Ts * vt2 + a[8]);
        for (int c = 0; c < 3; ++c) {
            double rot = Wj[0 * 3 + c] * m0 + Wj[1 * 3 + c] * m1 + Wj[2 * 3 + c] * m2;   // W' m
            double lin = Ts * P.inv_m * (Ts * (a[9 + c] - jh * a[12 + c]) + a[15 + c]);
            out[3 * s + c] = 2.0 * (rot + lin);
        }
    }
    g.sync();
}

// ---- reduction basis of one foot-step face --------------------------------------------------------
struct FaceZ { double fx, fy, fz, mx, my; };
MPC_HD FaceZ face_basis(double mu, int ax, int ay, int zt) {
    FaceZ z;
    z.fz = (zt == 0) ? 1.0 : 0.0;
    z.fx = (ax == 0 && zt != 2) ? 1.0 : 0.0;
    z.fy = (ay == 0 && zt != 2) ? 1.0 : 0.0;
    z.mx = (double)ax * mu;
    z.my = (double)ay * mu;
    return z;
}

// ---- phase: packed reduced Hessian  A = Z'(H + rho I)Z  over the stance variables ----------------
// H = 2(B'QB + R) (src/QPSolver.cpp:58) restricted to stance foot-steps; eliminated variables get
// a unit diagonal.  Work item = one pair of horizon steps (j >= l).
template <class WK, class G>
MPC_HD void build_hessian(const Tron1Const& P, WK& S, double rho, bool use_face, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const double Ts2 = P.Ts * P.Ts, Ts4 = Ts2 * Ts2, im2 = P.inv_m * P.inv_m;
    const double* q = P.q;
    // work item = one pair of STANCE foot-steps (compact indices ia >= ib) = one 3x3 block of the reduced
    // Hessian: every lane runs the same straight-line code whatever the contact pattern (a loop over step
    // pairs with inner loops over the feet in contact made the lanes of a warp diverge four ways)
    const int m = S.nc / 3;
    for (int pr = g.tid(); pr < m * (m + 1) / 2; pr += g.size()) {
        // unrank pr -> (ia, ib), ia >= ib
        int ia = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (ia * (ia + 1) / 2 > pr) --ia;
        while ((ia + 1) * (ia + 2) / 2 <= pr) ++ia;
        const int ib = pr - ia * (ia + 1) / 2;
        const int sa = S.cinv[ia], sb = S.cinv[ib];       // foot-steps (2 step + foot), sa >= sb
        const int j = sa >> 1, l = sb >> 1;               // horizon steps, j >= l
        const double* sw = S.SW + 8 * j;   // suffix sums over i > j  (j >= l so i > l too)
        double dcj = S.dc[j], dsj = S.ds[j], dcl = S.dc[l], dsl = S.ds[l];
        double Saa = sw[3] + (dcj + dcl) * sw[1] + dcj * dcl * sw[0];
        double Sbb = sw[4] + (dsj + dsl) * sw[2] + dsj * dsl * sw[0];
        double Sab = sw[5] + dcj * sw[2] + dsl * sw[1] + dcj * dsl * sw[0];
        double Sba = sw[5] + dsj * sw[1] + dcl * sw[2] + dsj * dcl * sw[0];
        double jh = (double)j + 0.5, lh = (double)l + 0.5;
        double Szz = sw[7] - (jh + lh) * sw[6] + jh * lh * sw[0];   // sum w (i-j-1/2)(i-l-1/2)
        // M = Ts^4 sum S(i,j)' Q_Theta S(i,l) + Ts^2 sum Q_omega      (2x2 block + zz)
        double M00 = Ts4 * (q[0] * Saa + q[1] * Sbb) + Ts2 * q[6] * sw[0];
        double M01 = Ts4 * (q[0] * Sab - q[1] * Sba);
        double M10 = Ts4 * (q[0] * Sba - q[1] * Sab);
        double M11 = Ts4 * (q[0] * Sbb + q[1] * Saa) + Ts2 * q[7] * sw[0];
        double M22 = Ts4 * q[2] * Szz + Ts2 * q[8] * sw[0];
        double Dp[3];
        for (int c = 0; c < 3; ++c) Dp[c] = im2 * (Ts4 * q[3 + c] * Szz + Ts2 * q[9 + c] * sw[0]);
        const double* Wa = S.W + 9 * sa;
        const double* Wb = S.W + 9 * sb;
        // block = Wa' M Wb + D: first L = Wa' M (3x3), then T = L Wb
        double L[9];
        for (int r = 0; r < 3; ++r) {   // row r of Wa' = column r of Wa
            double w0 = Wa[0 * 3 + r], w1 = Wa[1 * 3 + r], w2 = Wa[2 * 3 + r];
            L[r * 3 + 0] = w0 * M00 + w1 * M10;
            L[r * 3 + 1] = w0 * M01 + w1 * M11;
            L[r * 3 + 2] = w2 * M22;
        }
        double T[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                T[r * 3 + c] = L[r * 3 + 0] * Wb[0 * 3 + c] + L[r * 3 + 1] * Wb[1 * 3 + c] + L[r * 3 + 2] * Wb[2 * 3 + c];
        for (int c = 0; c < 3; ++c) T[c * 3 + c] += Dp[c];
        if (sa == sb) for (int c = 0; c < 3; ++c) T[c * 3 + c] += P.r + 0.5 * rho;
        for (int i = 0; i < 9; ++i) T[i] *= 2.0;
        // interior faces have Z = I: the first active-face iteration of every instance skips the reduction
        if (use_face && (S.ax[sa] | S.ay[sa] | S.zt[sa] | S.ax[sb] | S.ay[sb] | S.zt[sb]) != 0) {
            FaceZ Za = face_basis(P.mu, S.ax[sa], S.ay[sa], S.zt[sa]);
            FaceZ Zb = face_basis(P.mu, S.ax[sb], S.ay[sb], S.zt[sb]);
            for (int r = 0; r < 3; ++r) {   // columns:  T <- T Zb
                double t0 = T[r * 3], t1 = T[r * 3 + 1], t2 = T[r * 3 + 2];
                T[r * 3 + 2] = Zb.fz * (Zb.mx * t0 + Zb.my * t1 + t2);
                T[r * 3 + 0] = Zb.fx * t0;
                T[r * 3 + 1] = Zb.fy * t1;
            }
            for (int c = 0; c < 3; ++c) {   // rows:  T <- Za' T
                double t0 = T[c], t1 = T[3 + c], t2 = T[6 + c];
                T[6 + c] = Za.fz * (Za.mx * t0 + Za.my * t1 + t2);
                T[c] = Za.fx * t0;
                T[3 + c] = Za.fy * t1;
            }
            if (sa == sb) {
                if (Za.fx == 0.0) T[0] = 1.0;
                if (Za.fy == 0.0) T[4] = 1.0;
                if (Za.fz == 0.0) T[8] = 1.0;
            }
        }
        const int ra = 3 * ia, cb = 3 * ib;
        double* Ap = S.Ap();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                if (sa == sb && c > r) continue;
                Ap[WK::pk(ra + r, cb + c)] = T[r * 3 + c];
            }
    }
    g.sync();
}

#if defined(__CUDA_ARCH__)
// ---- device fast path: right-looking Cholesky with ONE ROW PER THREAD held in registers ------------
// Thread t owns row t of [H; rhs'] (rows 0..nc, row nc is the rhs).  Column k: the pivot comes from
// its owner by warp shuffle (one-warp groups) or through shared memory, every row scales its entry
// with rsqrt, publishes it in the packed factor (which doubles as the broadcast buffer), and the
// trailing update reads column k by uniform-address (broadcast) loads.  Static indexing keeps the row
// in registers; `k < n` / `j < n` predicates are group-uniform.  Result layout == generic path:
// packed L in S.A, 1/L_kk in S.dinv, L^-1 rhs in row nc.
// FULL = (nc == NC): the common case, compiled without the per-element `j < n` predicates.
template <class WK, class G, bool FULL>
__device__ __forceinline__ bool cholesky_regs_body(WK& S, const G& g) {
    constexpr int NC = WK::NC;
    const int n = FULL ? NC : S.nc, t = g.tid();
    double* A = S.Ap();
    double a[NC];
    const bool row = t <= n;
    const double* mine = A + MPC_PK(t, 0);
#pragma unroll
    for (int j = 0; j < NC; ++j) a[j] = (row && (FULL || j < n) && (j <= t || t == n)) ? mine[j] : 0.0;
    bool ok = true;
    // `piv` carries the next pivot candidate: after column k every thread forms a[k+1] - l*l from its own
    // registers; only the owner of row k+1 holds the true value, and it is the one the shuffle reads.
    // This keeps the shared-memory round trip (STS -> sync -> LDS) off the pivot-to-pivot critical path.
    double piv = a[0];
    if (G::kThreads != 32) {   // multi-warp groups exchange the pivot through shared memory
        if (t == 0) S.w[0] = piv;
        g.sync();
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        if (FULL || k < n) {
            double d;
            if (G::kThreads == 32) d = __shfl_sync(0xffffffffu, piv, k);
            else d = S.w[k];
            if (!(d > 0.0)) { ok = false; d = 1.0; }
            const double rs = rsqrt(d);
            const double l = a[k] * rs;
            if (k + 1 < NC) piv = a[k + 1 < NC ? k + 1 : k] - l * l;
            // column k is published twice: in the packed factor (kept for the solves) and in a contiguous,
            // double-buffered broadcast array that the trailing update reads with 128-bit loads
            double* cb = S.colbuf + (k & 1) * (NC + 2);
            if (row && t > k) { A[MPC_PK(t, k)] = l; cb[t] = l; }
            if (t == k) S.dinv[k] = rs;
            // the owner of row k+1 publishes the next pivot together with its column entry: one
            // barrier per column covers both
            if (G::kThreads != 32 && t == k + 1 && k + 1 < NC) S.w[k + 1] = piv;
            g.sync();
            int j = k + 1;
            if ((j & 1) && j < NC) { if (FULL || j < n) a[j] -= l * cb[j]; ++j; }
#pragma unroll
            for (; j + 1 < NC; j += 2) {
                if (FULL || j < n) {   // nc is a multiple of 3 and j even: both elements lie inside or the pair is cut
                    const double2 c2 = *reinterpret_cast<const double2*>(cb + j);
                    a[j] -= l * c2.x;
                    if (FULL || j + 1 < n) a[j + 1] -= l * c2.y;
                }
            }
            if (j < NC && (FULL || j < n)) a[j] -= l * cb[j];
        }
    }
    g.sync();
    return ok;
}

template <class WK, class G>
__device__ __noinline__ bool cholesky_regs(WK& S, const G& g) {
    if (S.nc == WK::NC) return cholesky_regs_body<WK, G, true>(S, g);
    return cholesky_regs_body<WK, G, false>(S, g);
}

// ---- device fast path of the active-face solve: Gauss-Jordan elimination of [A | b], ONE ROW PER THREAD ----
// A (nc x nc, symmetric positive definite) arrives as the packed lower triangle, b as its row nc.  Thread t
// keeps row t of the symmetric matrix in registers as a SLIDING WINDOW: at column k, w[p] = A[t][k + p].
// Column k: every row publishes its column entry w[0] (by symmetry these are the pivot row's entries to the
// right of the diagonal) at the RELATIVE index t - k of a double-buffered broadcast array, the pivot comes
// from its owner by warp shuffle, and every row except the pivot row does  w[p-1] = w[p] - m c[p],
// b -= m b_k  with m = w[0] / d_k.  Rows above the pivot are updated too (the Jordan half), so the solution
// is simply x_t = b_t / d_t: no backward substitution and no stored factor.  Per-thread work equals the
// trailing update of a row-per-thread Cholesky (lanes above the pivot would otherwise idle).
// The shift makes the loop body independent of k, so the columns run in ROLLED loops (5 stages whose window
// shrinks by NC/5 each): ~12x less code than the unrolled register Cholesky + backward solve it replaces --
// the unrolled version spent 40 % of its time in instruction-cache misses (profiles/r1d).
// Rows/columns >= nc are padded with the identity, so one code path serves every compact size.
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));   // MUFU.RCP64H, relative error ~2^-23
    // one third-order step: y (1 + e + e^2), e = 1 - d y  ->  relative error ~e^3 = 2^-69 (three dependent DFMAs)
    const double e = fma(-d, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);
}

#ifndef MPC_GJ_STAGES
#define MPC_GJ_STAGES 5
#endif
template <int W, int BUF, class WK, class G>
__device__ __forceinline__ void gj_column(double (&w)[WK::NC], double& b, double& inv, double& myinv, bool& ok,
                                          WK& S, const G& g, const int k) {
    const int t = g.tid();
    // pivot-to-pivot chain: inv_k -> next pivot candidate (one DFMA) -> SHFL -> MUFU -> 3 DFMA; everything else
    // (the multiplier, the pivot-row select, the positivity test) is formed off that chain
    const double w0m = (t == k) ? 0.0 : w[0];          // the pivot row only shifts
    const double q = w0m * w[0];
    const double piv = fma(-q, inv, w[W > 1 ? 1 : 0]);  // next pivot candidate (c[1] of row k+1 is its own w[0])
    const double m = w0m * inv;
    if (t == k) myinv = inv;
    double* cb = S.colbuf + BUF * WK::CBS;             // static toggle: the caller alternates BUF
    if (t > k) cb[t - k] = w[0];
    if (t == k) cb[0] = b;
    if (G::kThreads != 32 && t == k + 1 && k + 1 < WK::NC) S.w[k + 1] = piv;
    g.sync();
    double d;
    if (G::kThreads == 32) d = __shfl_sync(0xffffffffu, piv, (k + 1) & 31);
    else d = S.w[k + 1 < WK::NC ? k + 1 : k];
    if (k + 1 < WK::NC && !(d > 0.0)) ok = false;       // reported as a failed solve; the garbage that follows is discarded
    inv = fast_rcp(d);
    b -= m * cb[0];
    if (W > 1) w[0] = w[W > 1 ? 1 : 0] - m * cb[1];
#pragma unroll
    for (int p = 2; p + 1 < W; p += 2) {
        const double2 c2 = *reinterpret_cast<const double2*>(cb + p);
        w[p - 1] = w[p] - m * c2.x;
        w[p] = w[p + 1] - m * c2.y;
    }
    if (W > 2 && (W & 1)) w[W - 2] = w[W - 1] - m * cb[W - 1];
}

template <int W, int KS, class WK, class G>
__device__ __forceinline__ void gj_stage(double (&w)[WK::NC], double& b, double& inv, double& myinv, bool& ok,
                                         WK& S, const G& g, const int k0) {
    // `inv` = 1 / d_k of the column about to be eliminated.  The loop is ROTATED: the reciprocal of the next
    // pivot is started as soon as its candidate exists (own registers, before the barrier), so that its
    // SHFL -> MUFU -> Newton chain overlaps the barrier, the broadcast loads and the W-1 row updates of the
    // current column (the warp issues in order: without the rotation the chain and the updates serialise).
    // Columns run in even/odd pairs so that the double-buffer toggle is static (k0 is even).
    static_assert(KS % 2 == 0, "even/odd column pairs");
#pragma unroll 1
    for (int k = k0; k < k0 + KS; k += 2) {
        gj_column<W, 0, WK, G>(w, b, inv, myinv, ok, S, g, k);
        gj_column<W, 1, WK, G>(w, b, inv, myinv, ok, S, g, k + 1);
    }
}

template <int K0, class WK, class G>
__device__ __forceinline__ void gj_stages(double (&w)[WK::NC], double& b, double& inv, double& myinv, bool& ok,
                                          WK& S, const G& g) {
    constexpr int NC = WK::NC, KS = NC / MPC_GJ_STAGES;
    static_assert(NC % MPC_GJ_STAGES == 0, "stage split");
    if constexpr (K0 < NC) {
        gj_stage<NC - K0, KS, WK, G>(w, b, inv, myinv, ok, S, g, K0);
        gj_stages<K0 + KS, WK, G>(w, b, inv, myinv, ok, S, g);
    }
}

// solves A x = b for the packed system of face_solve; x goes to S.w[0..nc).  false = not positive definite.
template <class WK, class G>
__device__ __noinline__ bool gj_solve_regs(WK& S, const G& g) {
    constexpr int NC = WK::NC;
    static_assert(G::kThreads >= NC, "one row per thread");
    const int n = S.nc, t = g.tid();
    const double* A = S.Ap();
    double w[NC];
    const bool row = t < n;
    double b;
    if (n == NC) {
        // full size (the common case): two predicated loads per entry, no padding logic.  Lanes beyond the matrix
        // hold a copy of the last row: they are never pivots and everything they publish lands in window slots
        // beyond column NC-1, which never feed a real slot.
        const int tt = t < NC ? t : NC - 1;
        const double* rowp = A + MPC_PK(tt, 0);         // A[t][j], j <= t
        const double* colp = A + tt;                    // A[j][t] = colp[j (j + 1) / 2], j > t
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            double v;
            if (j <= tt) v = rowp[j]; else v = colp[j * (j + 1) / 2];
            w[j] = v;
        }
        b = A[MPC_PK(NC, tt)];
    } else {
        const double* rowp = A + MPC_PK(row ? t : 0, 0);
        const double* colp = A + (row ? t : 0);
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double* q = (j <= t) ? rowp + j : colp + j * (j + 1) / 2;
            w[j] = (row && j < n) ? *q : ((j == t) ? 1.0 : 0.0);
        }
        b = row ? A[MPC_PK(n, t)] : 0.0;
    }
    double myinv = 0.0;
    bool ok = true;
    double d0;
    if (G::kThreads != 32) {   // multi-warp groups exchange the pivot through shared memory
        if (t == 0) S.w[0] = w[0];
        g.sync();
        d0 = S.w[0];
    } else {
        d0 = __shfl_sync(0xffffffffu, w[0], 0);
    }
    if (!(d0 > 0.0)) ok = false;
    double inv = fast_rcp(d0);
    gj_stages<0, WK, G>(w, b, inv, myinv, ok, S, g);
    g.sync();                  // every pivot read of S.w is done before it is overwritten with the solution
    if (row) S.w[t] = b * myinv;
    g.sync();
    return ok;
}

// ---- 3x3-block variant of the elimination for the multi-warp classes: one step per stance foot-step ---------------------
// The reduced Hessian is made of 3x3 blocks (one per stance foot-step), so the elimination can pivot on a whole block:
// every row publishes its first three window entries (by symmetry: the three pivot rows to the right of the block), every
// thread inverts the 3x3 pivot block (adjugate + one reciprocal), forms three multipliers and updates its window with
// three FMAs per entry.  Same FMA count as three scalar columns, but ONE barrier, one reciprocal and one set of selects
// per three columns, no pivot exchange through a separate array, and the pivot-to-pivot chain is walked NC/3 times.
// Used where a column step is expensive (two-warp groups: a named barrier and a shared-memory pivot per column) and
// registers are available (168-254 per thread); in the one-warp class, at its 128-register cap, it spills and loses 7 %
// against the scalar columns, so that class keeps gj_solve_regs.  The broadcast arrays alias S.ee..S.g.
template <int W, int BUF, class WK, class G>
__device__ __forceinline__ void gj3_step(double (&w)[WK::NC], double& b, double (&minv)[3], bool& ok, double* buf,
                                         const G& g, const int k) {
    constexpr int L = WK::NC;                 // length of one broadcast array (threads beyond the matrix do not publish)
    double* c0 = buf + BUF * (3 * L + 4);     // c0 | c1 | c2 | b of the three pivot rows
    double* c1 = c0 + L;
    double* c2 = c1 + L;
    double* bq = c2 + L;
    const int t = g.tid(), rel = t - k;
    if (rel >= 0 && t < WK::NC) { c0[rel] = w[0]; c1[rel] = w[1]; c2[rel] = w[2]; }
    if (rel >= 0 && rel < 3) bq[rel] = b;
    g.sync();
    // pivot block D[i][j] = c_j[i] (row k+i, column k+j), symmetric
    const double d00 = c0[0], d01 = c1[0], d02 = c2[0], d11 = c1[1], d12 = c2[1], d22 = c2[2];
    const double C00 = d11 * d22 - d12 * d12, C01 = d02 * d12 - d01 * d22, C02 = d01 * d12 - d02 * d11;
    const double C11 = d00 * d22 - d02 * d02, C12 = d01 * d02 - d00 * d12, C22 = d00 * d11 - d01 * d01;
    const double det = d00 * C00 + d01 * C01 + d02 * C02;
    if (!(d00 > 0.0) || !(C22 > 0.0) || !(det > 0.0)) ok = false;     // leading minors of a positive definite block
    const double id = fast_rcp(det);
    const double i00 = C00 * id, i01 = C01 * id, i02 = C02 * id, i11 = C11 * id, i12 = C12 * id, i22 = C22 * id;
    const bool piv = rel >= 0 && rel < 3;
    if (piv) {                                                        // keep this row of the block inverse for the end
        minv[0] = rel == 0 ? i00 : (rel == 1 ? i01 : i02);
        minv[1] = rel == 0 ? i01 : (rel == 1 ? i11 : i12);
        minv[2] = rel == 0 ? i02 : (rel == 1 ? i12 : i22);
    }
    const double a0 = piv ? 0.0 : w[0], a1 = piv ? 0.0 : w[1], a2 = piv ? 0.0 : w[2];   // pivot rows only shift
    const double n0 = -(a0 * i00 + a1 * i01 + a2 * i02), n1 = -(a0 * i01 + a1 * i11 + a2 * i12), n2 = -(a0 * i02 + a1 * i12 + a2 * i22);
    b = fma(n2, bq[2], fma(n1, bq[1], fma(n0, bq[0], b)));
    if (W > 3) w[0] = fma(n2, c2[3], fma(n1, c1[3], fma(n0, c0[3], w[W > 3 ? 3 : 0])));
#pragma unroll
    for (int p = 4; p + 1 < W; p += 2) {
        const double2 x0 = *reinterpret_cast<const double2*>(c0 + p);
        const double2 x1 = *reinterpret_cast<const double2*>(c1 + p);
        const double2 x2 = *reinterpret_cast<const double2*>(c2 + p);
        w[p - 3] = fma(n2, x2.x, fma(n1, x1.x, fma(n0, x0.x, w[p])));
        w[p - 2] = fma(n2, x2.y, fma(n1, x1.y, fma(n0, x0.y, w[p + 1])));
    }
}

template <int K0, class WK, class G>
__device__ __forceinline__ void gj3_stages(double (&w)[WK::NC], double& b, double (&minv)[3], bool& ok, double* buf, const G& g) {
    constexpr int NC = WK::NC;
    static_assert(NC % 6 == 0, "two 3-column steps per stage");
    if constexpr (K0 < NC) {
        gj3_step<NC - K0, 0, WK, G>(w, b, minv, ok, buf, g, K0);
        gj3_step<NC - K0, 1, WK, G>(w, b, minv, ok, buf, g, K0 + 3);
        gj3_stages<K0 + 6, WK, G>(w, b, minv, ok, buf, g);
    }
}

template <class WK>
struct Gj3Fits {   // two buffers of (3 NC + 4) doubles must fit into ee + adj + g = 30 (N + 1) + 6 N doubles
    static constexpr bool value = 2 * (3 * WK::NC + 4) <= 30 * (WK::N + 1) + 6 * WK::N && WK::NC % 6 == 0;
};

template <class WK, class G>
__device__ __noinline__ bool gj3_solve_regs(WK& S, const G& g) {
    constexpr int NC = WK::NC;
    static_assert(G::kThreads >= NC, "one row per thread");
    static_assert(offsetof(WK, adj) == offsetof(WK, ee) + sizeof(S.ee) && offsetof(WK, g) == offsetof(WK, adj) + sizeof(S.adj),
                  "ee, adj, g must be contiguous");
    const int n = S.nc, t = g.tid();
    const double* A = S.Ap();
    double w[NC];
    const bool row = t < n;
    double b;
    if (n == NC) {
        const int tt = t < NC ? t : NC - 1;
        const double* rowp = A + MPC_PK(tt, 0);
        const double* colp = A + tt;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            double v;
            if (j <= tt) v = rowp[j]; else v = colp[j * (j + 1) / 2];
            w[j] = v;
        }
        b = A[MPC_PK(NC, tt)];
    } else {
        const double* rowp = A + MPC_PK(row ? t : 0, 0);
        const double* colp = A + (row ? t : 0);
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const double* q = (j <= t) ? rowp + j : colp + j * (j + 1) / 2;
            w[j] = (row && j < n) ? *q : ((j == t) ? 1.0 : 0.0);
        }
        b = row ? A[MPC_PK(n, t)] : 0.0;
    }
    double minv[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    g.sync();                       // the aliased arrays are dead from here (every thread has left the rhs staging)
    gj3_stages<0, WK, G>(w, b, minv, ok, S.ee, g);
    // x of a pivot triple = (block inverse) * (final right-hand sides of the triple), exchanged through the (now idle)
    // broadcast area -- not through S.y/S.z, which carry the ADMM iterates when this runs as a polish step
    g.sync();
    if (t < NC) S.ee[t] = b;
    g.sync();
    if (row) {
        const int kb = (t / 3) * 3;
        S.w[t] = minv[0] * S.ee[kb] + minv[1] * S.ee[kb + 1] + minv[2] * S.ee[kb + 2];
    }
    g.sync();
    return ok;
}

// forward solve L y = b (b in S.w) for one-warp groups; y goes to row nc of the packed factor, which is
// where backward_regs expects it.  Column access A[PK(i,k)] over i is bank-conflict free.
// The L entries each lane needs do not depend on the recurrence, so they are loaded up front and the
// loop body is shuffle -> multiply -> fma only.
template <class WK, class G, bool FULL>
__device__ __forceinline__ void forward_regs_body(WK& S, const G& g) {
    constexpr int NC = WK::NC;
    const int n = FULL ? NC : S.nc, t = g.tid();
    double* A = S.Ap();
    double y = (t < n) ? S.w[t] : 0.0;
    const double di = (t < n) ? S.dinv[t] : 0.0;
    double lrow[NC];
    const double* mine = A + MPC_PK(t < n ? t : 0, 0);
#pragma unroll
    for (int k = 0; k < NC; ++k) lrow[k] = (t < n && k < t) ? mine[k] : 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        if (FULL || k < n) {
            if (t == k) y *= di;                       // y_k is final once all columns < k are applied
            const double yk = __shfl_sync(0xffffffffu, y, k);
            if (t > k) y -= lrow[k] * yk;
        }
    }
    if (t < n) A[MPC_PK(n, t)] = y;
    g.sync();
}
template <class WK, class G>
__device__ __noinline__ void forward_regs(WK& S, const G& g) {
    if (S.nc == WK::NC) forward_regs_body<WK, G, true>(S, g);
    else forward_regs_body<WK, G, false>(S, g);
}

// backward solve L' x = y with y_i held by thread i; x_k is broadcast by shuffle (one warp) or through
// shared memory (one barrier per step).  Lane t needs column t of L (L[k][t], k > t): loaded up front.
template <class WK, class G, bool FULL>
__device__ __forceinline__ void backward_regs_body(WK& S, const G& g) {
    constexpr int NC = WK::NC;
    const int n = FULL ? NC : S.nc, t = g.tid();
    const double* A = S.Ap();
    double y = (t < n) ? A[MPC_PK(n, t)] : 0.0;
    if (G::kThreads == 32) {
        const double di = (t < n) ? S.dinv[t] : 0.0;
        double lcol[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) lcol[k] = (k < n && t < k) ? A[MPC_PK(k, t)] : 0.0;
#pragma unroll
        for (int k = NC - 1; k >= 0; --k) {
            if (FULL || k < n) {
                if (t == k) y *= di;
                const double xk = __shfl_sync(0xffffffffu, y, k);
                if (t < k) y -= lcol[k] * xk;
            }
        }
        if (t < n) S.w[t] = y;
        g.sync();
    } else {
        g.sync();   // S.w is free (the factorisation no longer reads its pivots)
        for (int k = n - 1; k >= 0; --k) {
            if (t == k) { y *= S.dinv[k]; S.w[k] = y; }
            g.sync();
            if (t < k) y -= A[MPC_PK(k, t)] * S.w[k];
        }
        g.sync();
    }
}
template <class WK, class G>
__device__ __noinline__ void backward_regs(WK& S, const G& g) {
    if (S.nc == WK::NC) backward_regs_body<WK, G, true>(S, g);
    else backward_regs_body<WK, G, false>(S, g);
}
#endif

// ---- phase: left-looking packed Cholesky of the leading nc x nc block; row nc (the rhs) is carried
// along so that on exit A[PK(nc, k)] = (L^-1 rhs)_k.  Returns false if not positive definite.
template <class WK, class G>
MPC_HD bool cholesky_with_rhs(WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const int n = S.nc;
    double* A = S.Ap();
    if (g.tid() == 0) S.flag = 0;
    g.sync();
    for (int k = 0; k < n; ++k) {
        const double* rk = A + MPC_PK(k, 0);
        // every thread forms the pivot redundantly (row k is broadcast-read)
        for (int i = k + 1 + g.tid(); i <= n; i += g.size()) {
            double* ri = A + MPC_PK(i, 0);
            double s0 = ri[k], s1 = 0.0, d0 = rk[k], d1 = 0.0;
            int jj = 0;
            for (; jj + 1 < k; jj += 2) {
                double a0 = rk[jj], a1 = rk[jj + 1];
                s0 -= ri[jj] * a0; s1 -= ri[jj + 1] * a1;
                d0 -= a0 * a0; d1 -= a1 * a1;
            }
            if (jj < k) { double a0 = rk[jj]; s0 -= ri[jj] * a0; d0 -= a0 * a0; }
            double d = d0 + d1;
            if (!(d > 0.0)) { S.flag = 1; d = 1.0; }
            double di = 1.0 / sqrt(d);
            ri[k] = (s0 + s1) * di;
            if (i == k + 1) S.dinv[k] = di;   // row k+1 <= nc always exists (the rhs row)
        }
        g.sync();
    }
    return S.flag == 0;
}

// backward solve L' x = y with y = A[row nc], result in S.w[0..nc)
template <class WK, class G>
MPC_HD void backward_solve(WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const int n = S.nc;
    double* A = S.Ap();
    for (int i = g.tid(); i < n; i += g.size()) S.w[i] = A[MPC_PK(n, i)];
    g.sync();
    // w[k] is final once the steps above k are done and is never written again by the loop (only entries i < k are),
    // so the scaling by 1/L_kk can wait until the end: one barrier per step instead of two
    for (int k = n - 1; k >= 0; --k) {
        const double xk = S.w[k] * S.dinv[k];
        const double* rk = A + MPC_PK(k, 0);
        for (int i = g.tid(); i < k; i += g.size()) S.w[i] -= rk[i] * xk;
        g.sync();
    }
    for (int i = g.tid(); i < n; i += g.size()) S.w[i] *= S.dinv[i];
    g.sync();
}

// forward solve L y = b with b in S.w (used by ADMM where the factor is reused)
template <class WK, class G>
MPC_HD void forward_solve(WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    const int n = S.nc;
    double* A = S.Ap();
    for (int k = 0; k < n; ++k) {       // same deferral of the 1/L_kk scaling as in backward_solve
        const double yk = S.w[k] * S.dinv[k];
        for (int i = k + 1 + g.tid(); i < n; i += g.size()) S.w[i] -= A[MPC_PK(i, k)] * yk;
        g.sync();
    }
    for (int i = g.tid(); i < n; i += g.size()) S.w[i] *= S.dinv[i];
    g.sync();
}

// =====================================================================================================
// Tiled storage (Work types with TILED = true, horizon 50): padding, tiled right-looking Cholesky whose trailing
// updates run on the FP64 tensor cores (DMMA.8x8x4), blocked triangular solves.
// What it replaces: the dense factorisation the reference pays inside qpOASES for every solve
// (reference src/QPSolver.cpp:87-96 on the H of :58); here on the reduced Hessian of one active face.
// =====================================================================================================

// rows nc..8NT-1 become the identity (a compact size below the capacity, the right-hand-side row NC and the rows that
// only pad the last tile); the diagonal entry of the right-hand-side row is large so that its pivot stays positive
// whatever the right-hand side is (that pivot is never used).  Ends with a barrier: the caller fills row NC next.
template <class WK, class G>
MPC_HD void tiled_pad(WK& S, const G& g) {
    double* A = S.Ap();
    for (int r = S.nc; r < 8 * WK::NT; ++r)
        for (int c = g.tid(); c <= r; c += g.size()) A[WK::pk(r, c)] = (c == r) ? (r == WK::NC ? 1.0e30 : 1.0) : 0.0;
    g.sync();
}

// any group size (host build, single-warp groups): scalar right-looking blocked factorisation with the same result
// layout as the tensor-core path: strictly-lower tiles hold L, diagonal tiles hold inv(L_KK) (upper part zero).
template <class WK, class G>
MPC_HD bool chol_tiled_generic(WK& S, const G& g) {
    constexpr int NT = WK::NT;
    double* A = S.Ap();
    if (g.tid() == 0) S.flag = 0;
    g.sync();
    for (int K = 0; K < NT; ++K) {
        if (g.tid() == 0) {
            double a[8][8], x[8][8];
            for (int i = 0; i < 8; ++i) for (int j = 0; j <= i; ++j) a[i][j] = A[WK::pk(8 * K + i, 8 * K + j)];
            for (int k = 0; k < 8; ++k) {
                double d = a[k][k];
                if (!(d > 0.0)) { S.flag = 1; d = 1.0; }
                const double rs = 1.0 / sqrt(d);
                a[k][k] = rs;                                     // keep 1 / L_kk on the diagonal
                for (int i = k + 1; i < 8; ++i) a[i][k] *= rs;
                for (int j = k + 1; j < 8; ++j) for (int i = j; i < 8; ++i) a[i][j] -= a[i][k] * a[j][k];
            }
            for (int j = 0; j < 8; ++j) {                         // column j of X = L^-1 by forward substitution
                double r[8];
                for (int i = 0; i < 8; ++i) r[i] = (i == j) ? 1.0 : 0.0;
                for (int k = 0; k < 8; ++k) {
                    x[k][j] = r[k] * a[k][k];
                    for (int i = k + 1; i < 8; ++i) r[i] -= a[i][k] * x[k][j];
                }
            }
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) S.pbuf[i * 8 + j] = (j <= i) ? x[i][j] : 0.0;
            for (int i = 0; i < 8; ++i) for (int j = 0; j <= i; ++j) A[WK::pk(8 * K + i, 8 * K + j)] = x[i][j];
        }
        g.sync();
        // panel: L_IK = A_IK inv(L_KK)', one row of one tile per work item (rows are independent)
        for (int idx = g.tid(); idx < (NT - 1 - K) * 8; idx += g.size()) {
            const int i = 8 * (K + 1) + idx;
            double row[8], out[8];
            for (int m = 0; m < 8; ++m) row[m] = A[WK::pk(i, 8 * K + m)];
            for (int c = 0; c < 8; ++c) { double v = 0.0; for (int m = 0; m <= c; ++m) v += row[m] * S.pbuf[c * 8 + m]; out[c] = v; }
            for (int c = 0; c < 8; ++c) A[WK::pk(i, 8 * K + c)] = out[c];
        }
        g.sync();
        for (int i = 8 * (K + 1) + g.tid(); i < 8 * NT; i += g.size())
            for (int j = 8 * (K + 1); j <= i; ++j) {
                double v = 0.0;
                for (int m = 0; m < 8; ++m) v += A[WK::pk(i, 8 * K + m)] * A[WK::pk(j, 8 * K + m)];
                A[WK::pk(i, j)] -= v;
            }
        g.sync();
    }
    return S.flag == 0;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));     // MUFU.RSQ64H, relative error ~2^-23
    const double h = 0.5 * d;
    double e = fma(-h * y, y, 0.5);                              // two Newton steps: ~2^-45, then rounding-limited
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}
// 8x8 diagonal tile: Cholesky factor L and its inverse X = L^-1, one warp.  Lane j (mod 8) owns COLUMN j of the symmetric
// tile (all 8 rows, so that a_jk is its own register k) and column j of X.  Every lane gets column k by shuffles ONE STEP
// AHEAD (the old values, to which it applies the pending rank-1 update itself), so the pivot-to-pivot chain is
// rsqrt -> multiply -> fma with no shuffle on it; the forward substitution for X rides along (r -= l_k x_k).
// D: 64 doubles in tile layout, in = A_KK (lower triangle valid), out = X (upper part zero).
// Measured alone: 1.2 k cycles (tools/microbench/chol_dmma_bench.cu); a shuffle-free variant in which every lane
// eliminates the whole triangle redundantly took 5.4 k (bound by the FP64 issue rate of a single warp).
__device__ __noinline__ bool factor_diag8(double* __restrict__ D, int lane) {
    const int j = lane & 7;
    double a[8], r[8], x[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int lo = i > j ? i : j, hi = i > j ? j : i;            // symmetric read of the lower triangle
        a[i] = D[((hi & 4) << 3) + (lo << 2) + (hi & 3)];
        r[i] = (i == j) ? 1.0 : 0.0;
        x[i] = 0.0;
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __shfl_sync(0xffffffffu, a[i], 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double d = c[k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        const double rs = fast_rsqrt(d);
        double l[8], cn[8];
        if (k + 1 < 8) {             // old values of column k+1, fetched before this column's update is applied to them
#pragma unroll
            for (int i = k + 1; i < 8; ++i) cn[i] = __shfl_sync(0xffffffffu, a[i], k + 1);
        }
#pragma unroll
        for (int i = k + 1; i < 8; ++i) l[i] = c[i] * rs;
        const double ljk = a[k] * rs;                                 // l_jk (a_jk = a_kj: own register)
        x[k] = r[k] * rs;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) {
            a[i] -= l[i] * ljk;                                       // trailing update of my column (rows > k)
            r[i] -= l[i] * x[k];                                      // forward substitution for my column of X
            if (k + 1 < 8) c[i] = cn[i] - l[i] * l[k + 1];            // column k+1 as every lane needs it next
        }
    }
    __syncwarp();
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) D[((j & 4) << 3) + (i << 2) + (j & 3)] = (i >= j) ? x[i] : 0.0;
    }
    __syncwarp();
    return ok;
}

// Tiled right-looking Cholesky, NW = G::kThreads / 32 warps on one matrix (shared or global memory).  Step K:
//   panel    L_IK = A_IK inv(L_KK)'            two DMMAs per tile, written to its final place and to S.pbuf
//   trailing A_IJ -= L_IK L_JK' (K < J <= I)   two DMMAs per tile; A/B fragments are 32 consecutive doubles of S.pbuf
//                                              (conflict-free 64-bit loads), the accumulator is one 128-bit load/store
//   look-ahead: warp 0 is the critical-path warp.  It solves panel tile (K+1,K), signals the panel barrier WITHOUT
//   waiting on it (bar.arrive), updates tile (K+1,K+1) and factors it (factor_diag8) while the other warps finish the
//   panel and work through the update; one full barrier closes the step.  The factor's latency chain is hidden whenever
//   the update is long enough.  The inverse diagonal tile is double-buffered (warp 0 runs ahead of the panel readers).
// Measured (tools/microbench/chol_dmma_bench.cu, n = 150, one CTA of 8 warps per SM): 49 k cycles against 244 k for the
// packed left-looking shared-memory Cholesky on the FP64 CUDA cores (cholesky_with_rhs) it replaces.
template <class WK, class G>
__device__ __forceinline__ bool chol_tiled_dmma(WK& S, const G& g) {
    constexpr int NT = WK::NT, NW = G::kThreads / 32;
    static_assert(NW >= 2, "needs a look-ahead warp and at least one updating warp");
    double* __restrict__ T = S.Ap();
    double* __restrict__ pbuf = S.pbuf;
    const int tid = g.tid(), lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int fo = (gid << 2) + tig;                                     // fragment offset inside a half tile
    const int co = ((tig >> 1) << 5) + (gid << 2) + ((tig & 1) << 1);    // accumulator (C layout) offset inside a tile
    const int bar_panel = 8 + g.gid;                                     // named barrier of the panel phase (ids 1..7: g.sync)
    bool ok = true;
    if (warp == 0) {
        double* dbuf = S.pbuf + NT * 64;
        *reinterpret_cast<double2*>(dbuf + 2 * lane) = *reinterpret_cast<const double2*>(T + 2 * lane);
        __syncwarp();
        ok &= factor_diag8(dbuf, lane);
        *reinterpret_cast<double2*>(T + 2 * lane) = *reinterpret_cast<const double2*>(dbuf + 2 * lane);
    }
    g.sync();
    for (int K = 0; K < NT - 1; ++K) {
        const double* dbuf = S.pbuf + NT * 64 + ((K & 1) << 6);
        const double b_lo = dbuf[fo], b_hi = dbuf[32 + fo];
        if (warp == 0) {
            // critical path: panel tile (K+1,K) -> tile (K+1,K+1) -> its factor
            double* P = T + (((K + 1) * (K + 2) / 2 + K) << 6);
            const double a_lo = P[fo], a_hi = P[32 + fo];
            double c0 = 0.0, c1 = 0.0;
            dmma884(c0, c1, a_lo, b_lo);
            dmma884(c0, c1, a_hi, b_hi);
            __syncwarp();
            *reinterpret_cast<double2*>(P + co) = make_double2(c0, c1);
            *reinterpret_cast<double2*>(pbuf + ((K + 1) << 6) + co) = make_double2(c0, c1);
            __syncwarp();
            asm volatile("bar.arrive %0, %1;" ::"r"(bar_panel), "n"(G::kThreads) : "memory");
            double* Dn = T + (((K + 1) * (K + 2) / 2 + K + 1) << 6);
            double* dnext = S.pbuf + NT * 64 + (((K + 1) & 1) << 6);
            const double al = pbuf[((K + 1) << 6) + fo], ah = pbuf[((K + 1) << 6) + 32 + fo];
            double2 c = *reinterpret_cast<double2*>(Dn + co);
            dmma884(c.x, c.y, -al, al);
            dmma884(c.x, c.y, -ah, ah);
            __syncwarp();
            *reinterpret_cast<double2*>(dnext + co) = c;
            __syncwarp();
            ok &= factor_diag8(dnext, lane);
            *reinterpret_cast<double2*>(Dn + 2 * lane) = *reinterpret_cast<const double2*>(dnext + 2 * lane);
        } else {
            for (int I = K + 1 + warp; I < NT; I += NW - 1) {             // panel tiles K+2.. over the other warps
                double* P = T + ((I * (I + 1) / 2 + K) << 6);
                const double a_lo = P[fo], a_hi = P[32 + fo];
                double c0 = 0.0, c1 = 0.0;
                dmma884(c0, c1, a_lo, b_lo);
                dmma884(c0, c1, a_hi, b_hi);
                __syncwarp();
                *reinterpret_cast<double2*>(P + co) = make_double2(c0, c1);
                *reinterpret_cast<double2*>(pbuf + (I << 6) + co) = make_double2(c0, c1);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_panel), "n"(G::kThreads) : "memory");
            // tiles 1..M-1 in row-major order (tile 0 is the next diagonal tile), cut into contiguous ranges: a range mostly
            // stays inside one tile row, whose A fragments are reused; two tiles in flight per warp
            const int m = NT - 1 - K, M = m * (m + 1) / 2;
            const int t0 = 1 + ((M - 1) * (warp - 1)) / (NW - 1), t1 = 1 + ((M - 1) * warp) / (NW - 1);
            if (t0 < t1) {
                int r = (int)((sqrtf(8.0f * (float)t0 + 1.0f) - 1.0f) * 0.5f);
                while (r * (r + 1) / 2 > t0) --r;
                while ((r + 1) * (r + 2) / 2 <= t0) ++r;
                int I = K + 1 + r, J = K + 1 + (t0 - r * (r + 1) / 2);
                double a_lo = -pbuf[(I << 6) + fo], a_hi = -pbuf[(I << 6) + 32 + fo];
                for (int t = t0; t < t1; t += 2) {
                    double* C0 = T + ((I * (I + 1) / 2 + J) << 6) + co;
                    const double a0l = a_lo, a0h = a_hi;
                    const double b0l = pbuf[(J << 6) + fo], b0h = pbuf[(J << 6) + 32 + fo];
                    double2 c0 = *reinterpret_cast<double2*>(C0);
                    if (J == I) { ++I; J = K + 1; const int Ic = I < NT ? I : NT - 1; a_lo = -pbuf[(Ic << 6) + fo]; a_hi = -pbuf[(Ic << 6) + 32 + fo]; } else ++J;
                    const bool two = t + 1 < t1;
                    double* C1 = T + ((I * (I + 1) / 2 + J) << 6) + co;
                    const double a1l = a_lo, a1h = a_hi;
                    double b1l = 0.0, b1h = 0.0;
                    double2 c1 = make_double2(0.0, 0.0);
                    if (two) {
                        b1l = pbuf[(J << 6) + fo]; b1h = pbuf[(J << 6) + 32 + fo];
                        c1 = *reinterpret_cast<double2*>(C1);
                        if (J == I) { ++I; J = K + 1; const int Ic = I < NT ? I : NT - 1; a_lo = -pbuf[(Ic << 6) + fo]; a_hi = -pbuf[(Ic << 6) + 32 + fo]; } else ++J;
                    }
                    dmma884(c0.x, c0.y, a0l, b0l);
                    dmma884(c1.x, c1.y, a1l, b1l);
                    dmma884(c0.x, c0.y, a0h, b0h);
                    dmma884(c1.x, c1.y, a1h, b1h);
                    *reinterpret_cast<double2*>(C0) = c0;
                    if (two) *reinterpret_cast<double2*>(C1) = c1;
                }
            }
        }
        g.sync();
    }
    // only warp 0 has seen the pivots
    if (tid == 0) S.flag = ok ? 0 : 1;
    g.sync();
    return S.flag == 0;
}
#endif

template <class WK, class G>
MPC_HD bool chol_tiled(WK& S, const G& g) {
#if defined(__CUDA_ARCH__) && !defined(MPC_TILED_NO_DMMA)
    if constexpr (G::kThreads >= 64 && G::kThreads % 32 == 0) return chol_tiled_dmma<WK>(S, g);
    else
#endif
        return chol_tiled_generic<WK>(S, g);
}

// L y = b with b in S.w[0..nc) (zero beyond), blocked over the tile rows: y_K = inv(L_KK) s_K, then every later row
// subtracts L_rK y_K.  y goes to S.st (zero in the rows >= NC), where tiled_backward(.., false) picks it up.
// The loads of a step do not depend on the running sums, so they are issued before the dependent arithmetic
// (fixed trip counts: the upper parts of the inverse tiles are stored as zeros).
template <class WK, class G>
MPC_HD void tiled_forward(WK& S, const G& g) {
    constexpr int NT = WK::NT, NC = WK::NC;
    const double* A = S.Ap();
    for (int r = g.tid(); r < 8 * NT; r += g.size()) S.st[r] = (r < S.nc) ? S.w[r] : 0.0;
    g.sync();
    for (int K = 0; K < NT; ++K) {
        const double* XK = A + ((K * (K + 1) / 2 + K) << 6);
        for (int a = g.tid(); a < 8; a += g.size()) {
            double xr[8], sv[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) { xr[b] = XK[((b & 4) << 3) + (a << 2) + (b & 3)]; sv[b] = S.st[8 * K + b]; }
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int b = 0; b < 8; b += 2) { v0 = fma(xr[b], sv[b], v0); v1 = fma(xr[b + 1], sv[b + 1], v1); }
            S.xt[8 * K + a] = v0 + v1;
        }
        g.sync();
        for (int r = 8 * K + g.tid(); r < 8 * NT; r += g.size()) {
            if (r < 8 * K + 8) { S.st[r] = S.xt[r]; continue; }       // block K itself: s becomes y
            const double* Lr = A + (((r >> 3) * ((r >> 3) + 1) / 2 + K) << 6) + ((r & 7) << 2);
            double lv[8], yv[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) { lv[b] = Lr[((b & 4) << 3) + (b & 3)]; yv[b] = S.xt[8 * K + b]; }
            double a0 = S.st[r], a1 = 0.0;
#pragma unroll
            for (int b = 0; b < 8; b += 2) { a0 = fma(-lv[b], yv[b], a0); a1 = fma(-lv[b + 1], yv[b + 1], a1); }
            S.st[r] = a0 + a1;
        }
        g.sync();
    }
    for (int r = NC + g.tid(); r < 8 * NT; r += g.size()) S.st[r] = 0.0;
    g.sync();
}

// L' x = y, blocked: x_K = inv(L_KK)' s_K, then every earlier column c subtracts sum_a L[8K+a][c] x[8K+a].
// from_factor = true : y is row NC of the factor (the right-hand side that was carried through chol_tiled).  The last
//   tile's part of that row was overwritten by the tile's inverse X; its solution is read off X directly:
//   X[q][a] = -(1/delta) (L_qq^-T y_last)_a and X[q][q] = 1/delta (q = NC mod 8), so x_last = -X[q][.] / X[q][q].
// from_factor = false: y is in S.st (tiled_forward).
// The inverse diagonal tiles are first copied to S.pbuf (idle after the factorisation) so that the short dependent part
// of every step runs out of shared memory even when the factor lives in global memory; a column's eight factor entries
// of a step are loaded BEFORE the barrier that publishes x_K, so their latency overlaps it.
// Result in S.w[0..NC).
template <class WK, class G>
MPC_HD void tiled_backward(WK& S, const G& g, bool from_factor) {
    constexpr int NT = WK::NT, NC = WK::NC, KL = NT - 1;
    const double* A = S.Ap();
    for (int idx = g.tid(); idx < NT * 64; idx += g.size()) {
        const int K = idx >> 6;
        S.pbuf[idx] = A[((K * (K + 1) / 2 + K) << 6) + (idx & 63)];
    }
    if (from_factor) {
        const double xqq = A[WK::pk(NC, NC)];
        for (int c = g.tid(); c < 8 * NT; c += g.size()) {
            S.st[c] = (c < 8 * KL) ? A[WK::pk(NC, c)] : 0.0;
            S.xt[c] = (c >= 8 * KL && c < NC) ? -A[WK::pk(NC, c)] / xqq : 0.0;
        }
    }
    g.sync();
    for (int K = KL; K >= 0; --K) {
        // this step's factor entries of my column(s): independent of x, in flight across the barrier below
        const int c = g.tid();
        const bool mine = c < 8 * K;
        double lv[8];
        const int rot = c >> 2;     // every group of four columns starts at a different row: spreads the shared-memory banks
        if (mine) {
            const double* Lc = A + ((K * (K + 1) / 2 + (c >> 3)) << 6) + ((c & 4) << 3) + (c & 3);
#pragma unroll
            for (int a = 0; a < 8; ++a) lv[a] = Lc[((a + rot) & 7) << 2];
        }
        if (!(from_factor && K == KL)) {
            const double* XK = S.pbuf + (K << 6);
            for (int a = g.tid(); a < 8; a += g.size()) {
                double xc[8], sv[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) { xc[b] = XK[((a & 4) << 3) + (b << 2) + (a & 3)]; sv[b] = S.st[8 * K + b]; }
                double v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int b = 0; b < 8; b += 2) { v0 = fma(xc[b], sv[b], v0); v1 = fma(xc[b + 1], sv[b + 1], v1); }
                S.xt[8 * K + a] = (8 * K + a < NC) ? v0 + v1 : 0.0;
            }
            g.sync();
        }
        if (mine) {
            double xv[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) xv[a] = S.xt[8 * K + ((a + rot) & 7)];
            double a0 = S.st[c], a1 = 0.0;
#pragma unroll
            for (int a = 0; a < 8; a += 2) { a0 = fma(-lv[a], xv[a], a0); a1 = fma(-lv[a + 1], xv[a + 1], a1); }
            S.st[c] = a0 + a1;
        }
        // groups smaller than the matrix (host build, single-warp groups): the remaining columns, plain loop
        for (int c2 = g.tid() + g.size(); c2 < 8 * K; c2 += g.size()) {
            double acc = S.st[c2];
            for (int a = 0; a < 8; ++a) acc -= A[WK::pk(8 * K + a, c2)] * S.xt[8 * K + a];
            S.st[c2] = acc;
        }
        g.sync();
    }
    for (int i = g.tid(); i < NC; i += g.size()) S.w[i] = S.xt[i];
    g.sync();
}

// exact Euclidean projection onto {|x|<=mu z, |y|<=mu z, 0<=z<=fmax}; returns face codes
MPC_HD void project_pyramid(double mu, double fmax, const double v[3], double out[3], int& ax, int& ay, int& zt) {
    double x = v[0], y = v[1], w = v[2];
    double fxa = fabs(x), fya = fabs(y);
    double a = fxa > fya ? fxa : fya, b = fxa > fya ? fya : fxa, t;
    if (mu * w >= a) t = w;
    else {
        t = (w + mu * a) / (1.0 + mu * mu);
        if (mu * t < b) t = (w + mu * (a + b)) / (1.0 + 2.0 * mu * mu);
    }
    zt = 0;
    if (t >= fmax) { t = fmax; zt = 1; }
    else if (t <= 0.0) { t = 0.0; zt = 2; }
    double lim = mu * t;
    ax = x > lim ? 1 : (x < -lim ? -1 : 0);
    ay = y > lim ? 1 : (y < -lim ? -1 : 0);
    if (zt == 2) { ax = 0; ay = 0; }
    out[0] = x > lim ? lim : (x < -lim ? -lim : x);
    out[1] = y > lim ? lim : (y < -lim ? -lim : y);
    out[2] = t;
}

// ---- gradient g = H u + f  at u (full layout) -----------------------------------------------------
template <class WK, class G>
MPC_HD void gradient(const Tron1Const& P, WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    if constexpr (WK::RICCATI) {
        // the forward sweep left the full tracking error e0 + B u in S.ee: g = 2 B' Qbar (e0 + B u) + 2 r u in one adjoint pass
        adjoint<WK>(P, S, S.ee, S.gp(), g);
        MPC_TICK(S, g, 11);
        for (int i = g.tid(); i < 6 * N; i += g.size()) S.gp()[i] += 2.0 * P.r * S.up()[i];
        g.sync();
        return;
    }
    input_response<WK>(P, S, S.up(), g);
    MPC_TICK(S, g, 10);
    adjoint<WK>(P, S, S.ee, S.gp(), g);
    MPC_TICK(S, g, 11);
    for (int i = g.tid(); i < 6 * N; i += g.size()) S.gp()[i] += S.f[i] + 2.0 * P.r * S.up()[i];
    g.sync();
}

// ---- Riccati form of the active-face solve (work types with RICCATI = true) ----------------------------------------------
// The face-restricted problem  min_w  sum_i w_i (d_i + e0_i)' Q (d_i + e0_i) + r sum_k |u_k|^2,  u_k = c_k + Z_k w_k,
// d_{k+1} = A_k d_k + B_k u_k, d_0 = 0  (d = input response, e0 = free-response error; the same minimiser as the condensed
// system Z'HZ w = -Z'(f + H c) of face_solve, src/QPSolver.cpp:58-60) is a linear-quadratic tracking problem: one backward
// sweep over the horizon with a 12 x 12 value matrix, one forward sweep with the stored gains.  O(N) work, no matrix.
//   A_k = [[I, 0, Ts Rz_k', 0], [0, I, 0, Ts I], [0, 0, I, 0], [0, 0, 0, I]]   on (Theta, p, omega, v)
//   B_k = [Ts^2/2 Rz_k' W_ka ; Ts^2/(2m) I ; Ts W_ka ; Ts/m I]  per stance foot a       (input_response, same model)
// Eliminated variables of a face (Z column zero) keep a unit pivot and a zero right-hand side, exactly like build_hessian.
// Velocity-level form used by the sweeps.  With  beta = Ts (tau ; phi) = Bv u  (6 x 3 per foot: Bv_a = Ts [W_a ; I / m]) the
// step is  d' = A d + T beta,  T = [Ts/2 D ; I6],  D = blkdiag(Rz', I3),  and the last six columns of A are 2 T - [0 ; I6].
// Hence every product with A' or T' is the same three-term "left transform" over rows  lt(x)_i = Ts/2 (D' x[0:6])_i + x[6 + i],
// every product with A or T the same transform along a row.  With the face basis folded into the input matrix
// (Bz = Bv Z, 6 x mm; beta_c = Bv c) one backward step is
//   PT = P T,  X = P A = [P[:, 0:6] | 2 PT - P[:, 6:12]],  M = PT Bz,  t = PT beta_c + s                (row-local)
//   A'X = [X[0:6] ; 2 lt(X) - X[6:12]],  U6 = lt(M) = T'M,  V = A'M = [M[0:6] ; 2 U6 - M[6:12]],  A't likewise   (rows 6..11 read rows 0..5)
//   G = Bz' U6 + r Z'Z (mm x mm),  h = Bz' lt(t),   [K | kappa] = G^-1 [V' | h]
//   P_k = w_k Q + A'X - V K,   s_k = w_k Q e0_k + A't - V kappa,   and forward  w = -(K d + kappa).
// (Z'c = 0 for every face: a fixed component never shares a column with a free one.)
// Mapping: ONE WARP per instance with fixed lane roles -- lane r < 12 owns row r of the value matrix in registers (and its
// entry of s), lane a < mm builds row a of G, lanes 0..12 each solve one column of [V' | h] against G (the small LDL'
// factorisation is repeated per column in registers), and the four exchanges of a step go through a few hundred bytes of shared
// memory with a warp barrier.  Every loop has a compile-time trip count per stance-foot count (mm = 0, 3, 6); nothing is
// indexed dynamically.  The host build runs the same phase functions over an array of 32 lane states.
struct RicLane {
    double Pr[12];   // row r of P_{k+1}
    double s;        // s_{k+1}[r]
    double X[12];    // row r of P A, then of A'PA
    double M[6];     // row r of M, then of V = A'M
    double t;        // t[r], then (A't)[r]
};
struct RicStep {     // uniform per step
    int k, f0, sr;   // step, foot-step of the first stance foot (compact columns 0..2), right foot-step (columns 3..5 when both stand)
    double cz, sz, wk;
    double two;      // 2.0 the compiler cannot see through: 2 x - y stays one fused operation instead of (x + x) - y
    FaceZ Z[2];      // face basis per compact slot
    double cfx[2], cfy[2], cfz[2];   // fixed part c of u = c + Z w per slot (non-zero only with the normal force at its upper bound)
    double* Ku;      // this step's forward gains in force space: 6 rows [Ku (12) | ku0] (left foot xyz, right foot xyz), u = ku0 + Ku d
};

template <int MM, class WK>
MPC_HD RicStep ric_make_step(const Tron1Const& P, WK& S, int k) {
    RicStep st = {};
    st.k = k; st.sr = 2 * k + 1;
    st.f0 = S.contact[2 * k] ? 2 * k : 2 * k + 1;
    st.cz = S.cs[2 * k]; st.sz = S.cs[2 * k + 1];
    st.wk = step_weight<WK::N>(P, k);
    st.two = 2.0 + 0.0 * P.Ts;
#pragma unroll
    for (int a = 0; a < MM / 3; ++a) {
        const int s_ = a == 0 ? st.f0 : st.sr;
        st.Z[a] = face_basis(P.mu, S.ax[s_], S.ay[s_], S.zt[s_]);
        const double fz = S.zt[s_] == 1 ? P.f_max : 0.0;
        st.cfx[a] = (double)S.ax[s_] * P.mu * fz; st.cfy[a] = (double)S.ay[s_] * P.mu * fz; st.cfz[a] = fz;
    }
    st.Ku = S.Kp() + 78 * k;
    return st;
}

// run a phase: the device warp calls it once per thread, the host build once per emulated lane
template <class G, class F>
MPC_HD void ric_lanes(const G& g, RicLane* L, F f) {
    if constexpr (G::kThreads == 32) f(g.tid(), L[0]);
    else for (int l = 0; l < 32; ++l) f(l, L[l]);
}

// inverse of a symmetric positive definite 3 x 3 matrix by cofactors, lower triangle in the order (0,0) (1,0) (1,1) (2,0) (2,1) (2,2);
// false when a leading minor is not positive
MPC_HD bool ric_inv3(const double* a, double* inv) {
    const double c00 = a[2] * a[5] - a[4] * a[4], c10 = a[4] * a[3] - a[1] * a[5], c20 = a[1] * a[4] - a[2] * a[3];
    const double c11 = a[0] * a[5] - a[3] * a[3], c21 = a[1] * a[3] - a[0] * a[4], c22 = a[0] * a[2] - a[1] * a[1];
    const double det = a[0] * c00 + a[1] * c10 + a[3] * c20;
    const double r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = c10 * r; inv[2] = c11 * r; inv[3] = c20 * r; inv[4] = c21 * r; inv[5] = c22 * r;
    return a[0] > 0.0 && c22 > 0.0 && det > 0.0;
}

template <int MM, class WK, class G>
MPC_HD bool riccati_backward_step(const Tron1Const& P, WK& S, const G& g, RicLane* LL, int k) {
    static_assert(G::kThreads == 32 || G::kThreads == 1, "the Riccati class maps one warp to an instance");
    const RicStep st = ric_make_step<MM, WK>(P, S, k);
    const double Ts = P.Ts, hTs = 0.5 * P.Ts, im = P.inv_m;
    double* XM = S.rc + WK::RC_XM;     // rows 0..5 of [X (12) | M (6) | t], stride 20
    double* U6 = S.rc + WK::RC_U6;     // [U6 row i (6) | lt(t)_i], stride 8
    double* GH = S.rc + WK::RC_GH;     // [G row a (6) | h_a], stride 8
    double* KT = S.rc + WK::RC_KT;     // column j of [K | kappa] (6 entries), stride 6
    const double* wq = S.rc + WK::RC_WQ;   // Q diagonal
    bool ok = true;
    // phase A (lanes 0..11, registers only): PT, X = P A, M = PT Bz, t; rows 0..5 are published
    ric_lanes(g, LL, [&](int lane, RicLane& L) {
        if (lane >= 12) return;
        double PT[6];
        PT[0] = hTs * (st.cz * L.Pr[0] - st.sz * L.Pr[1]) + L.Pr[6];
        PT[1] = hTs * (st.sz * L.Pr[0] + st.cz * L.Pr[1]) + L.Pr[7];
        PT[2] = hTs * L.Pr[2] + L.Pr[8];
#pragma unroll
        for (int j = 3; j < 6; ++j) PT[j] = hTs * L.Pr[j] + L.Pr[6 + j];
#pragma unroll
        for (int j = 0; j < 6; ++j) { L.X[j] = L.Pr[j]; L.X[6 + j] = fma(st.two, PT[j], -L.Pr[6 + j]); }
        double t = L.s;
#pragma unroll
        for (int a = 0; a < MM / 3; ++a) {
            const double* Wa = S.W + 9 * (a == 0 ? st.f0 : st.sr);
            double Mv[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) Mv[c] = Ts * (PT[0] * Wa[c] + PT[1] * Wa[3 + c] + PT[2] * Wa[6 + c] + im * PT[3 + c]);
            const FaceZ& Z = st.Z[a];
            L.M[3 * a] = Z.fx * Mv[0];
            L.M[3 * a + 1] = Z.fy * Mv[1];
            L.M[3 * a + 2] = Z.fz * (Z.mx * Mv[0] + Z.my * Mv[1] + Mv[2]);
            if (st.cfz[a] != 0.0) t += st.cfx[a] * Mv[0] + st.cfy[a] * Mv[1] + st.cfz[a] * Mv[2];   // PT beta_c
        }
        L.t = t;
        if (lane < 6) {
            double* row = XM + 20 * lane;
#pragma unroll
            for (int j = 0; j < 12; ++j) row[j] = L.X[j];
#pragma unroll
            for (int b = 0; b < MM; ++b) row[12 + b] = L.M[b];
            row[18] = t;
        }
    });
    g.sync();
    MPC_RTICK(1, S, g, 4);
    // phase B (lanes 6..11): the left transform of rows 0..5 turns X into A'X, M into V and t into A't; U6 and lt(t) are published
    ric_lanes(g, LL, [&](int lane, RicLane& L) {
        if (lane < 6 || lane >= 12) return;
        const int i = lane - 6;
        const double* ra = XM + 20 * (i < 2 ? 0 : i);
        const double* rb = XM + 20 * (i < 2 ? 1 : i);
        const double ca = hTs * (i == 0 ? st.cz : (i == 1 ? st.sz : 1.0)), cb = hTs * (i == 0 ? -st.sz : (i == 1 ? st.cz : 0.0));
        double* urow = U6 + 8 * i;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const double lt = ca * ra[j] + cb * rb[j] + L.X[j];
            L.X[j] = fma(st.two, lt, -L.X[j]);
        }
#pragma unroll
        for (int b = 0; b < MM; ++b) {
            const double lt = ca * ra[12 + b] + cb * rb[12 + b] + L.M[b];
            urow[b] = lt;
            L.M[b] = fma(st.two, lt, -L.M[b]);
        }
        const double lt = ca * ra[18] + cb * rb[18] + L.t;
        urow[6] = lt;
        L.t = fma(st.two, lt, -L.t);
    });
    if constexpr (MM > 0) {
        g.sync();
        // phase C (lanes a < mm): column a of Bz, row a of G = Bz' U6 + r Z'Z (unit pivot for an eliminated variable), h_a
        ric_lanes(g, LL, [&](int lane, RicLane&) {
            if (lane >= MM) return;
            const int a = lane / 3, c = lane - 3 * a;
            FaceZ Z;          // selects, not an indexed access: the step record stays in registers
            Z.fx = a == 0 ? st.Z[0].fx : st.Z[1].fx; Z.fy = a == 0 ? st.Z[0].fy : st.Z[1].fy; Z.fz = a == 0 ? st.Z[0].fz : st.Z[1].fz;
            Z.mx = a == 0 ? st.Z[0].mx : st.Z[1].mx; Z.my = a == 0 ? st.Z[0].my : st.Z[1].my;
            const double k0 = c == 0 ? Z.fx : (c == 2 ? Z.fz * Z.mx : 0.0), k1 = c == 1 ? Z.fy : (c == 2 ? Z.fz * Z.my : 0.0),
                         k2 = c == 2 ? Z.fz : 0.0;
            const double* Wa = S.W + 9 * (a == 0 ? st.f0 : st.sr);
            double Bz[6];
#pragma unroll
            for (int l = 0; l < 3; ++l) Bz[l] = Ts * (k0 * Wa[3 * l] + k1 * Wa[3 * l + 1] + k2 * Wa[3 * l + 2]);
            Bz[3] = Ts * im * k0; Bz[4] = Ts * im * k1; Bz[5] = Ts * im * k2;
            const double zz = k0 * k0 + k1 * k1 + k2 * k2;
            double* grow = GH + 8 * lane;
#pragma unroll
            for (int b = 0; b < MM; ++b) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) v = fma(Bz[i], U6[8 * i + b], v);
                if (b == lane) v += (zz == 0.0) ? 1.0 : P.r * zz;
                grow[b] = v;
            }
            double h = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) h = fma(Bz[i], U6[8 * i + 6], h);
            grow[6] = h;
        });
        g.sync();
        MPC_RTICK(1, S, g, 5);
        // phase D (lanes 0..12): one column of [K | kappa] = G^-1 [V' | h] per lane
        ric_lanes(g, LL, [&](int lane, RicLane& L) {
            if (lane > 12) return;
            // 3 x 3 blocks with cofactor inverses (one reciprocal per block, shallow dependency chains) instead of a scalar LDL'
            // with six reciprocals in sequence:  G = [[A, B'], [B, C]],  T = B A^-1,  S = C - T B',
            // x2 = S^-1 (r2 - T r1),  x1 = A^-1 r1 - T' x2
            double rhs[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) rhs[i] = (lane < 12) ? L.M[i] : GH[8 * i + 6];
            double A[6], iA[6];
            A[0] = GH[0]; A[1] = GH[8]; A[2] = GH[9]; A[3] = GH[16]; A[4] = GH[17]; A[5] = GH[18];
            if (!ric_inv3(A, iA)) ok = false;
            if constexpr (MM == 3) {
                const double r0 = rhs[0], r1 = rhs[1], r2 = rhs[2];
                rhs[0] = iA[0] * r0 + iA[1] * r1 + iA[3] * r2;
                rhs[1] = iA[1] * r0 + iA[2] * r1 + iA[4] * r2;
                rhs[2] = iA[3] * r0 + iA[4] * r1 + iA[5] * r2;
            } else {
                double B[3][3], T[3][3], C[6], iS[6];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) B[i][j] = GH[8 * (3 + i) + j];
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    T[i][0] = B[i][0] * iA[0] + B[i][1] * iA[1] + B[i][2] * iA[3];
                    T[i][1] = B[i][0] * iA[1] + B[i][1] * iA[2] + B[i][2] * iA[4];
                    T[i][2] = B[i][0] * iA[3] + B[i][1] * iA[4] + B[i][2] * iA[5];
                }
                int q = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j <= i; ++j, ++q)
                        C[q] = GH[8 * (3 + i) + 3 + j] - (T[i][0] * B[j][0] + T[i][1] * B[j][1] + T[i][2] * B[j][2]);
                }
                if (!ric_inv3(C, iS)) ok = false;
                const double r0 = rhs[0], r1 = rhs[1], r2 = rhs[2];
                const double s0 = rhs[3] - (T[0][0] * r0 + T[0][1] * r1 + T[0][2] * r2);
                const double s1 = rhs[4] - (T[1][0] * r0 + T[1][1] * r1 + T[1][2] * r2);
                const double s2 = rhs[5] - (T[2][0] * r0 + T[2][1] * r1 + T[2][2] * r2);
                const double x3 = iS[0] * s0 + iS[1] * s1 + iS[3] * s2;
                const double x4 = iS[1] * s0 + iS[2] * s1 + iS[4] * s2;
                const double x5 = iS[3] * s0 + iS[4] * s1 + iS[5] * s2;
                rhs[0] = iA[0] * r0 + iA[1] * r1 + iA[3] * r2 - (T[0][0] * x3 + T[1][0] * x4 + T[2][0] * x5);
                rhs[1] = iA[1] * r0 + iA[2] * r1 + iA[4] * r2 - (T[0][1] * x3 + T[1][1] * x4 + T[2][1] * x5);
                rhs[2] = iA[3] * r0 + iA[4] * r1 + iA[5] * r2 - (T[0][2] * x3 + T[1][2] * x4 + T[2][2] * x5);
                rhs[3] = x3; rhs[4] = x4; rhs[5] = x5;
            }
#pragma unroll
            for (int i = 0; i < MM; ++i) KT[6 * lane + i] = rhs[i];      // this step's update reads the compact gains
            // the forward sweep reads the gains in force space, u = c + Z w = (c - Z kappa) - (Z K) d: no face logic there
#pragma unroll
            for (int a = 0; a < MM / 3; ++a) {
                const FaceZ& Z = st.Z[a];
                const int ft = (a == 0 ? st.f0 : st.sr) & 1;
                const double kz = Z.fz * rhs[3 * a + 2];
                double ux = -(Z.fx * rhs[3 * a] + Z.mx * kz), uy = -(Z.fy * rhs[3 * a + 1] + Z.my * kz), uz = -kz;
                if (lane == 12) { ux += st.cfx[a]; uy += st.cfy[a]; uz += st.cfz[a]; }
                double* row = st.Ku + 39 * ft + lane;
                row[0] = ux; row[13] = uy; row[26] = uz;
            }
        });
    }
    if (k > 0) {     // step 0 needs neither P_0 nor s_0
        g.sync();
        MPC_RTICK(1, S, g, 6);
        // phase E (lanes 0..11): P_k = w_k Q + A'X - V K,  s_k = w_k Q e0_k + A't - V kappa
        ric_lanes(g, LL, [&](int lane, RicLane& L) {
            if (lane >= 12) return;
            const double wqr = st.wk * wq[lane];
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                double v = L.X[j];
                if constexpr (MM > 0) {
#pragma unroll
                    for (int a = 0; a < MM; ++a) v = fma(-L.M[a], KT[6 * j + a], v);
                }
#if defined(__CUDA_ARCH__)
                // one compare and one predicated add (the compiler's select form costs an add and two selects per entry)
                asm("{\n.reg .pred p;\nsetp.eq.s32 p, %1, %2;\n@p add.f64 %0, %0, %3;\n}" : "+d"(v) : "r"(lane), "r"(j), "d"(wqr));
#else
                if (j == lane) v += wqr;
#endif
                L.Pr[j] = v;
            }
            double v = L.t;
            if constexpr (MM > 0) {
#pragma unroll
                for (int a = 0; a < MM; ++a) v = fma(-L.M[a], KT[6 * 12 + a], v);
            }
            L.s = v + wqr * S.ee[12 * k + lane];
        });
        // no barrier here: the rows stay in their lanes' registers, and every exchange buffer is next written behind another barrier
        MPC_RTICK(1, S, g, 9);
    }
    return ok;
}

// forward sweep, one step: u = ku0 + Ku d (gains in force space), d' = A d + T Bv u; the state is double-buffered in shared
// memory.  External gains arrive through a four-slot ring in shared memory filled by asynchronous copies three steps ahead.
template <class WK, class G>
MPC_HD void ric_fetch_gains(WK& S, const G& g, int k) {
    if constexpr (!WK::AINL) {
        if (k < WK::N) {
            // the whole 78-double block of the step, 16 bytes per copy (rows of a swing foot hold stale data and are never read)
            const double* Kn = S.Kp() + 78 * k;
            double* dst = S.rc + WK::RC_KF + 78 * (k & 3);
            for (int it = g.tid(); it < 39; it += G::kThreads) {
#if defined(__CUDA_ARCH__)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + 2 * it)), "l"(Kn + 2 * it) : "memory");
#else
                dst[2 * it] = Kn[2 * it]; dst[2 * it + 1] = Kn[2 * it + 1];
#endif
            }
        }
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    }
}

template <class WK, class G>
MPC_HD void riccati_forward_step(const Tron1Const& P, WK& S, const G& g, int k) {
    const double Ts = P.Ts;
    const double* d = S.rc + WK::RC_D + 12 * (k & 1);
    double* dn = S.rc + WK::RC_D + 12 * ((k + 1) & 1);
    const bool inl = S.contact[2 * k] != 0, inr = S.contact[2 * k + 1] != 0;
    if constexpr (!WK::AINL) {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.wait_group 2;" ::: "memory");    // the copies of step k have landed (steps k+1, k+2 may be in flight)
#endif
        g.sync();
        ric_fetch_gains<WK>(S, g, k + 3);
    }
    MPC_RTICK(2, S, g, 4);
    // the forces of this step: read back by the state update below from shared memory (external storage: a copy goes there)
    double* uk = WK::UEXT ? S.rc + WK::RC_UK : S.up() + 6 * k;
    {
        const double* Kk = WK::AINL ? S.Kp() + 78 * k : S.rc + WK::RC_KF + 78 * (k & 3);
        for (int i = g.tid(); i < 6; i += G::kThreads) {
            double v = 0.0;
            if (i < 3 ? inl : inr) {
                const double* row = Kk + 13 * i;
                double v0 = row[12], v1 = 0.0, v2 = 0.0;
#pragma unroll
                for (int l = 0; l < 12; l += 3) { v0 = fma(row[l], d[l], v0); v1 = fma(row[l + 1], d[l + 1], v1); v2 = fma(row[l + 2], d[l + 2], v2); }
                v = v0 + (v1 + v2);
            }
            uk[i] = v;
            if constexpr (WK::UEXT) S.up()[6 * k + i] = v;
        }
        g.sync();
    }
    MPC_RTICK(2, S, g, 5);
    // d'[r] = d[r] + gam (rho . d_omega) + mu d[9 + c] + al (rho . tau) + lam (uL[c] + uR[c]),  tau = sum_a W_a u_a: the same code in
    // every lane, the row's coefficients selected without branches
    //   Theta rows: rho = row r of Rz', gam = Ts, al = Ts^2/2;  p rows: mu = Ts, lam = Ts^2/(2m);  omega rows: rho = e_c, al = Ts;  v rows: lam = Ts/m
    const double cz = S.cs[2 * k], sz = S.cs[2 * k + 1];
    for (int r = g.tid(); r < 12; r += G::kThreads) {
        const int c = r % 3, blk = r / 3;      // blk: 0 Theta, 1 p, 2 omega, 3 v
        const double* W0 = S.W + 18 * k;
        double tau[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            tau[i] = (W0[3 * i] * uk[0] + W0[3 * i + 1] * uk[1] + W0[3 * i + 2] * uk[2]) + (W0[9 + 3 * i] * uk[3] + W0[10 + 3 * i] * uk[4] + W0[11 + 3 * i] * uk[5]);
        const bool rot = blk == 0;
        const double rho0 = rot ? (c == 0 ? cz : (c == 1 ? -sz : 0.0)) : (c == 0 ? 1.0 : 0.0);
        const double rho1 = rot ? (c == 0 ? sz : (c == 1 ? cz : 0.0)) : (c == 1 ? 1.0 : 0.0);
        const double rho2 = c == 2 ? 1.0 : 0.0;
        const double gam = blk == 0 ? Ts : 0.0, mu = blk == 1 ? Ts : 0.0;
        const double al = blk == 0 ? 0.5 * Ts * Ts : (blk == 2 ? Ts : 0.0);
        const double lam = blk == 1 ? 0.5 * Ts * Ts * P.inv_m : (blk == 3 ? Ts * P.inv_m : 0.0);
        const double rw = rho0 * d[6] + rho1 * d[7] + rho2 * d[8];
        const double rt = rho0 * tau[0] + rho1 * tau[1] + rho2 * tau[2];
        const double v = d[r] + gam * rw + mu * d[9 + c] + al * rt + lam * (uk[c] + uk[3 + c]);
        dn[r] = v;
        S.ee[12 * (k + 1) + r] += v;     // e0 + B u: the full tracking error for the gradient pass (e0 is rebuilt for a further face)
    }
    g.sync();
    MPC_RTICK(2, S, g, 6);
}

template <class WK, class G>
MPC_HD bool riccati_face_solve(const Tron1Const& P, WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    RicLane LL[G::kThreads == 32 ? 1 : 32];
    bool ok = true;
    {   // terminal value: P_N = w_N Q, s_N = w_N Q e0_N;  d_0 = 0;  the Q diagonal where a lane can index it
        const double wN = step_weight<N>(P, N);
        for (int i = g.tid(); i < 12; i += G::kThreads) { (S.rc + WK::RC_WQ)[i] = P.q[i]; (S.rc + WK::RC_D)[i] = 0.0; }
        g.sync();
        ric_lanes(g, LL, [&](int lane, RicLane& L) {
            const int r = lane < 12 ? lane : 0;
            const double wq = wN * (S.rc + WK::RC_WQ)[r];
#pragma unroll
            for (int j = 0; j < 12; ++j) L.Pr[j] = (j == r) ? wq : 0.0;
            L.s = wq * S.ee[12 * N + r];
#pragma unroll
            for (int j = 0; j < 6; ++j) L.M[j] = 0.0;
        });
    }
    for (int k = N - 1; k >= 0; --k) {
        const int nf = (S.contact[2 * k] ? 1 : 0) + (S.contact[2 * k + 1] ? 1 : 0);
        bool okk;
        if (nf == 2) okk = riccati_backward_step<6, WK>(P, S, g, LL, k);
        else if (nf == 1) okk = riccati_backward_step<3, WK>(P, S, g, LL, k);
        else okk = riccati_backward_step<0, WK>(P, S, g, LL, k);
        ok = ok && okk;
    }
    MPC_TICK(S, g, 7);
    g.sync();       // the gains of step 0 were written by other lanes
    ric_fetch_gains<WK>(S, g, 0);
    ric_fetch_gains<WK>(S, g, 1);
    ric_fetch_gains<WK>(S, g, 2);
    for (int k = 0; k < N; ++k) riccati_forward_step<WK>(P, S, g, k);
#if defined(__CUDA_ARCH__)
    if constexpr (!WK::AINL) asm volatile("cp.async.wait_all;" ::: "memory");
#endif
    MPC_TICK(S, g, 8);
    // the pivots were seen by the lanes that solved: combine
    if (g.tid() == 0) S.flag = 0;
    g.sync();
    if (!ok) S.flag = 1;
    g.sync();
    const bool all_ok = S.flag == 0;
    g.sync();
    return all_ok;
}

// ---- one active-face solve: u = argmin q on the affine hull of the current face -------------------
// returns false when the reduced Hessian is not positive definite (never for valid inputs)
template <class WK, class G>
MPC_HD bool face_solve(const Tron1Const& P, WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    if constexpr (WK::RICCATI) {
        if (g.tid() == 0) S.interior = 0;
        const bool okr = riccati_face_solve<WK>(P, S, g);
        MPC_TICK(S, g, 7);
        return okr;
    } else {
    // fixed part: z = fmax on zt==1 foot-steps
    bool any_fixed = false;
    [[maybe_unused]] bool any_reduced = false;   // some stance foot-step sits on a face other than the interior one
#if defined(__CUDA_ARCH__)
    if constexpr (G::kThreads >= 2 * N) {     // one foot-step per thread, group-wide votes
        const int s = g.tid();
        const bool st = s < 2 * N && S.contact[s];
        any_fixed = g.any(st && S.zt[s] == 1);
#if defined(MPC_REDUCED_CERTIFICATE)
        any_reduced = g.any(st && (S.ax[s] | S.ay[s] | S.zt[s]) != 0);
        if (g.tid() == 0) S.interior = any_reduced ? 0 : 1;
#endif
    } else
#endif
    {
        for (int s = 0; s < 2 * N; ++s) any_fixed |= (S.contact[s] && S.zt[s] == 1);
        if (g.tid() == 0) S.interior = 0;
    }
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        double fz = (S.contact[s] && S.zt[s] == 1) ? P.f_max : 0.0;
        S.up()[3 * s] = (double)S.ax[s] * P.mu * fz;
        S.up()[3 * s + 1] = (double)S.ay[s] * P.mu * fz;
        S.up()[3 * s + 2] = fz;
    }
    g.sync();
    if (any_fixed) gradient<WK>(P, S, g);   // g0 = f + H u_fix
    MPC_TICK(S, g, 4);
    build_hessian<WK>(P, S, 0.0, true, g);
    MPC_TICK(S, g, 5);
    const int n = S.nc;
    if constexpr (WK::TILED) tiled_pad<WK>(S, g);
    const int R = S.rhs_row();
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        if (!S.contact[s]) continue;
        FaceZ Z = face_basis(P.mu, S.ax[s], S.ay[s], S.zt[s]);
        const double* g0 = any_fixed ? S.gp() + 3 * s : S.f + 3 * s;
        double* A = S.Ap();
        const int c0 = 3 * S.cidx[s];
        A[WK::pk(R, c0)] = -Z.fx * g0[0];
        A[WK::pk(R, c0 + 1)] = -Z.fy * g0[1];
        A[WK::pk(R, c0 + 2)] = -Z.fz * (Z.mx * g0[0] + Z.my * g0[1] + g0[2]);
    }
    if constexpr (!WK::TILED) { if (g.tid() == 0) S.Ap()[MPC_PK(n, n)] = 1.0; }
    g.sync();
    MPC_TICK(S, g, 6);
    bool ok;
    if constexpr (WK::TILED) {
        ok = chol_tiled<WK>(S, g);
        MPC_TICK(S, g, 7);
        tiled_backward<WK>(S, g, true);
    } else
#if defined(__CUDA_ARCH__)
    if constexpr (G::kThreads >= WK::NC && WK::NC <= 60) {
        // measured: +12 % for the double-support class of horizon 10 (168 registers), -9 % for horizon 20 (already at 254)
        if constexpr (G::kThreads > 32 && WK::NC == 6 * WK::N && Gj3Fits<WK>::value) ok = gj3_solve_regs<WK>(S, g);
        else ok = gj_solve_regs<WK>(S, g);
        MPC_TICK(S, g, 7);
    } else
#endif
    {
        ok = cholesky_with_rhs<WK>(S, g);
        MPC_TICK(S, g, 7);
        backward_solve<WK>(S, g);
    }
    MPC_TICK(S, g, 8);
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        if (!S.contact[s]) continue;
        const double* w = S.w + 3 * S.cidx[s];
        int zt = S.zt[s];
        double fz = zt == 0 ? w[2] : (zt == 1 ? P.f_max : 0.0);
        S.up()[3 * s + 2] = fz;
        S.up()[3 * s] = S.ax[s] != 0 ? (double)S.ax[s] * P.mu * fz : (zt == 2 ? 0.0 : w[0]);
        S.up()[3 * s + 1] = S.ay[s] != 0 ? (double)S.ay[s] * P.mu * fz : (zt == 2 ? 0.0 : w[1]);
    }
    g.sync();
    MPC_TICK(S, g, 9);
    return ok;
    }
}

// ---- optimality check of S.up(): natural residual |u - P_C(u - gamma g)|_inf, predicts the next face --
// returns true if converged; `changed` tells whether the predicted face differs from the current one
template <class WK, class G>
MPC_HD bool check_optimality(const Tron1Const& P, WK& S, const G& g, bool& changed, double& resid) {
    [[maybe_unused]] constexpr int N = WK::N;
#if defined(__CUDA_ARCH__) && defined(MPC_REDUCED_CERTIFICATE)
    // MEASURED TWICE AND NOT USED (profiles/r2_reduced_certificate.log): compile-time option only.
    // Interior face (Z = I, no fixed part: 99.7 % of the BASELINE config-2 instances) at full compact size: the gradient on the
    // stance variables is the residual of the linear system that was just solved, g_c = A w + f_c, and A -- the compact
    // Hessian build_hessian wrote -- is still intact in shared memory because the register elimination (gj_solve_regs /
    // gj3_solve_regs) never writes it back.  One symmetric mat-vec out of shared memory, fully unrolled with the same
    // predicated row/column loads the elimination uses to fill its window (all loads independent), replaces the matrix-free
    // input-response + adjoint pass.  Swing foot-steps are not read by the check.  Any other face needs the multipliers of
    // the fixed variables, i.e. the full gradient.  Measured: 48.6 us against 45.6 us per 4096-instance batch for this
    // unrolled version (47.2 us for a first version with lane-dependent loop bounds); horizon 20 7 % slower.
    if constexpr (G::kThreads >= WK::NC && WK::NC <= 60 && !WK::TILED) {
        if (S.interior && S.nc == WK::NC) {
            constexpr int NC = WK::NC;
            const int t = g.tid(), tt = t < NC ? t : NC - 1;
            const double* A = S.Ap();
            const double* rowp = A + MPC_PK(tt, 0);
            const double* colp = A + tt;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j += 2) {
                double v0, v1;
                if (j <= tt) v0 = rowp[j]; else v0 = colp[j * (j + 1) / 2];
                if (j + 1 <= tt) v1 = rowp[j + 1]; else v1 = colp[(j + 1) * (j + 2) / 2];
                a0 = fma(v0, S.w[j], a0);
                a1 = fma(v1, S.w[j + 1], a1);
            }
            if (t < NC) {
                const int s = S.cinv[t / 3], c = t - 3 * (t / 3);
                S.gp()[3 * s + c] = a0 + a1 + S.f[3 * s + c];
            }
            g.sync();
        } else gradient<WK>(P, S, g);
    } else
#endif
    gradient<WK>(P, S, g);
#if defined(__CUDA_ARCH__)
    if constexpr (G::kThreads == 32 && 2 * N <= 32) {
        // one foot-step per lane, reductions by warp shuffles
        const int s = g.tid();
        double r = 0.0, um = 1.0;
        bool ch = false;
        if (s < 2 * N && S.contact[s]) {
            double v[3], o[3];
            int ax, ay, zt;
            for (int c = 0; c < 3; ++c) v[c] = S.up()[3 * s + c] - P.gamma * S.gp()[3 * s + c];
            project_pyramid(P.mu, P.f_max, v, o, ax, ay, zt);
            for (int c = 0; c < 3; ++c) {
                double d = fabs(S.up()[3 * s + c] - o[c]);
                r = (!(d <= r) && r == r) ? d : r;   // NaN-propagating max
                double a = fabs(S.up()[3 * s + c]);
                um = a > um ? a : um;
            }
            S.nax[s] = (int8_t)ax; S.nay[s] = (int8_t)ay; S.nzt[s] = (int8_t)zt;
            ch = (ax != S.ax[s]) || (ay != S.ay[s]) || (zt != S.zt[s]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double r2 = __shfl_xor_sync(0xffffffffu, r, off), u2 = __shfl_xor_sync(0xffffffffu, um, off);
            r = (!(r2 <= r) && r == r) ? r2 : r;   // NaN-propagating max
            um = u2 > um ? u2 : um;
        }
        changed = __any_sync(0xffffffffu, ch);
        resid = r;
        g.sync();
        MPC_TICK(S, g, 12);
        return r <= P.tol * um;
    } else if constexpr (G::kThreads >= 2 * N && G::kThreads > 32) {
        // multi-warp groups: one foot-step per thread, warp-shuffle reductions, the per-warp results combined through
        // shared memory (S.resp() is free here), the face-change flag by a reducing barrier
        const int s = g.tid();
        double r = 0.0, um = 1.0;
        bool ch = false;
        if (s < 2 * N && S.contact[s]) {
            double v[3], o[3];
            int ax, ay, zt;
            for (int c = 0; c < 3; ++c) v[c] = S.up()[3 * s + c] - P.gamma * S.gp()[3 * s + c];
            project_pyramid(P.mu, P.f_max, v, o, ax, ay, zt);
            for (int c = 0; c < 3; ++c) {
                double d = fabs(S.up()[3 * s + c] - o[c]);
                r = (!(d <= r) && r == r) ? d : r;   // NaN-propagating max
                double a = fabs(S.up()[3 * s + c]);
                um = a > um ? a : um;
            }
            S.nax[s] = (int8_t)ax; S.nay[s] = (int8_t)ay; S.nzt[s] = (int8_t)zt;
            ch = (ax != S.ax[s]) || (ay != S.ay[s]) || (zt != S.zt[s]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double r2 = __shfl_xor_sync(0xffffffffu, r, off), u2 = __shfl_xor_sync(0xffffffffu, um, off);
            r = (!(r2 <= r) && r == r) ? r2 : r;
            um = u2 > um ? u2 : um;
        }
        constexpr int NW = G::kThreads / 32;
        static_assert(2 * NW <= 2 * N, "S.resp() holds the per-warp partial results");
        if ((s & 31) == 0) { S.resp()[2 * (s >> 5)] = r; S.resp()[2 * (s >> 5) + 1] = um; }
        changed = g.any(ch);                  // also the barrier that publishes S.resp()
        r = S.resp()[0]; um = S.resp()[1];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const double r2 = S.resp()[2 * w], u2 = S.resp()[2 * w + 1];
            r = (!(r2 <= r) && r == r) ? r2 : r;
            um = u2 > um ? u2 : um;
        }
        resid = r;
        g.sync();                             // S.resp() is reused by the next check
        MPC_TICK(S, g, 12);
        return r <= P.tol * um;
    } else if constexpr (G::kThreads == 32 && WK::RICCATI) {
        // one warp, more foot-steps than lanes: strided over the foot-steps, then the same shuffle reductions
        double r = 0.0, um = 1.0;
        bool ch = false;
        for (int s = g.tid(); s < 2 * N; s += 32) {
            if (!S.contact[s]) continue;
            double v[3], o[3];
            int ax, ay, zt;
            for (int c = 0; c < 3; ++c) v[c] = S.up()[3 * s + c] - P.gamma * S.gp()[3 * s + c];
            project_pyramid(P.mu, P.f_max, v, o, ax, ay, zt);
            for (int c = 0; c < 3; ++c) {
                double d = fabs(S.up()[3 * s + c] - o[c]);
                r = (!(d <= r) && r == r) ? d : r;   // NaN-propagating max
                double a = fabs(S.up()[3 * s + c]);
                um = a > um ? a : um;
            }
            S.nax[s] = (int8_t)ax; S.nay[s] = (int8_t)ay; S.nzt[s] = (int8_t)zt;
            ch = ch || (ax != S.ax[s]) || (ay != S.ay[s]) || (zt != S.zt[s]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double r2 = __shfl_xor_sync(0xffffffffu, r, off), u2 = __shfl_xor_sync(0xffffffffu, um, off);
            r = (!(r2 <= r) && r == r) ? r2 : r;   // NaN-propagating max
            um = u2 > um ? u2 : um;
        }
        changed = __any_sync(0xffffffffu, ch);
        resid = r;
        g.sync();
        MPC_TICK(S, g, 12);
        return r <= P.tol * um;
    } else
#endif
    {
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        double r = 0.0, um = 0.0;
        if (S.contact[s]) {
            double v[3], o[3];
            int ax, ay, zt;
            for (int c = 0; c < 3; ++c) v[c] = S.up()[3 * s + c] - P.gamma * S.gp()[3 * s + c];
            project_pyramid(P.mu, P.f_max, v, o, ax, ay, zt);
            for (int c = 0; c < 3; ++c) {
                double d = fabs(S.up()[3 * s + c] - o[c]);
                r = (!(d <= r) && r == r) ? d : r;   // NaN-propagating max
                double a = fabs(S.up()[3 * s + c]);
                um = a > um ? a : um;
            }
            S.nax[s] = (int8_t)ax; S.nay[s] = (int8_t)ay; S.nzt[s] = (int8_t)zt;
        }
        S.resp()[s] = r;
        S.tau()[s] = um;   // scratch: per foot-step |u|_inf
    }
    g.sync();
    double r = 0.0, um = 1.0;
    bool ch = false;
    for (int s = 0; s < 2 * N; ++s) {
        r = (!(S.resp()[s] <= r) && r == r) ? S.resp()[s] : r;   // NaN-propagating max
        um = S.tau()[s] > um ? S.tau()[s] : um;
        if (S.contact[s]) ch |= (S.nax[s] != S.ax[s]) || (S.nay[s] != S.ay[s]) || (S.nzt[s] != S.zt[s]);
    }
    g.sync();
    changed = ch;
    resid = r;
    MPC_TICK(S, g, 12);
    return r <= P.tol * um;
    }
}

template <class WK, class G>
MPC_HD void adopt_predicted_face(WK& S, const G& g) {
    [[maybe_unused]] constexpr int N = WK::N;
    for (int s = g.tid(); s < 2 * N; s += g.size()) { S.ax[s] = S.nax[s]; S.ay[s] = S.nay[s]; S.zt[s] = S.nzt[s]; }
    g.sync();
}

// ---- setup shared by solve and dump: inputs must already be in S.x0 / S.feet / S.contact ---------
template <class WK, class G>
MPC_HD void setup_instance(const Tron1Const& P, WK& S, const double* xref, const G& g, bool warm = false) {
    [[maybe_unused]] constexpr int N = WK::N;
#if defined(__CUDA_ARCH__)
    if constexpr (G::kThreads == 32 && 2 * N <= 32) {
        // one foot-step per lane: ranks from a ballot instead of a serial scan by one thread
        const int s = g.tid();
        const bool in = s < 2 * N && S.contact[s];
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        const int rank = __popc(mask & ((1u << s) - 1u));
        if (s < 2 * N) S.cidx[s] = (int16_t)rank;
        if (in) S.cinv[rank] = (int8_t)s;
        if (s == 0) S.nc = 3 * __popc(mask);
    } else if constexpr (G::kThreads >= 2 * N && G::kThreads > 32) {
        // multi-warp groups: per-warp ballots, warp offsets from the other warps' masks (through S.resp() as scratch)
        const int s = g.tid(), wid = s >> 5, lane = s & 31;
        const bool in = s < 2 * N && S.contact[s];
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        unsigned* masks = reinterpret_cast<unsigned*>(S.resp());
        if (lane == 0) masks[wid] = mask;
        g.sync();
        int base = 0, total = 0;
        constexpr int NW = G::kThreads / 32;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = __popc(masks[w]);
            if (w < wid) base += c;
            total += c;
        }
        const int rank = base + __popc(mask & ((1u << lane) - 1u));
        if (s < 2 * N) S.cidx[s] = (int16_t)rank;
        if (in) S.cinv[rank] = (int8_t)s;
        if (s == 0) S.nc = 3 * total;
        g.sync();                             // S.resp() is scratch again
    } else
#endif
    if (g.tid() == 0) {
        int c = 0;
        for (int s = 0; s < 2 * N; ++s) {
            S.cidx[s] = (int16_t)c;
            if (S.contact[s]) { S.cinv[c] = (int8_t)s; ++c; }
        }
        S.nc = 3 * c;
    }
    if (!warm) {
        for (int s = g.tid(); s < 2 * N; s += g.size()) { S.ax[s] = 0; S.ay[s] = 0; S.zt[s] = 0; }
    } else {
        // warm start (closed-loop rollout): the optimal face of the previous control step, shifted by one
        // horizon step, is the first guess; the new last step starts in the interior
        g.sync();
        for (int s = g.tid(); s < 2 * N; s += g.size()) {
            const bool in = s + 2 < 2 * N;
            S.nax[s] = in ? S.ax[s + 2] : (int8_t)0;
            S.nay[s] = in ? S.ay[s + 2] : (int8_t)0;
            S.nzt[s] = in ? S.zt[s + 2] : (int8_t)0;
        }
        g.sync();
        for (int s = g.tid(); s < 2 * N; s += g.size()) { S.ax[s] = S.nax[s]; S.ay[s] = S.nay[s]; S.zt[s] = S.nzt[s]; }
    }
    for (int k = g.tid(); k < N; k += g.size()) model_step<WK>(P, S, xref, k);
    g.sync();
    MPC_TICK(S, g, 0);
    horizon_sums<WK>(P, S, g);
    MPC_TICK(S, g, 1);
    free_response<WK>(P, S, xref, g);
    MPC_TICK(S, g, 2);
    if constexpr (!WK::RICCATI) adjoint<WK>(P, S, S.ee, S.f, g);   // f = 2 B' Q (A x0 - x_ref)   (src/QPSolver.cpp:59-60)
    MPC_TICK(S, g, 3);
}

// ---- full QP solve of one instance.  On exit S.up() holds the forces (full layout). -----------------
// iters = face solves + ADMM iterations.
// `after_setup` runs once the staged inputs (xref, S.x0, S.feet) are dead: everything the iterations need has been
// condensed into S by then, so a persistent kernel starts fetching its next instance into the same staging area there.
struct NoHook { MPC_HD void operator()() const {} };
// `restage` (Riccati work type only) brings the instance's inputs (S.x0, S.feet, xref) back before a further active-face
// iteration: that work type keeps no copy of the free-response error once the forward sweep has run, and its caller has
// reused the input staging area in between.
template <class WK, class G, class Hook = NoHook, class Hook2 = NoHook>
MPC_HD int solve_instance(const Tron1Const& P, WK& S, const double* xref, const G& g, int& iters, bool warm = false,
                          Hook after_setup = Hook(), Hook2 restage = Hook2()) {
    [[maybe_unused]] constexpr int N = WK::N;
    setup_instance<WK>(P, S, xref, g, warm);
    after_setup();
    iters = 0;
    // non-finite inputs (NaN/inf state, reference or feet) poison f: report failure instead of iterating
    // (a NaN would otherwise slip through the max-reductions of the optimality check)
    // (a non-finite H is caught by the factorisation's pivot test)
    bool bad;
    {
        double acc = 0.0;
        if constexpr (WK::RICCATI) {      // f is not formed: the free-response error and the lever-arm matrices carry the same inputs
            for (int i = g.tid(); i < 12 * (N + 1); i += g.size()) acc += S.ee[i] * 0.0;
            for (int i = g.tid(); i < 18 * N; i += g.size()) acc += S.W[i] * 0.0;
        } else
        for (int i = g.tid(); i < 6 * N; i += g.size()) acc += S.f[i] * 0.0;   // 0 unless f[i] is NaN/inf
        bad = (acc != 0.0 || acc != acc);
    }
#if defined(__CUDA_ARCH__)
    if constexpr (G::kThreads == 32) bad = __any_sync(0xffffffffu, bad);
    else
#endif
    {
        if (g.tid() == 0) S.flag = 0;
        g.sync();
        if (bad) S.flag = 1;
        g.sync();
        bad = S.flag != 0;
        g.sync();
    }
    if (bad) {
        for (int i = g.tid(); i < 6 * N; i += g.size()) S.up()[i] = 0.0;
        g.sync();
        return ST_FAILED;
    }
    if (S.nc == 0) {
        for (int i = g.tid(); i < 6 * N; i += g.size()) S.up()[i] = 0.0;
        g.sync();
        return ST_SOLVED;
    }
    bool changed;
    double resid;
    // phase A: active-face (semismooth Newton) iterations, cold start from the interior face
    for (int it = 0; it < P.max_newton; ++it) {
        ++iters;
        if constexpr (WK::RICCATI) {
            if (it > 0) {         // S.ee holds e0 + B u of the previous face: rebuild e0
                restage();
                free_response<WK>(P, S, xref, g);
            }
        }
        if (!face_solve<WK>(P, S, g)) return ST_FAILED;
        if (check_optimality<WK>(P, S, g, changed, resid)) return ST_SOLVED;
        if (!changed) break;   // same face predicted but not optimal: numerical stall -> ADMM
        adopt_predicted_face<WK>(S, g);
    }
    if constexpr (WK::RICCATI) {
        return ST_DEFER;     // the dense class (factorisation + ADMM) takes over from scratch
    } else {
    // phase B: ADMM on  min q(u) + I_C(z), u = z  with periodic active-face polish
    const int n = S.nc;
    double hmax = 0.0;
    build_hessian<WK>(P, S, 0.0, false, g);
    for (int i = 0; i < n; ++i) hmax = S.Ap()[WK::pk(i, i)] > hmax ? S.Ap()[WK::pk(i, i)] : hmax;
    g.sync();
    const double rho = sqrt(2.0 * P.r * hmax * 4.0);
    // start from the projection of the last face solution
    for (int s = g.tid(); s < 2 * N; s += g.size()) {
        if (!S.contact[s]) continue;
        double o[3];
        int ax, ay, zt;
        project_pyramid(P.mu, P.f_max, S.up() + 3 * s, o, ax, ay, zt);
        for (int c = 0; c < 3; ++c) { S.z[3 * S.cidx[s] + c] = o[c]; S.y[3 * S.cidx[s] + c] = 0.0; }
    }
    g.sync();
    bool have_factor = false;
    int stable = 0, since_polish = 0;
    for (int it = 0; it < P.max_admm; ++it) {
        ++iters;
        if (!have_factor) {
            build_hessian<WK>(P, S, rho, false, g);
            bool okf;
            if constexpr (WK::TILED) {
                tiled_pad<WK>(S, g);
                for (int i = g.tid(); i < WK::NC; i += g.size()) S.Ap()[WK::pk(WK::NC, i)] = 0.0;
                g.sync();
                okf = chol_tiled<WK>(S, g);
            } else {
            for (int i = g.tid(); i <= n; i += g.size()) S.Ap()[MPC_PK(n, i)] = (i == n) ? 1.0 : 0.0;
            g.sync();
#if defined(__CUDA_ARCH__)
            if constexpr (G::kThreads >= WK::NC + 1 && WK::NC <= 60) okf = cholesky_regs<WK>(S, g);
            else
#endif
                okf = cholesky_with_rhs<WK>(S, g);
            }
            if (!okf) return ST_FAILED;
            have_factor = true;
        }
        for (int s = g.tid(); s < 2 * N; s += g.size()) {
            if (!S.contact[s]) continue;
            int b = 3 * S.cidx[s];
            for (int c = 0; c < 3; ++c) S.w[b + c] = rho * (S.z[b + c] - S.y[b + c]) - S.f[3 * s + c];
        }
        g.sync();
        if constexpr (WK::TILED) {
            tiled_forward<WK>(S, g);
            tiled_backward<WK>(S, g, false);
        } else
#if defined(__CUDA_ARCH__)
        if constexpr (G::kThreads == 32 && WK::NC <= 31) {
            forward_regs<WK>(S, g);
            backward_regs<WK>(S, g);
        } else
#endif
        {
            forward_solve<WK>(S, g);
            for (int i = g.tid(); i < n; i += g.size()) S.Ap()[MPC_PK(n, i)] = S.w[i];
            g.sync();
            backward_solve<WK>(S, g);
        }
        if (g.tid() == 0) S.flag = 0;
        g.sync();
        for (int s = g.tid(); s < 2 * N; s += g.size()) {
            if (!S.contact[s]) continue;
            int b = 3 * S.cidx[s];
            double v[3], o[3], uh[3];
            int ax, ay, zt;
            for (int c = 0; c < 3; ++c) {
                uh[c] = P.admm_alpha * S.w[b + c] + (1.0 - P.admm_alpha) * S.z[b + c];
                v[c] = uh[c] + S.y[b + c];
            }
            project_pyramid(P.mu, P.f_max, v, o, ax, ay, zt);
            for (int c = 0; c < 3; ++c) { S.y[b + c] += uh[c] - o[c]; S.z[b + c] = o[c]; }
            if (ax != S.ax[s] || ay != S.ay[s] || zt != S.zt[s]) S.flag = 1;
            S.ax[s] = (int8_t)ax; S.ay[s] = (int8_t)ay; S.zt[s] = (int8_t)zt;
        }
        g.sync();
        stable = S.flag ? 0 : stable + 1;
        ++since_polish;
        g.sync();
        if ((stable >= 3 && since_polish >= 8) || since_polish >= 40) {
            since_polish = 0;
            have_factor = false;
            if (!face_solve<WK>(P, S, g)) return ST_FAILED;
            if (check_optimality<WK>(P, S, g, changed, resid)) return ST_SOLVED;
        }
    }
    // not certified: return the last ADMM iterate (feasible by construction)
    for (int s = g.tid(); s < 2 * N; s += g.size())
        for (int c = 0; c < 3; ++c) S.up()[3 * s + c] = S.contact[s] ? S.z[3 * S.cidx[s] + c] : 0.0;
    g.sync();
    return ST_MAXITER;
    }
}

// ---- reference trajectory of mpcQP (reference include/mpcQP.h:74-97): x_ref 13 x (N+1), step-major ------
template <class G>
MPC_HD void make_reference(const double* x0, double omega_yaw, double velocity_x, double Ts, int N, double* xr, const G& g) {
    for (int idx = g.tid(); idx < 13 * (N + 1); idx += g.size()) {
        const int i = idx / 13, c = idx % 13;
        const double t = (double)i * Ts;
        double v = x0[c];
        if (c == 2) v = x0[2] + t * omega_yaw;
        else if (c == 3) v = x0[3] + t * velocity_x;
        else if (c == 9) v = (i == 0) ? x0[9] : velocity_x;
        else if (c == 12) v = -9.8;
        xr[idx] = v;
    }
}

// nominal foot positions under the base: p_xy + Rz(yaw) offset_xy, on the ground (z = 0)
// (offsets: reference include/MPCParam.h:64-73; the same rule synth.py uses without the random term)
MPC_HD void nominal_feet(const double* x, const double* off_l, const double* off_r, double* feet) {
    double s, c;
    sincos(x[2], &s, &c);
    feet[0] = x[3] + c * off_l[0] - s * off_l[1];
    feet[1] = x[4] + s * off_l[0] + c * off_l[1];
    feet[2] = 0.0;
    feet[3] = x[3] + c * off_r[0] - s * off_r[1];
    feet[4] = x[4] + s * off_r[0] + c * off_r[1];
    feet[5] = 0.0;
}

// ---- plant update of the closed loop: x <- Ad x + Bd u0 (reference src/QPSolver.cpp:108-111) with the
// step-0 model already in S (closed-form ZOH).  u0 = S.up()[0..5].  One thread does the 12 updates.
template <class WK, class G>
MPC_HD void integrate_state(const Tron1Const& P, WK& S, double* x, const G& g) {
    if (g.tid() == 0) {
        const double Ts = P.Ts;
        const double* W0 = S.W;
        const double* u = S.up();
        double tau[3], phi[3];
        for (int i = 0; i < 3; ++i) {
            tau[i] = W0[i * 3] * u[0] + W0[i * 3 + 1] * u[1] + W0[i * 3 + 2] * u[2]
                   + W0[9 + i * 3] * u[3] + W0[9 + i * 3 + 1] * u[4] + W0[9 + i * 3 + 2] * u[5];
            phi[i] = (u[i] + u[3 + i]) * P.inv_m;
        }
        phi[2] += x[12];   // gravity state enters v_z_dot
        const double cz = S.cs[0], sz = S.cs[1];
        const double a0 = x[6] + 0.5 * Ts * tau[0], a1 = x[7] + 0.5 * Ts * tau[1], a2 = x[8] + 0.5 * Ts * tau[2];
        x[0] += Ts * (cz * a0 + sz * a1);
        x[1] += Ts * (-sz * a0 + cz * a1);
        x[2] += Ts * a2;
        for (int i = 0; i < 3; ++i) {
            x[3 + i] += Ts * (x[9 + i] + 0.5 * Ts * phi[i]);
            x[6 + i] += Ts * tau[i];
            x[9 + i] += Ts * phi[i];
        }
    }
    g.sync();
}

// ---- parity dump helpers (closed-form prediction matrices, column-major like Eigen) --------------
// A_aug: 13(N+1) x 13, block i = A_{i-1}...A_0   (src/QPSolver.cpp:36-40)
template <class WK>
MPC_HD double a_aug_entry(const Tron1Const& P, const WK& S, int i, int r, int c) {
    [[maybe_unused]] constexpr int N = WK::N;
    double v = (r == c) ? 1.0 : 0.0;
    double di = (double)i, Ts = P.Ts;
    if (r < 3 && c >= 6 && c < 9) {   // Theta <- omega : Ts C_i
        int rr = r, cc = c - 6;
        double C[9] = {S.cc[i], S.ss[i], 0, -S.ss[i], S.cc[i], 0, 0, 0, di};
        v += Ts * C[rr * 3 + cc];
    }
    if (r >= 3 && r < 6 && c == r + 6) v += di * Ts;              // p <- v
    if (r == 11 && c == 12) v += di * Ts;                         // v_z <- g
    if (r == 5 && c == 12) v += 0.5 * di * di * Ts * Ts;          // p_z <- g
    return v;
}
// B_aug block (i,j), i > j: 13 x 6 (src/QPSolver.cpp:42-47 generalised to the per-step model)
template <class WK>
MPC_HD double b_aug_entry(const Tron1Const& P, const WK& S, int i, int j, int r, int c) {
    [[maybe_unused]] constexpr int N = WK::N;
    if (i <= j) return 0.0;
    const double Ts = P.Ts;
    const int a = c / 3, cc = c % 3;
    const double* W = S.W + 18 * j + 9 * a;
    if (r < 3) {
        double sa = S.cc[i] + S.dc[j], sb = S.ss[i] + S.ds[j], sz = (double)(i - j) - 0.5;
        double Sm[9] = {sa, sb, 0, -sb, sa, 0, 0, 0, sz};
        double v = 0.0;
        for (int k = 0; k < 3; ++k) v += Sm[r * 3 + k] * W[k * 3 + cc];
        return Ts * Ts * v;
    }
    if (r < 6) return (r - 3 == cc) ? Ts * Ts * P.inv_m * ((double)(i - j) - 0.5) : 0.0;
    if (r < 9) return Ts * W[(r - 6) * 3 + cc];
    if (r < 12) return (r - 9 == cc) ? Ts * P.inv_m : 0.0;
    return 0.0;
}

}  // namespace mpcb200
