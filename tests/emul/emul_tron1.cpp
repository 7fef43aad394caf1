// emul_tron1.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's group-cooperative device source (mpc_limx_control_b200/csrc/tron1_core.cuh)
// for the host with a one-thread "group", so the kernel mathematics can be checked against the
// oracle in the CPU test tier (no GPU in the build container).  It is NOT a CPU fallback: nothing
// under mpc_limx_control_b200/ or include/ links or loads this file, and the product library fails
// loudly without CUDA.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <new>

#include "../../mpc_limx_control_b200/csrc/tron1_core.cuh"
#include "../../mpc_limx_control_b200/csrc/tron1_params.h"

using namespace mpcb200;

struct GrpSerial {
    static constexpr int kThreads = 1;
    int tid() const { return 0; }
    int size() const { return 1; }
    void sync() const {}
};

template <int N, int NC, bool TILED = (N == 50)>
static int run_solve_nc(const Tron1Const& P, const double* x0, const double* xref, const double* feet,
                        const uint8_t* contact, double* forces, int* iters) {
    // same storage rule as the kernel wrapper (mpc_b200.cu: SolveWork): horizon 50 uses the tiled layout, so this build
    // checks the tile indexing, the blocked triangular solves and the padding (the tensor-core factorisation itself is
    // device code; chol_tiled_generic stands in with the same result layout)
    using Work = Tron1Work<N, NC, true, TILED>;
    auto* S = new Work();
    S->Aext = nullptr;
    S->x0 = x0;
    S->feet = feet;
    for (int s = 0; s < 2 * N; ++s) S->contact[s] = contact[s] ? 1 : 0;
    GrpSerial g;
    int it = 0;
    int st = solve_instance<Work>(P, *S, xref, g, it);
    std::memcpy(forces, S->up(), sizeof(double) * 6 * N);
    if (iters) *iters = it;
    delete S;
    return st;
}

// same routing as the kernel wrapper: small capacity class when the instance fits, else the large one
template <int N>
static int run_solve(const Tron1Const& P, const double* x0, const double* xref, const double* feet,
                     const uint8_t* contact, double* forces, int* iters) {
    int c = 0;
    for (int s = 0; s < 2 * N; ++s) c += contact[s] ? 1 : 0;
    if (3 * c <= 3 * N) return run_solve_nc<N, 3 * N>(P, x0, xref, feet, contact, forces, iters);
    return run_solve_nc<N, 6 * N>(P, x0, xref, feet, contact, forces, iters);
}

template <int N>
static void run_dump(const Tron1Const& P, const double* x0, const double* xref, const double* feet,
                     double* H, double* f, double* A_aug, double* B_aug) {
    using Work = Tron1Work<N, 6 * N>;
    auto* S = new Work();
    S->Aext = nullptr;
    S->x0 = x0;
    S->feet = feet;
    for (int s = 0; s < 2 * N; ++s) S->contact[s] = 1;
    GrpSerial g;
    setup_instance<Work>(P, *S, xref, g);
    build_hessian<Work>(P, *S, 0.0, false, g);
    const int n = 6 * N, p = 13 * (N + 1);
    if (H)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) { H[i + n * j] = S->Ap()[MPC_PK(i, j)]; H[j + n * i] = S->Ap()[MPC_PK(i, j)]; }
    if (f) std::memcpy(f, S->f, sizeof(double) * n);
    if (A_aug)
        for (int i = 0; i <= N; ++i)
            for (int r = 0; r < 13; ++r)
                for (int c = 0; c < 13; ++c) A_aug[(13 * i + r) + p * c] = a_aug_entry<Work>(P, *S, i, r, c);
    if (B_aug)
        for (int i = 0; i <= N; ++i)
            for (int j = 0; j < N; ++j)
                for (int r = 0; r < 13; ++r)
                    for (int c = 0; c < 6; ++c) B_aug[(13 * i + r) + p * (6 * j + c)] = b_aug_entry<Work>(P, *S, i, j, r, c);
    delete S;
}

template <int N, int NC>
static int run_rollout(const Tron1Const& P, int steps, double* x, double oy, double vx, int it0, double* u_traj, int* iters) {
    using Work = Tron1Work<N, NC>;
    auto* S = new Work();
    S->Aext = nullptr;
    double feet[6], xr[13 * (N + 1)];
    S->x0 = x;
    S->feet = feet;
    GrpSerial g;
    int bad = 0, tot = 0;
    for (int s = 0; s < steps; ++s) {
        nominal_feet(x, P.foot_off_l, P.foot_off_r, feet);
        make_reference(x, oy, vx, P.Ts, N, xr, g);
        for (int k = 0; k < N; ++k) {
            int l, r;
            gait_contact(P, it0 < 0 ? it0 : it0 + (s + k) * P.gait_mpc_step, l, r);
            S->contact[2 * k] = (int8_t)l;
            S->contact[2 * k + 1] = (int8_t)r;
        }
        int its = 0;
        int code = solve_instance<Work>(P, *S, xr, g, its, s > 0);
        if (u_traj) std::memcpy(u_traj + 6 * s, S->up(), sizeof(double) * 6);
        bad += code != 0;
        tot += its;
        integrate_state<Work>(P, *S, x, g);
    }
    if (iters) *iters = tot;
    delete S;
    return bad;
}

// the Riccati work type (O(N) active-face solves, no matrix); `deferred` reports whether the instance was handed to the
// dense class, as the kernel wrapper does for instances whose active-face iteration does not certify
template <int N, bool AINL>
static int run_solve_riccati(const Tron1Const& P, const double* x0, const double* xref, const double* feet,
                             const uint8_t* contact, double* forces, int* iters, int* deferred) {
    using Work = Tron1Work<N, 6 * N, AINL, false, true>;
    auto* S = new Work();
    std::vector<double> gains(AINL ? 1 : Work::ASZ);      // external gain storage (global memory on the device)
    S->Aext = AINL ? nullptr : gains.data();
    std::vector<double> adjs(18 * (N + 1));               // adjoint scratch (the dead input staging area on the device)
    S->adjx = adjs.data();
    S->x0 = x0;
    S->feet = feet;
    for (int s = 0; s < 2 * N; ++s) S->contact[s] = contact[s] ? 1 : 0;
    GrpSerial g;
    int it = 0;
    int st = solve_instance<Work>(P, *S, xref, g, it);
    std::memcpy(forces, S->up(), sizeof(double) * 6 * N);
    delete S;
    if (deferred) *deferred = (st == ST_DEFER);
    if (st == ST_DEFER) {
        int it2 = 0;
        st = run_solve<N>(P, x0, xref, feet, contact, forces, &it2);
        it += it2;
    }
    if (iters) *iters = it;
    return st;
}

extern "C" {

int emul_tron1_solve_riccati(const mpc_b200_tron1_params* prm, int N, int ext_gains, const double* x0, const double* xref,
                             const double* feet, const uint8_t* contact, double* forces, int* iters, int* deferred) {
    Tron1Const P;
    if (make_tron1_const(*prm, P)) return -1;
    switch (N) {
        case 4: return ext_gains ? run_solve_riccati<4, false>(P, x0, xref, feet, contact, forces, iters, deferred)
                                  : run_solve_riccati<4, true>(P, x0, xref, feet, contact, forces, iters, deferred);
        case 10: return ext_gains ? run_solve_riccati<10, false>(P, x0, xref, feet, contact, forces, iters, deferred)
                                  : run_solve_riccati<10, true>(P, x0, xref, feet, contact, forces, iters, deferred);
        case 20: return ext_gains ? run_solve_riccati<20, false>(P, x0, xref, feet, contact, forces, iters, deferred)
                                  : run_solve_riccati<20, true>(P, x0, xref, feet, contact, forces, iters, deferred);
        case 50: return ext_gains ? run_solve_riccati<50, false>(P, x0, xref, feet, contact, forces, iters, deferred)
                                  : run_solve_riccati<50, true>(P, x0, xref, feet, contact, forces, iters, deferred);
        default: return -2;
    }
}


int emul_tron1_rollout(const mpc_b200_tron1_params* prm, int N, int steps, double* x, double oy, double vx, int it0,
                       double* u_traj, int* iters) {
    Tron1Const P;
    if (make_tron1_const(*prm, P)) return -1;
    if (N != 10) return -2;
    return it0 < 0 ? run_rollout<10, 60>(P, steps, x, oy, vx, it0, u_traj, iters)
                   : run_rollout<10, 30>(P, steps, x, oy, vx, it0, u_traj, iters);
}

int emul_tron1_solve(const mpc_b200_tron1_params* prm, int N, const double* x0, const double* xref,
                     const double* feet, const uint8_t* contact, double* forces, int* iters) {
    Tron1Const P;
    if (make_tron1_const(*prm, P)) return -1;
    switch (N) {
        case 4: return run_solve<4>(P, x0, xref, feet, contact, forces, iters);
        case 10: return run_solve<10>(P, x0, xref, feet, contact, forces, iters);
        case 20: return run_solve<20>(P, x0, xref, feet, contact, forces, iters);
        case 50: return run_solve<50>(P, x0, xref, feet, contact, forces, iters);
        default: return -2;
    }
}

// the latency class of horizon 10 (mpc_b200.cu: MPC_N10_WPI_LAT) keeps the 60-variable system in the tiled layout
int emul_tron1_solve_tiled60(const mpc_b200_tron1_params* prm, const double* x0, const double* xref,
                             const double* feet, const uint8_t* contact, double* forces, int* iters) {
    Tron1Const P;
    if (make_tron1_const(*prm, P)) return -1;
    return run_solve_nc<10, 60, true>(P, x0, xref, feet, contact, forces, iters);
}

int emul_tron1_dump(const mpc_b200_tron1_params* prm, int N, const double* x0, const double* xref,
                    const double* feet, double* H, double* f, double* A_aug, double* B_aug) {
    Tron1Const P;
    if (make_tron1_const(*prm, P)) return -1;
    switch (N) {
        case 4: run_dump<4>(P, x0, xref, feet, H, f, A_aug, B_aug); return 0;
        case 10: run_dump<10>(P, x0, xref, feet, H, f, A_aug, B_aug); return 0;
        case 20: run_dump<20>(P, x0, xref, feet, H, f, A_aug, B_aug); return 0;
        case 50: run_dump<50>(P, x0, xref, feet, H, f, A_aug, B_aug); return 0;
        default: return -2;
    }
}

void emul_gait_contact(const mpc_b200_tron1_params* prm, int iter, int N, uint8_t* contact) {
    Tron1Const P;
    make_tron1_const(*prm, P);
    for (int k = 0; k < N; ++k) {
        int l, r;
        gait_contact(P, iter < 0 ? iter : iter + k * P.gait_mpc_step, l, r);
        contact[2 * k] = (uint8_t)l;
        contact[2 * k + 1] = (uint8_t)r;
    }
}
}
