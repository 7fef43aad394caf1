// lti_b200.cu -- kernels + C ABI of the generic condensed-MPC path (reference class QPSolver,
// include/QPSolver.h:13-37, src/QPSolver.cpp).  One CTA per problem instance, workspaces in HBM.
// Host-pointer entry points: the facade objects live on the caller's stack like the reference's.
#include <cuda_runtime.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/mpc_b200.h"
#include "lti_core.cuh"

using namespace mpcb200;

namespace {

struct GrpCta {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

__global__ void lti_discretize_kernel(int NX, int NU, double Ts, const double* Ac, const double* Bc, double* Ad,
                                      double* Bd, double* work) {
    const int b = blockIdx.x, m = NX + NU;
    GrpCta g;
    lti_discretize(NX, NU, Ts, Ac + (size_t)b * NX * NX, Bc + (size_t)b * NX * NU, Ad + (size_t)b * NX * NX,
                   Bd + (size_t)b * NX * NU, work + (size_t)b * 3 * m * m, g);
}

struct BuildPtrs {
    const double *Ad, *Bd, *Q, *R, *P, *x_min, *x_max, *xi0, *xi_ref;
    double *A_aug, *B_aug, *H, *f, *A_eq, *b_eq, *lb, *ub, *A_ineq, *lbA, *ubA, *work;
    double u_min, u_max;
};

__global__ void lti_build_kernel(LtiDims d, BuildPtrs q) {
    const size_t b = blockIdx.x;
    const size_t NX = d.NX, N = d.N, p = d.p(), n = d.n(), mi = 2 * NX * N, me = NX * N;
    GrpCta g;
    auto off = [&](double* base, size_t stride) { return base ? base + b * stride : (double*)0; };
    lti_build(d, q.Ad, q.Bd, q.Q, q.R, q.P, q.x_min, q.x_max, q.u_min, q.u_max, q.xi0 + b * NX, q.xi_ref + b * p,
              q.A_aug + b * p * NX, q.B_aug + b * p * n, off(q.H, n * n), off(q.f, n), off(q.A_eq, me * n),
              off(q.b_eq, me), off(q.lb, n), off(q.ub, n), off(q.A_ineq, mi * n), off(q.lbA, mi), off(q.ubA, mi),
              q.work + b * (p * n + p), g);
}

__global__ void qp_dense_kernel(int n, int m, const double* H, const double* f, const double* A, const double* lb,
                                const double* ub, const double* lbA, const double* ubA, double* U, int32_t* status,
                                int32_t* iters, double* work, size_t work_stride, int max_newton, int max_admm, double tol) {
    const size_t b = blockIdx.x;
    GrpCta g;
    QpWork W = qp_dense_carve(work + b * work_stride, n, m);
    int its = 0;
    int st = qp_dense_solve(n, H + b * n * n, f + b * n, m, m ? A + b * (size_t)m * n : (const double*)0, lb + b * n,
                            ub + b * n, m ? lbA + b * m : (const double*)0, m ? ubA + b * m : (const double*)0,
                            U + b * n, &its, W, max_newton, max_admm, tol, g);
    if (threadIdx.x == 0) {
        if (status) status[b] = st;
        if (iters) iters[b] = its;
    }
}

__global__ void lti_update_kernel(int NX, int NU, const double* Ad, const double* Bd, double* xi, const double* u, double* work) {
    const size_t b = blockIdx.x;
    GrpCta g;
    lti_update(NX, NU, Ad, Bd, xi + b * NX, u + b * NU, work + b * NX, g);
}

}  // namespace

struct mpc_b200_lti {
    int device = 0;
    cudaStream_t stream = nullptr;
    char* arena = nullptr;
    size_t cap = 0, used = 0;
    std::string err;
    int64_t launches = 0;
};

static int lti_err(mpc_b200_lti* c, int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (c) {
        c->err = what;
        if (ce != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(ce); }
    }
    return code;
}
#define LCU(c, call)                                                             \
    do {                                                                         \
        cudaError_t ce_ = (call);                                                \
        if (ce_ != cudaSuccess) return lti_err((c), MPC_B200_ECUDA, #call, ce_); \
    } while (0)

// grow-only device arena; reserve() must be called with the total before any take()
static int arena_reserve(mpc_b200_lti* c, size_t bytes) {
    c->used = 0;
    if (bytes <= c->cap) return MPC_B200_OK;
    if (c->arena) cudaFree(c->arena);
    c->arena = nullptr; c->cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&c->arena, want) != cudaSuccess) { cudaGetLastError(); return lti_err(c, MPC_B200_ENOMEM, "arena"); }
    c->cap = want;
    return MPC_B200_OK;
}
template <class T>
static T* arena_take(mpc_b200_lti* c, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(c->arena + c->used);
    c->used += bytes;
    return p;
}
static size_t a256(size_t count, size_t elem = 8) { return (count * elem + 255) & ~size_t(255); }

template <class T>
static T* up(mpc_b200_lti* c, const T* host, size_t count, cudaError_t& ce) {
    if (!host) return nullptr;
    T* d = arena_take<T>(c, count);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    return d;
}
template <class T>
static void down(mpc_b200_lti* c, T* host, const T* dev, size_t count, cudaError_t& ce) {
    if (host && ce == cudaSuccess) ce = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream);
}

extern "C" {

int mpc_b200_lti_create(int device, mpc_b200_lti** out) {
    if (!out) return MPC_B200_EINVAL;
    *out = nullptr;
    if (device < 0 || device >= mpc_b200_device_count()) return MPC_B200_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return MPC_B200_ENODEV;
    mpc_b200_lti* c = new (std::nothrow) mpc_b200_lti();
    if (!c) return MPC_B200_ENOMEM;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return MPC_B200_ECUDA;
    }
    *out = c;
    return MPC_B200_OK;
}

int mpc_b200_lti_destroy(mpc_b200_lti* c) {
    if (!c) return MPC_B200_EINVAL;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->arena) cudaFree(c->arena);
    delete c;
    return MPC_B200_OK;
}

const char* mpc_b200_lti_last_error(const mpc_b200_lti* c) { return c ? c->err.c_str() : ""; }
int64_t mpc_b200_lti_launch_count(const mpc_b200_lti* c) { return c ? c->launches : 0; }

int mpc_b200_lti_discretize(mpc_b200_lti* c, int B, int NX, int NU, double Ts, const double* Ac, const double* Bc,
                            double* Ad, double* Bd) {
    if (!c || !Ac || !Bc || !Ad || !Bd || B < 1 || NX < 1 || NU < 1 || NX + NU > 64 || !(Ts > 0.0))
        return lti_err(c, MPC_B200_EINVAL, "discretize: bad argument");
    LCU(c, cudaSetDevice(c->device));
    const size_t m = NX + NU, aa = (size_t)B * NX * NX, ab = (size_t)B * NX * NU;
    int rc = arena_reserve(c, 2 * a256(aa) + 2 * a256(ab) + a256((size_t)B * 3 * m * m));
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    double* dAc = up(c, Ac, aa, ce); double* dBc = up(c, Bc, ab, ce);
    double* dAd = arena_take<double>(c, aa); double* dBd = arena_take<double>(c, ab);
    double* work = arena_take<double>(c, (size_t)B * 3 * m * m);
    LCU(c, ce);
    lti_discretize_kernel<<<B, 128, 0, c->stream>>>(NX, NU, Ts, dAc, dBc, dAd, dBd, work);
    LCU(c, cudaGetLastError());
    c->launches++;
    down(c, Ad, dAd, aa, ce); down(c, Bd, dBd, ab, ce);
    LCU(c, ce);
    LCU(c, cudaStreamSynchronize(c->stream));
    return MPC_B200_OK;
}

int mpc_b200_lti_build_qp(mpc_b200_lti* c, int B, int NX, int NU, int N, const double* Ad, const double* Bd,
                          const double* Q, const double* R, const double* P, const double* x_min, const double* x_max,
                          double u_min, double u_max, const double* xi0, const double* xi_ref, double* H, double* f,
                          double* A_eq, double* b_eq, double* lb, double* ub, double* A_ineq, double* lbA, double* ubA,
                          double* A_aug, double* B_aug) {
    if (!c || !Ad || !Bd || !Q || !R || !P || !x_min || !x_max || !xi0 || !xi_ref || B < 1 || NX < 1 || NU < 1 || N < 1)
        return lti_err(c, MPC_B200_EINVAL, "build_qp: bad argument");
    LCU(c, cudaSetDevice(c->device));
    LtiDims d{NX, NU, N};
    const size_t p = d.p(), n = d.n(), mi = 2 * (size_t)NX * N, me = (size_t)NX * N, Bs = B;
    size_t total = a256((size_t)NX * NX) * 3 + a256((size_t)NX * NU) + a256((size_t)NU * NU) + 2 * a256(NX) +
                   a256(Bs * NX) + a256(Bs * p) + a256(Bs * p * NX) + a256(Bs * p * n) + a256(Bs * n * n) + 3 * a256(Bs * n) +
                   a256(Bs * me * n) + a256(Bs * me) + a256(Bs * mi * n) + 2 * a256(Bs * mi) + a256(Bs * (p * n + p));
    int rc = arena_reserve(c, total);
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    BuildPtrs q;
    q.Ad = up(c, Ad, (size_t)NX * NX, ce); q.Bd = up(c, Bd, (size_t)NX * NU, ce);
    q.Q = up(c, Q, (size_t)NX * NX, ce); q.R = up(c, R, (size_t)NU * NU, ce); q.P = up(c, P, (size_t)NX * NX, ce);
    q.x_min = up(c, x_min, NX, ce); q.x_max = up(c, x_max, NX, ce);
    q.xi0 = up(c, xi0, Bs * NX, ce); q.xi_ref = up(c, xi_ref, Bs * p, ce);
    q.A_aug = arena_take<double>(c, Bs * p * NX); q.B_aug = arena_take<double>(c, Bs * p * n);
    q.H = H ? arena_take<double>(c, Bs * n * n) : nullptr; q.f = f ? arena_take<double>(c, Bs * n) : nullptr;
    q.A_eq = A_eq ? arena_take<double>(c, Bs * me * n) : nullptr; q.b_eq = b_eq ? arena_take<double>(c, Bs * me) : nullptr;
    q.lb = lb ? arena_take<double>(c, Bs * n) : nullptr; q.ub = ub ? arena_take<double>(c, Bs * n) : nullptr;
    q.A_ineq = A_ineq ? arena_take<double>(c, Bs * mi * n) : nullptr;
    q.lbA = lbA ? arena_take<double>(c, Bs * mi) : nullptr; q.ubA = ubA ? arena_take<double>(c, Bs * mi) : nullptr;
    q.work = arena_take<double>(c, Bs * (p * n + p));
    q.u_min = u_min; q.u_max = u_max;
    LCU(c, ce);
    lti_build_kernel<<<B, 256, 0, c->stream>>>(d, q);
    LCU(c, cudaGetLastError());
    c->launches++;
    down(c, H, q.H, Bs * n * n, ce); down(c, f, q.f, Bs * n, ce);
    down(c, A_eq, q.A_eq, Bs * me * n, ce); down(c, b_eq, q.b_eq, Bs * me, ce);
    down(c, lb, q.lb, Bs * n, ce); down(c, ub, q.ub, Bs * n, ce);
    down(c, A_ineq, q.A_ineq, Bs * mi * n, ce); down(c, lbA, q.lbA, Bs * mi, ce); down(c, ubA, q.ubA, Bs * mi, ce);
    down(c, A_aug, q.A_aug, Bs * p * NX, ce); down(c, B_aug, q.B_aug, Bs * p * n, ce);
    LCU(c, ce);
    LCU(c, cudaStreamSynchronize(c->stream));
    return MPC_B200_OK;
}

int mpc_b200_qp_solve_dense(mpc_b200_lti* c, int B, int n, int m, const double* H, const double* f, const double* A,
                            const double* lb, const double* ub, const double* lbA, const double* ubA, double* U,
                            int32_t* status, int32_t* iters) {
    if (!c || !H || !f || !lb || !ub || !U || B < 1 || n < 1 || m < 0 || (m > 0 && (!A || !lbA || !ubA)))
        return lti_err(c, MPC_B200_EINVAL, "qp_solve_dense: bad argument");
    LCU(c, cudaSetDevice(c->device));
    const size_t Bs = B, ws = qp_dense_work_doubles(n, m);
    size_t total = a256(Bs * n * n) + 4 * a256(Bs * n) + a256(Bs * m * n) + 2 * a256(Bs * m) + 2 * a256(Bs, 4) + a256(Bs * ws) + 4096;
    int rc = arena_reserve(c, total);
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    double* dH = up(c, H, Bs * n * n, ce); double* df = up(c, f, Bs * n, ce);
    double* dA = m ? up(c, A, Bs * m * n, ce) : nullptr;
    double* dlb = up(c, lb, Bs * n, ce); double* dub = up(c, ub, Bs * n, ce);
    double* dlbA = m ? up(c, lbA, Bs * m, ce) : nullptr; double* dubA = m ? up(c, ubA, Bs * m, ce) : nullptr;
    double* dU = arena_take<double>(c, Bs * n);
    int32_t* dst = arena_take<int32_t>(c, Bs); int32_t* dit = arena_take<int32_t>(c, Bs);
    double* work = arena_take<double>(c, Bs * ws);
    LCU(c, ce);
    qp_dense_kernel<<<B, 256, 0, c->stream>>>(n, m, dH, df, dA, dlb, dub, dlbA, dubA, dU, dst, dit, work, ws, 30, 4000, 1e-9);
    LCU(c, cudaGetLastError());
    c->launches++;
    down(c, U, dU, Bs * n, ce); down(c, status, dst, Bs, ce); down(c, iters, dit, Bs, ce);
    LCU(c, ce);
    LCU(c, cudaStreamSynchronize(c->stream));
    return MPC_B200_OK;
}

int mpc_b200_lti_update_state(mpc_b200_lti* c, int B, int NX, int NU, const double* Ad, const double* Bd, double* xi,
                              const double* u) {
    if (!c || !Ad || !Bd || !xi || !u || B < 1 || NX < 1 || NU < 1) return lti_err(c, MPC_B200_EINVAL, "update_state: bad argument");
    LCU(c, cudaSetDevice(c->device));
    const size_t Bs = B;
    int rc = arena_reserve(c, a256((size_t)NX * NX) + a256((size_t)NX * NU) + 2 * a256(Bs * NX) + a256(Bs * NU));
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    double* dAd = up(c, Ad, (size_t)NX * NX, ce); double* dBd = up(c, Bd, (size_t)NX * NU, ce);
    double* dxi = up(c, (const double*)xi, Bs * NX, ce); double* du = up(c, u, Bs * NU, ce);
    double* work = arena_take<double>(c, Bs * NX);
    LCU(c, ce);
    lti_update_kernel<<<B, 32, 0, c->stream>>>(NX, NU, dAd, dBd, dxi, du, work);
    LCU(c, cudaGetLastError());
    c->launches++;
    down(c, xi, dxi, Bs * NX, ce);
    LCU(c, ce);
    LCU(c, cudaStreamSynchronize(c->stream));
    return MPC_B200_OK;
}

}  // extern "C"
