// chol_dmma_bench.cu -- prototype + microbenchmark of the tiled right-looking Cholesky with FP64 tensor-core
// (DMMA.8x8x4, mma.sync.aligned.m8n8k4.f64) trailing updates that the horizon-50 classes use, against the packed
// left-looking shared-memory Cholesky it replaces (tron1_core.cuh: cholesky_with_rhs).  Also measures DMMA and DFMA
// throughput per SM so that the comparison has a denominator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o chol_dmma_bench chol_dmma_bench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define PK(i, j) ((i) * ((i) + 1) / 2 + (j))
// tile layout: 8x8 tiles of the lower triangle, tile (I,J) at (I(I+1)/2+J)*64; inside a tile two 8x4 halves
// [half][row][4] so that an MMA A/B fragment load (row = lane/4, col = 4*half + lane%4) reads 32 consecutive doubles
__host__ __device__ inline int TL(int i, int j) {
    const int I = i >> 3, J = j >> 3;
    return ((I * (I + 1) / 2 + J) << 6) + ((j & 4) << 3) + ((i & 7) << 2) + (j & 3);
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- throughput probes ------------------------------------------------------------------------------------------
__global__ void dmma_peak(double* out, int reps) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1e-3 * threadIdx.x, b = 1e-3;
    for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
}
__global__ void dmma_latency(double* out, long long* cyc, int reps) {
    double c0 = threadIdx.x, c1 = 1.0, a = 1e-3, b = 1e-3;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) dmma(c0, c1, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
    if (c0 + c1 == 123.456) out[0] = c0;
}

// ---- the tiled factorisation ------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));       // MUFU.RCP64H, ~2^-23
    const double e = fma(-d, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);                                         // third-order step: ~2^-69
}
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));     // MUFU.RSQ64H, ~2^-23
    const double h = 0.5 * d;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}
// 8x8 diagonal tile: Cholesky factor L and its inverse X = L^-1, one warp.  Lane j (mod 8) owns COLUMN j of the symmetric
// tile (all 8 rows, so that a_jk is its own register k) and column j of X.  Column k: every lane gets column k by
// shuffles ONE STEP AHEAD (the old values, then applies the pending rank-1 update itself), so the pivot-to-pivot chain
// is rsqrt -> multiply -> fma with no shuffle on it; the forward substitution for X rides along (column-oriented:
// r -= l_k x_k).  D: 64 doubles in tile layout, in = A_KK (lower triangle valid), out = X (upper part zero).
// (A shuffle-free variant in which every lane eliminates the whole triangle redundantly measured 4x slower: 5.4 k cycles
// against 1.2 k for a lone warp -- it is bound by the FP64 issue rate of one warp, not by the dependency chain.)
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
    for (int i = 0; i < 8; ++i) c[i] = __shfl_sync(0xffffffffu, a[i], 0);   // column 0
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

// T: tile layout (shared or global memory), rows/cols 0..8*NT-1.  In: lower triangle of A (rows beyond the matrix are
// extra rows carried along, e.g. the right-hand side; their diagonal entries must stay positive).  Out: strictly-lower
// tiles hold L, diagonal tiles hold inv(L_KK).  pbuf: NT*64 + 64 doubles of shared memory (current panel column + the
// diagonal tile being factored).
// Look-ahead: the warp that owns the head of the trailing update does tile (K+1,K+1) first and factors it while the
// other warps work through the rest of the update, so the factor's latency chain is hidden whenever the update is long.
template <int NT, int NW, bool IDLE_MATES = false>
__device__ __forceinline__ bool chol_dmma_tiled(double* __restrict__ T, double* __restrict__ pbuf, int tid) {
    const int lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int fo = (gid << 2) + tig;                                     // fragment offset inside a half tile
    const int co = ((tig >> 1) << 5) + (gid << 2) + ((tig & 1) << 1);    // accumulator (C layout) offset inside a tile
    double* dbuf = pbuf + NT * 64;
    bool ok = true;
    if (warp == 0) {
        *reinterpret_cast<double2*>(dbuf + 2 * lane) = *reinterpret_cast<const double2*>(T + 2 * lane);
        __syncwarp();
        ok &= factor_diag8(dbuf, lane);
        *reinterpret_cast<double2*>(T + 2 * lane) = *reinterpret_cast<const double2*>(dbuf + 2 * lane);
    }
    __syncthreads();
    for (int K = 0; K < NT - 1; ++K) {
        // panel: L_IK = A_IK inv(L_KK)'  (two DMMAs per tile); goes to its final place and to the panel buffer
        {
            const double b_lo = dbuf[fo], b_hi = dbuf[32 + fo];
            for (int I = K + 1 + warp; I < NT; I += NW) {
                double* P = T + ((I * (I + 1) / 2 + K) << 6);
                const double a_lo = P[fo], a_hi = P[32 + fo];
                double c0 = 0.0, c1 = 0.0;
                dmma(c0, c1, a_lo, b_lo);
                dmma(c0, c1, a_hi, b_hi);
                __syncwarp();
                *reinterpret_cast<double2*>(P + co) = make_double2(c0, c1);
                *reinterpret_cast<double2*>(pbuf + (I << 6) + co) = make_double2(c0, c1);
            }
        }
        __syncthreads();
        // trailing update: A_IJ -= L_IK L_JK'  for K < J <= I.  Tile 0 (= the next diagonal tile) and the next HEAD tiles'
        // worth of time belong to warp 0; the remaining tiles, in row-major order, are cut into contiguous ranges (a range
        // mostly stays inside one tile row: its A fragments are reused); two tiles in flight per warp.
        const int m = NT - 1 - K, M = m * (m + 1) / 2;
        int t0, t1;
        {
            // warp 0 only looks ahead; with IDLE_MATES the other warps of its scheduler (warp % 4 == 0) sit the update out
            const int nwork = IDLE_MATES ? NW - NW / 4 : NW - 1;
            const int widx = IDLE_MATES ? warp - 1 - warp / 4 : warp - 1;
            const bool works = warp != 0 && (!IDLE_MATES || (warp & 3) != 0);
            t0 = works ? 1 + ((M - 1) * widx) / nwork : 0;
            t1 = works ? 1 + ((M - 1) * (widx + 1)) / nwork : 0;
        }
        if (warp == 0) {
            // next diagonal tile first, then factor it (overlaps the other warps' share of the update)
            double* Dn = T + (((K + 1) * (K + 2) / 2 + K + 1) << 6);
            const double al = -pbuf[((K + 1) << 6) + fo], ah = -pbuf[((K + 1) << 6) + 32 + fo];
            double2 c = *reinterpret_cast<double2*>(Dn + co);
            dmma(c.x, c.y, al, -al);
            dmma(c.x, c.y, ah, -ah);
            __syncwarp();
            *reinterpret_cast<double2*>(dbuf + co) = c;
            __syncwarp();
            ok &= factor_diag8(dbuf, lane);
            *reinterpret_cast<double2*>(Dn + 2 * lane) = *reinterpret_cast<const double2*>(dbuf + 2 * lane);
        }
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
                if (J == I) { ++I; J = K + 1; a_lo = -pbuf[(I < NT ? I : NT - 1) * 64 + fo]; a_hi = -pbuf[(I < NT ? I : NT - 1) * 64 + 32 + fo]; } else ++J;
                const bool two = t + 1 < t1;
                double* C1 = T + ((I * (I + 1) / 2 + J) << 6) + co;
                const double a1l = a_lo, a1h = a_hi;
                double b1l = 0.0, b1h = 0.0;
                double2 c1 = make_double2(0.0, 0.0);
                if (two) {
                    b1l = pbuf[(J << 6) + fo]; b1h = pbuf[(J << 6) + 32 + fo];
                    c1 = *reinterpret_cast<double2*>(C1);
                    if (J == I) { ++I; J = K + 1; a_lo = -pbuf[(I < NT ? I : NT - 1) * 64 + fo]; a_hi = -pbuf[(I < NT ? I : NT - 1) * 64 + 32 + fo]; } else ++J;
                }
                dmma(c0.x, c0.y, a0l, b0l);
                dmma(c1.x, c1.y, a1l, b1l);
                dmma(c0.x, c0.y, a0h, b0h);
                dmma(c1.x, c1.y, a1h, b1h);
                *reinterpret_cast<double2*>(C0) = c0;
                if (two) *reinterpret_cast<double2*>(C1) = c1;
            }
        }
        __syncthreads();
    }
    return ok;
}

// packed left-looking variant (the code it replaces; copied in shape from tron1_core.cuh for a like-for-like timing)
__device__ bool chol_packed(double* A, double* dinv, int n, int tid, int nthreads, int* flag) {
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double* rk = A + PK(k, 0);
        for (int i = k + 1 + tid; i <= n; i += nthreads) {
            double* ri = A + PK(i, 0);
            double s0 = ri[k], s1 = 0.0, d0 = rk[k], d1 = 0.0;
            int jj = 0;
            for (; jj + 1 < k; jj += 2) {
                double a0 = rk[jj], a1 = rk[jj + 1];
                s0 -= ri[jj] * a0; s1 -= ri[jj + 1] * a1;
                d0 -= a0 * a0; d1 -= a1 * a1;
            }
            if (jj < k) { double a0 = rk[jj]; s0 -= ri[jj] * a0; d0 -= a0 * a0; }
            double d = d0 + d1;
            if (!(d > 0.0)) { *flag = 1; d = 1.0; }
            double di = 1.0 / sqrt(d);
            ri[k] = (s0 + s1) * di;
            if (i == k + 1) dinv[k] = di;
        }
        __syncthreads();
    }
    return *flag == 0;
}

template <int NT, int NW, bool IM = false>
__global__ void __launch_bounds__(32 * NW, 1) k_tiled(const double* Ag, double* Lg, long long* cyc, int reps) {
    extern __shared__ double sm[];
    constexpr int SZ = (NT * (NT + 1) / 2) * 64;
    long long tot = 0;
    bool ok = true;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = threadIdx.x; i < SZ; i += blockDim.x) sm[i] = Ag[i];
        __syncthreads();
        long long t0 = clock64();
        ok &= chol_dmma_tiled<NT, NW, IM>(sm, sm + SZ, threadIdx.x);
        tot += clock64() - t0;
        __syncthreads();
    }
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < SZ; i += blockDim.x) Lg[i] = sm[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = ok ? tot / reps : -1;
}

template <int NT, int NW>
__global__ void __launch_bounds__(32 * NW, 1) k_tiled_global(const double* Ag, double* Wg, long long* cyc, int reps) {
    extern __shared__ double sm[];
    constexpr int SZ = (NT * (NT + 1) / 2) * 64;
    double* T = Wg + (size_t)blockIdx.x * SZ;
    long long tot = 0;
    bool ok = true;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = threadIdx.x; i < SZ; i += blockDim.x) T[i] = Ag[i];
        __syncthreads();
        long long t0 = clock64();
        ok &= chol_dmma_tiled<NT, NW>(T, sm, threadIdx.x);
        tot += clock64() - t0;
        __syncthreads();
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = ok ? tot / reps : -1;
}

template <int NC>
__global__ void __launch_bounds__(256, 1) k_packed(const double* Ag, double* Lg, long long* cyc, int reps) {
    extern __shared__ double sm[];
    constexpr int SZ = (NC + 1) * (NC + 2) / 2;
    double* dinv = sm + SZ;
    __shared__ int flag;
    long long tot = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = threadIdx.x; i < SZ; i += blockDim.x) sm[i] = Ag[i];
        __syncthreads();
        long long t0 = clock64();
        chol_packed(sm, dinv, NC, threadIdx.x, blockDim.x, &flag);
        tot += clock64() - t0;
        __syncthreads();
    }
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < SZ; i += blockDim.x) Lg[i] = sm[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = tot / reps;
}

__global__ void dfma_peak(double* out, int reps, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[threadIdx.x] = s;
}

template <int NT, int NW>
static void run_tiled_global(const std::vector<double>& Afull, int n, int NR, const std::vector<double>& Lref, int sms, int ctas_per_sm) {
    constexpr int SZ = (NT * (NT + 1) / 2) * 64;
    std::vector<double> h(SZ, 0.0);
    for (int i = 0; i < 8 * NT; ++i)
        for (int j = 0; j <= i; ++j) h[TL(i, j)] = (i < NR && j < NR) ? Afull[(size_t)i * NR + j] : (i == j ? 1.0 : 0.0);
    double *dA, *dW;
    long long* dc;
    const int grid = sms * ctas_per_sm;
    cudaMalloc(&dA, SZ * 8); cudaMalloc(&dW, (size_t)SZ * 8 * grid); cudaMalloc(&dc, 8 * 1024);
    cudaMemcpy(dA, h.data(), SZ * 8, cudaMemcpyHostToDevice);
    auto k = k_tiled_global<NT, NW>;
    k<<<grid, 32 * NW, (NT * 64 + 64) * 8>>>(dA, dW, dc, 4);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> L(SZ);
    long long c[2];
    cudaMemcpy(L.data(), dW, SZ * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
    double err = 0, scale = 0;
    for (int i = 0; i < NR; ++i)
        for (int j = 0; j < n && j <= i; ++j) {
            if ((i >> 3) == (j >> 3)) continue;
            double r = Lref[(size_t)i * NR + j];
            err = fmax(err, fabs(L[TL(i, j)] - r)); scale = fmax(scale, fabs(r));
        }
    printf("tiled DMMA (matrix in global/L2) NT=%d (n=%d) warps=%2d, %d CTA/SM: %7lld cycles per factorisation  max|L-Lref|/max|L| = %.2e  %s\n",
           NT, n, NW, ctas_per_sm, c[0], err / scale, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dW); cudaFree(dc);
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

template <int NT, int NW, bool IM = false>
static void run_tiled(const std::vector<double>& Afull, int n, int NR, const std::vector<double>& Lref, int sms) {
    constexpr int SZ = (NT * (NT + 1) / 2) * 64;
    std::vector<double> h(SZ, 0.0);
    for (int i = 0; i < 8 * NT; ++i)
        for (int j = 0; j <= i; ++j) h[TL(i, j)] = (i < NR && j < NR) ? Afull[(size_t)i * NR + j] : (i == j ? 1.0 : 0.0);
    double *dA, *dL;
    long long* dc;
    cudaMalloc(&dA, SZ * 8); cudaMalloc(&dL, SZ * 8); cudaMalloc(&dc, 8 * 1024);
    cudaMemcpy(dA, h.data(), SZ * 8, cudaMemcpyHostToDevice);
    auto k = k_tiled<NT, NW, IM>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (SZ + NT * 64 + 64) * 8);
    k<<<sms, 32 * NW, (SZ + NT * 64 + 64) * 8>>>(dA, dL, dc, 10);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> L(SZ);
    long long c[2];
    cudaMemcpy(L.data(), dL, SZ * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
    // compare strictly-lower tiles with the host factor; diagonal tiles hold inv(L_KK): check inv * L = I
    double err = 0, scale = 0;
    for (int i = 0; i < NR; ++i)
        for (int j = 0; j < n && j <= i; ++j) {
            if ((i >> 3) == (j >> 3)) continue;
            double r = Lref[(size_t)i * NR + j];
            err = fmax(err, fabs(L[TL(i, j)] - r)); scale = fmax(scale, fabs(r));
        }
    double ierr = 0;
    for (int K = 0; K * 8 < n; ++K)
        for (int i = 0; i < 8 && 8 * K + i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = 0;
                for (int m2 = j; m2 <= i; ++m2) s += L[TL(8 * K + i, 8 * K + m2)] * Lref[(size_t)(8 * K + m2) * NR + 8 * K + j];
                ierr = fmax(ierr, fabs(s - (i == j ? 1.0 : 0.0)));
            }
    printf("tiled DMMA %s NT=%d (n=%d) warps=%2d: %7lld cycles per factorisation  max|L-Lref|/max|L| = %.2e  |inv(Lkk) Lkk - I| = %.2e  %s\n",
           IM ? "(idle mates)" : "", NT, n, NW, c[0], err / scale, ierr, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dL); cudaFree(dc);
}

int main() {
    int dev = 0, sms = 148;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    sms = prop.multiProcessorCount;
    double* dout; long long* dc;
    cudaMalloc(&dout, 8 * 4096); cudaMalloc(&dc, 64);
    const double clk = prop.clockRate * 1e3;
    // ---- pipe throughput
    for (int wps : {4, 8, 16}) {
        const int reps = 20000;
        float ms = time_ms([&] { dmma_peak<<<sms, 32 * wps>>>(dout, reps); });
        double fl = 2.0 * 256 * 8 * (double)reps * wps * sms;
        printf("DMMA.8x8x4 %2d warps/SM: %.2f TFLOP/s  (%.1f cycles per DMMA per SM sub-partition at %.0f MHz nominal)\n", wps,
               fl / (ms * 1e-3) / 1e12, (ms * 1e-3) * clk / ((double)reps * 8 * wps / 4.0), clk / 1e6);
        ms = time_ms([&] { dfma_peak<<<sms, 32 * wps>>>(dout, reps / 16, 0.999999, 1e-9); });
        fl = 2.0 * 32 * 8 * 16 * (double)(reps / 16) * wps * sms;
        printf("DFMA       %2d warps/SM: %.2f TFLOP/s\n", wps, fl / (ms * 1e-3) / 1e12);
    }
    dmma_latency<<<1, 32>>>(dout, dc, 4096);
    cudaDeviceSynchronize();
    long long lat;
    cudaMemcpy(&lat, dc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA.8x8x4 dependent-issue latency: %lld cycles\n", lat);

    // ---- factorisation: n = 150 (+ rhs row), the single-stance class of horizon 50
    for (int n : {150, 300}) {
        const int NR = n + 1;
        std::vector<double> A((size_t)NR * NR, 0.0), L((size_t)NR * NR, 0.0);
        srand(7);
        std::vector<double> G((size_t)n * 8);
        for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = (i == j) ? 0.2 : 0.0;
                for (int k = 0; k < 8; ++k) s += G[i * 8 + k] * G[j * 8 + k] * (1.0 + 30.0 * k);
                A[(size_t)i * NR + j] = s;
            }
        for (int j = 0; j < n; ++j) A[(size_t)n * NR + j] = rand() / (double)RAND_MAX - 0.5;   // rhs row
        A[(size_t)n * NR + n] = 1e30;
        L = A;
        for (int k = 0; k < n; ++k) {
            double d = sqrt(L[(size_t)k * NR + k]);
            L[(size_t)k * NR + k] = d;
            for (int i = k + 1; i < NR; ++i) L[(size_t)i * NR + k] /= d;
            for (int j = k + 1; j < NR; ++j)
                for (int i = j; i < NR; ++i) L[(size_t)i * NR + j] -= L[(size_t)i * NR + k] * L[(size_t)j * NR + k];
        }
        if (n == 300) {
            run_tiled_global<38, 8>(A, n, NR, L, sms, 1);
            run_tiled_global<38, 16>(A, n, NR, L, sms, 1);
            run_tiled_global<38, 8>(A, n, NR, L, sms, 2);
            run_tiled_global<38, 16>(A, n, NR, L, sms, 2);
            run_tiled_global<38, 8>(A, n, NR, L, sms, 4);
        }
        if (n == 150) {
            run_tiled_global<19, 8>(A, n, NR, L, sms, 2);
            run_tiled<19, 4>(A, n, NR, L, sms);
            run_tiled<19, 8>(A, n, NR, L, sms);
            run_tiled<19, 16>(A, n, NR, L, sms);
            run_tiled<19, 8, true>(A, n, NR, L, sms);
            run_tiled<19, 16, true>(A, n, NR, L, sms);
            // packed left-looking baseline
            constexpr int NC = 150, SZ = (NC + 1) * (NC + 2) / 2;
            std::vector<double> h(SZ);
            for (int i = 0; i <= NC; ++i) for (int j = 0; j <= i; ++j) h[PK(i, j)] = A[(size_t)i * NR + j];
            double *dA, *dL;
            cudaMalloc(&dA, SZ * 8); cudaMalloc(&dL, SZ * 8);
            cudaMemcpy(dA, h.data(), SZ * 8, cudaMemcpyHostToDevice);
            auto k = k_packed<NC>;
            size_t smem = (SZ + NC + 2) * 8;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<sms, 256, smem>>>(dA, dL, dc, 4);
            cudaError_t e = cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            std::vector<double> Lp(SZ);
            cudaMemcpy(Lp.data(), dL, SZ * 8, cudaMemcpyDeviceToHost);
            double err = 0, scale = 0;
            for (int i = 0; i <= NC; ++i) for (int j = 0; j < NC && j <= i; ++j) { err = fmax(err, fabs(Lp[PK(i, j)] - L[(size_t)i * NR + j])); scale = fmax(scale, fabs(L[(size_t)i * NR + j])); }
            printf("packed left-looking DFMA n=150 warps= 8: %7lld cycles per factorisation  max|L-Lref|/max|L| = %.2e  %s\n", c, err / scale, cudaGetErrorString(e));
        }
    }
    return 0;
}
