// Isolated timing of the register Gauss-Jordan solve of csrc/tron1_core.cuh (one warp per system).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mpc_limx_control_b200/csrc -o gj_bench gj_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tron1_core.cuh"
using namespace mpcb200;
struct Grp {
    static constexpr int kThreads = 32;
    int t;
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
using Work = Tron1Work<10, 30, true>;
__global__ void kern(const double* Ag, double* out, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char raw[];
    Work* works = reinterpret_cast<Work*>(raw);
    const int wid = threadIdx.x / 32;
    Work& S = works[wid];
    Grp g; g.t = threadIdx.x % 32;
    S.nc = 30;
    long long tot = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = g.t; i < Work::PKN; i += 32) S.Astore[i] = Ag[i];
        __syncwarp();
        long long t0 = clock64();
        bool ok = true;
#if defined(__CUDA_ARCH__)
        ok = gj_solve_regs<Work>(S, g);
#endif
        tot += clock64() - t0;
        if (!ok) tot = -1;
    }
    if (g.t == 0) cyc[blockIdx.x * (blockDim.x / 32) + wid] = tot / reps;
    if (g.t < 30) out[(blockIdx.x * (blockDim.x / 32) + wid) * 32 + g.t] = S.w[g.t];
}
int main() {
    const int NC = 30, SZ = Work::PKN;
    static double h[SZ], Afull[NC][NC], b[NC];
    for (int i = 0; i <= NC; i++) for (int j = 0; j <= i; j++) h[MPC_PK(i, j)] = (i == j) ? 40.0 + i : 0.3 + 0.01 * (i + j);
    for (int i = 0; i < NC; i++) { for (int j = 0; j < NC; j++) Afull[i][j] = h[MPC_PK(i > j ? i : j, i > j ? j : i)]; b[i] = h[MPC_PK(NC, i)]; }
    double *dA, *dout; long long* dc;
    cudaMalloc(&dA, sizeof(h)); cudaMalloc(&dout, 8 * 32 * 4096); cudaMalloc(&dc, 8 * 4096);
    cudaMemcpy(dA, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int wpb : {1, 4}) for (int blocks : {148, 148 * 4}) {
        size_t smem = wpb * sizeof(Work);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, 32 * wpb, smem>>>(dA, dout, dc, 20);
        cudaDeviceSynchronize();
        long long hc[8]; double x[32];
        cudaMemcpy(hc, dc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(x, dout, 256, cudaMemcpyDeviceToHost);
        double res = 0;
        for (int i = 0; i < NC; i++) { double r = -b[i]; for (int j = 0; j < NC; j++) r += Afull[i][j] * x[j]; res = fmax(res, fabs(r)); }
        printf("warps/block %d blocks %d (=%d warps/SM): cycles per solve %lld  residual %.2e  %s\n", wpb, blocks, wpb * blocks / 148, hc[0], res,
               cudaGetErrorString(cudaGetLastError()));
    }
}
