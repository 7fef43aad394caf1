// Dependent-issue latencies / issue intervals on sm_100a (one warp per scheduler, clock64 around N ops).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define PIN(x) asm volatile("" : "+d"(x)::"memory")
__device__ __forceinline__ double rcp_approx(double d) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); return y; }
__global__ void k(double* out, long long* cyc, double a, double b, int lane_src) {
    __shared__ __align__(16) double sm[128];
    const int t = threadIdx.x & 31;
    sm[t] = a; sm[t + 32] = b; sm[t + 64] = a; sm[t + 96] = b;
    __syncthreads();
    double x = a + t * 1e-9;
    long long t0, t1;
    int r = 0;
#define REC() do { t1 = clk(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r; } while (0)
    PIN(x); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    PIN(x); REC();                                         // 0 DFMA dependent
    PIN(x); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane_src + i) & 31) + 0.0 * i;
    PIN(x); REC();                                         // 1 SHFL(double)+ (dependent; DADD folded? see sass)
    int xi = t + lane_src;
    asm volatile("" : "+r"(xi)); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) xi = __shfl_sync(0xffffffffu, xi, xi & 31);
    asm volatile("" : "+r"(xi)); REC();                    // 2 SHFL(int) dependent (+LOP)
    PIN(x); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) x = rcp_approx(x);
    PIN(x); REC();                                         // 3 MUFU.RCP64H dependent
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3, y4 = x + 4, y5 = x + 5, y6 = x + 6, y7 = x + 7;
    PIN(y0); PIN(y1); PIN(y2); PIN(y3); PIN(y4); PIN(y5); PIN(y6); PIN(y7); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) { y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a); y4 = fma(y4, b, a); y5 = fma(y5, b, a); y6 = fma(y6, b, a); y7 = fma(y7, b, a); }
    PIN(y0); PIN(y1); PIN(y2); PIN(y3); PIN(y4); PIN(y5); PIN(y6); PIN(y7); REC();   // 4 8 independent DFMA (per group of 8)
    // 5: LDS.128 broadcast + 2 DFMA (the row-update pattern), 16 accumulators
    double w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = x + i;
    t0 = clk();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int p = 0; p < 16; p += 2) {
            const double2 c2 = *reinterpret_cast<const double2*>(sm + ((i & 3) * 16) + p);
            w[p] = fma(-b, c2.x, w[p]); w[p + 1] = fma(-b, c2.y, w[p + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) PIN(w[i]);
    REC();                                                 // 5 per (N/8) x [8 LDS.128 + 16 DFMA]
    PIN(x); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) { sm[t] = x; __syncwarp(); x = sm[(t + 1) & 31]; __syncwarp(); }
    PIN(x); REC();                                         // 6 STS+sync+LDS+sync
    PIN(x); t0 = clk();
#pragma unroll
    for (int i = 0; i < N; ++i) { x = (x > 0.0) ? x : 1.0; x = x * b; }
    PIN(x); REC();                                         // 7 DSETP+FSEL+DMUL
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += w[i];
    out[threadIdx.x] = x + y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7 + xi + s;
}
int main() {
    double* out; long long* cyc; long long h[16];
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 128);
    const char* nm[] = {"DFMA dependent", "SHFL(double) dependent", "SHFL(int)+LOP dependent", "MUFU.RCP64H dependent", "8 independent DFMA (per 8)",
                        "8 LDS.128 + 16 DFMA block", "STS+sync+LDS+sync", "DSETP+FSEL+DMUL dependent"};
    const double div[] = {N, N, N, N, N, N / 8, N, N};
    for (int warps = 1; warps <= 16; warps *= 4) {
        for (int r = 0; r < 2; ++r) k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, 3);
        cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
        printf("block of %d warp(s) (%d per scheduler)\n", warps, (warps + 3) / 4);
        for (int i = 0; i < 8; ++i) printf("  %-28s %7.2f cycles\n", nm[i], (double)h[i] / div[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
