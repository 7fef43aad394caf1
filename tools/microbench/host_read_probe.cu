// host_read_probe.cu -- how fast can G GPUs of one box pull bytes out of pinned host memory at the same time?
// Separates the two candidates for the end-to-end ceiling at 8 GPUs: the PCIe link of each GPU
// versus the host's memory / root complexes shared by all of them.  One process per GPU (launched side by side by
// tools/host_bw_probe.sh with CUDA_VISIBLE_DEVICES), each measuring, for a pinned buffer larger than the host LLC:
//   (a) SM-issued zero-copy reads (what the solve kernel's zero-copy path does), default-pinned and write-combined;
//   (b) cudaMemcpyAsync H2D through the copy engine;
//   (c) SM-issued zero-copy WRITES back to the host (the forces going home).
// All processes rendezvous on a wall-clock deadline so that the measurements overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o host_read_probe host_read_probe.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <cuda_runtime.h>

__global__ void read_kernel(const double4* __restrict__ src, size_t n, double* sink) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double4 v = src[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456) *sink = acc;
}
__global__ void write_kernel(double4* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_double4(1.0, 2.0, 3.0, (double)i);
}

static double now_s() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const double start_at = argc > 1 ? atof(argv[1]) : 0.0;   // wall-clock second at which every process starts measuring
    const char* tag = argc > 2 ? argv[2] : "gpu?";
    const size_t bytes = (size_t)256 << 20;
    void *h_def, *h_wc, *d;
    if (cudaHostAlloc(&h_def, bytes, cudaHostAllocMapped) != cudaSuccess || cudaHostAlloc(&h_wc, bytes, cudaHostAllocMapped | cudaHostAllocWriteCombined) != cudaSuccess ||
        cudaMalloc(&d, bytes) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
    memset(h_def, 1, bytes); memset(h_wc, 1, bytes);
    double* sink; cudaMalloc(&sink, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timed = [&](auto f) { f(); cudaDeviceSynchronize(); cudaEventRecord(a); for (int r = 0; r < 4; ++r) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return 4.0 * bytes / (ms * 1e-3) / 1e9; };
    while (start_at > 0.0 && now_s() < start_at) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    const size_t n = bytes / sizeof(double4);
    const double r_def = timed([&] { read_kernel<<<148 * 8, 256>>>((const double4*)h_def, n, sink); });
    const double r_wc = timed([&] { read_kernel<<<148 * 8, 256>>>((const double4*)h_wc, n, sink); });
    const double c_def = timed([&] { cudaMemcpyAsync(d, h_def, bytes, cudaMemcpyHostToDevice); });
    const double w_def = timed([&] { write_kernel<<<148 * 8, 256>>>((double4*)h_def, n); });
    printf("%s: zero-copy read %.1f GB/s (write-combined %.1f) | cudaMemcpyAsync H2D %.1f GB/s | zero-copy write %.1f GB/s  %s\n", tag, r_def, r_wc, c_def, w_def,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
