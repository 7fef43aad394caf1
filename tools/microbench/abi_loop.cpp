// abi_loop.cpp -- times mpc_b200_tron1_solve_device back-to-back from C++ (no Python in the loop) to
// separate GPU time from host launch overhead.  usage: abi_loop B [steps]
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/mpc_b200.h"
int main(int argc, char** argv) {
    int B = argc > 1 ? atoi(argv[1]) : 4096, steps = argc > 2 ? atoi(argv[2]) : 500, N = 10;
    mpc_b200_tron1_params p; mpc_b200_tron1_default_params(&p);
    mpc_b200_engine* e; if (mpc_b200_create(&p, N, B, 0, &e)) { printf("create failed\n"); return 1; }
    std::vector<double> x0(13 * (size_t)B), xr(13 * 11 * (size_t)B), feet(6 * (size_t)B); std::vector<int32_t> it(B);
    for (int b = 0; b < B; ++b) {
        double* x = &x0[13 * (size_t)b];
        double yaw = 0.001 * b;
        double v[13] = {0.05, -0.03, yaw, 0.1, 0.2, 0.8, 0.1, -0.1, 0.05, 0.3, 0.1, 0, -9.8};
        for (int i = 0; i < 13; ++i) x[i] = v[i];
        for (int k = 0; k <= N; ++k) { double* r = &xr[(13 * 11) * (size_t)b + 13 * k]; for (int i = 0; i < 13; ++i) r[i] = x[i]; r[3] += 0.005 * k * 0.5; r[9] = k ? 0.5 : x[9]; }
        double c = cos(yaw), s = sin(yaw);
        feet[6 * (size_t)b + 0] = x[3] + c * -0.026 - s * -0.105; feet[6 * (size_t)b + 1] = x[4] + s * -0.026 + c * -0.105; feet[6 * (size_t)b + 2] = 0;
        feet[6 * (size_t)b + 3] = x[3] + c * -0.026 - s * 0.105; feet[6 * (size_t)b + 4] = x[4] + s * -0.026 + c * 0.105; feet[6 * (size_t)b + 5] = 0;
        it[b] = (b * 37) % 100000;
    }
    double *dx0, *dxr, *dfeet, *dF; int32_t *dit, *dst, *dits;
    cudaMalloc(&dx0, x0.size() * 8); cudaMalloc(&dxr, xr.size() * 8); cudaMalloc(&dfeet, feet.size() * 8); cudaMalloc(&dF, 60 * 8 * (size_t)B);
    cudaMalloc(&dit, 4 * (size_t)B); cudaMalloc(&dst, 4 * (size_t)B); cudaMalloc(&dits, 4 * (size_t)B);
    cudaMemcpy(dx0, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dxr, xr.data(), xr.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dfeet, feet.data(), feet.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dit, it.data(), 4 * (size_t)B, cudaMemcpyHostToDevice);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int i = 0; i < 20; ++i) mpc_b200_tron1_solve_device(e, B, dx0, dxr, dfeet, nullptr, dit, dF, dst, dits, s);
    cudaStreamSynchronize(s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(a, s);
    for (int i = 0; i < steps; ++i) mpc_b200_tron1_solve_device(e, B, dx0, dxr, dfeet, nullptr, dit, dF, dst, dits, s);
    cudaEventRecord(b, s);
    auto t1 = std::chrono::steady_clock::now();
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    std::vector<int32_t> st(B); cudaMemcpy(st.data(), dst, 4 * (size_t)B, cudaMemcpyDeviceToHost);
    int bad = 0; for (int v : st) bad += v != 0;
    printf("B %6d: %.2f us/step on device (%.1f M solves/s), host enqueue %.2f us/step, uncertified %d\n", B, 1e3 * ms / steps,
           B * (double)steps / (ms * 1e-3) / 1e6, std::chrono::duration<double, std::micro>(t1 - t0).count() / steps, bad);
    // host-buffer entry points with pinned memory (no Python in the loop)
    double *hx0, *hxr, *hfeet, *hF, *hoy, *hvx, *hu0; int32_t *hit, *hst, *hits;
    cudaHostAlloc((void**)&hx0, x0.size() * 8, 0); cudaHostAlloc((void**)&hxr, xr.size() * 8, 0); cudaHostAlloc((void**)&hfeet, feet.size() * 8, 0);
    cudaHostAlloc((void**)&hF, 60 * 8 * (size_t)B, 0); cudaHostAlloc((void**)&hit, 4 * (size_t)B, 0); cudaHostAlloc((void**)&hst, 4 * (size_t)B, 0);
    cudaHostAlloc((void**)&hits, 4 * (size_t)B, 0); cudaHostAlloc((void**)&hoy, 8 * (size_t)B, 0); cudaHostAlloc((void**)&hvx, 8 * (size_t)B, 0);
    cudaHostAlloc((void**)&hu0, 48 * (size_t)B, 0);
    memcpy(hx0, x0.data(), x0.size() * 8); memcpy(hxr, xr.data(), xr.size() * 8); memcpy(hfeet, feet.data(), feet.size() * 8);
    memcpy(hit, it.data(), 4 * (size_t)B);
    for (int b2 = 0; b2 < B; ++b2) { hoy[b2] = 0.1; hvx[b2] = 0.5; }
    int hs = steps / 4 + 5;
    for (int i = 0; i < 5; ++i) mpc_b200_tron1_solve_host(e, B, hx0, hxr, hfeet, nullptr, hit, hF, hst, hits);
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < hs; ++i) mpc_b200_tron1_solve_host(e, B, hx0, hxr, hfeet, nullptr, hit, hF, hst, hits);
    double us_full = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / hs;
    for (int i = 0; i < 5; ++i) mpc_b200_tron1_control_host(e, B, hx0, hoy, hvx, hfeet, nullptr, hit, hu0, hst, hits);
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < hs; ++i) mpc_b200_tron1_control_host(e, B, hx0, hoy, hvx, hfeet, nullptr, hit, hu0, hst, hits);
    double us_ctl = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / hs;
    printf("         solve_host %.1f us/call (%.1f M solves/s e2e)   control_host %.1f us/call (%.1f M solves/s e2e)\n", us_full,
           B / us_full, us_ctl, B / us_ctl);
    mpc_b200_destroy(e);
    return 0;
}
