// kf_b200.cu -- sm_100a kernel and C ABI of the batched base-state Kalman filter (reference
// include/stateEstimator.h:184-337, a file that is in none of the reference's build targets).  One warp per robot,
// eight robots per CTA, the 12x12 / 14x14 workspaces in shared memory; see kf_core.cuh for the algebra.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/mpc_b200.h"
#include "kf_core.cuh"

using namespace mpcb200;

namespace {

constexpr int kRobotsPerCta = 8;

struct GrpWarp {
    int t;
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

__global__ void __launch_bounds__(32 * kRobotsPerCta)
kf_update_kernel(const __grid_constant__ KfParams K, const __grid_constant__ LegModel L, int B, double dt, const double* __restrict__ quat,
                 const double* __restrict__ gyro, const double* __restrict__ accel, const double* __restrict__ q,
                 const double* __restrict__ dq, const uint8_t* __restrict__ contact, double* __restrict__ xhat,
                 double* __restrict__ P, double* __restrict__ odom) {
    extern __shared__ __align__(16) unsigned char raw[];
    KfWork* works = reinterpret_cast<KfWork*>(raw);
    const int w = threadIdx.x / 32;
    const size_t b = (size_t)blockIdx.x * kRobotsPerCta + w;
    if (b >= (size_t)B) return;
    GrpWarp g{(int)(threadIdx.x % 32)};
    kf_update(K, L, dt, quat + 4 * b, gyro + 3 * b, accel + 3 * b, q + 6 * b, dq + 6 * b, contact + 2 * b, xhat + 12 * b, P + 144 * b,
              odom ? odom + 13 * b : nullptr, works[w], g);
}

__global__ void kf_reset_kernel(int B, double p0, double* __restrict__ xhat, double* __restrict__ P) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)B * 144; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % 144);
        P[i] = (e / 12 == e % 12) ? p0 : 0.0;
        if (e < 12) xhat[(i / 144) * 12 + e] = 0.0;
    }
}

KfParams to_kf(const mpc_b200_kf_params& p) {
    KfParams K;
    K.foot_radius = p.foot_radius; K.imu_noise_pos = p.imu_process_noise_position; K.imu_noise_vel = p.imu_process_noise_velocity;
    K.foot_noise_pos = p.foot_process_noise_position; K.foot_sensor_pos = p.foot_sensor_noise_position;
    K.foot_sensor_vel = p.foot_sensor_noise_velocity; K.foot_height_noise = p.foot_height_sensor_noise;
    K.suspect = p.high_suspect_number; K.accel_transpose = p.accel_transpose;
    return K;
}
LegModel to_model(const mpc_b200_leg_model& m) {
    LegModel M;
    memcpy(M.offset, m.offset, sizeof(M.offset));
    memcpy(M.axis, m.axis, sizeof(M.axis));
    return M;
}

}  // namespace

extern "C" {

int mpc_b200_kf_default_params(mpc_b200_kf_params* p) {
    if (!p) return MPC_B200_EINVAL;
    p->foot_radius = 0.02;                       // include/stateEstimator.h:124-130
    p->imu_process_noise_position = 0.02;
    p->imu_process_noise_velocity = 0.02;
    p->foot_process_noise_position = 0.002;
    p->foot_sensor_noise_position = 0.005;
    p->foot_sensor_noise_velocity = 0.1;
    p->foot_height_sensor_noise = 0.01;
    p->high_suspect_number = 100.0;              // :262
    p->accel_transpose = 1;                      // :281 as written
    return MPC_B200_OK;
}

int mpc_b200_kf_reset_device(int B, double p0, double* d_xhat, double* d_P, void* stream) {
    if (B < 1 || !d_xhat || !d_P) return MPC_B200_EINVAL;
    kf_reset_kernel<<<(int)(((size_t)B * 144 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, p0, d_xhat, d_P);
    return cudaGetLastError() == cudaSuccess ? MPC_B200_OK : MPC_B200_ECUDA;
}

int mpc_b200_kf_update_device(const mpc_b200_kf_params* p, const mpc_b200_leg_model* m, int B, double dt, const double* d_quat,
                              const double* d_gyro_local, const double* d_accel_local, const double* d_q, const double* d_dq,
                              const uint8_t* d_contact, double* d_xhat, double* d_P, double* d_odom, void* stream) {
    if (!p || !m || B < 1 || !(dt > 0.0) || !d_quat || !d_gyro_local || !d_accel_local || !d_q || !d_dq || !d_contact || !d_xhat || !d_P)
        return MPC_B200_EINVAL;
    const size_t smem = sizeof(KfWork) * kRobotsPerCta;
    static bool configured[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MPC_B200_ECUDA;
    if (!configured[dev & 63]) {
        if (cudaFuncSetAttribute(kf_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MPC_B200_ECUDA;
        configured[dev & 63] = true;
    }
    kf_update_kernel<<<(B + kRobotsPerCta - 1) / kRobotsPerCta, 32 * kRobotsPerCta, smem, (cudaStream_t)stream>>>(
        to_kf(*p), to_model(*m), B, dt, d_quat, d_gyro_local, d_accel_local, d_q, d_dq, d_contact, d_xhat, d_P, d_odom);
    return cudaGetLastError() == cudaSuccess ? MPC_B200_OK : MPC_B200_ECUDA;
}

int mpc_b200_kf_update_host(int device, const mpc_b200_kf_params* p, const mpc_b200_leg_model* m, int B, double dt, const double* quat,
                            const double* gyro_local, const double* accel_local, const double* q, const double* dq,
                            const uint8_t* contact, double* xhat, double* P, double* odom) {
    if (!p || !m || B < 1 || !quat || !gyro_local || !accel_local || !q || !dq || !contact || !xhat || !P) return MPC_B200_EINVAL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ENODEV; }
    if (device < 0 || device >= n) return MPC_B200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return MPC_B200_ECUDA;
    const size_t nb = (size_t)B, bytes = 8 * nb * (4 + 3 + 3 + 6 + 6 + 12 + 144 + 13) + 2 * nb + 256;
    unsigned char* d = nullptr;
    if (cudaMalloc((void**)&d, bytes) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ENOMEM; }
    double* dq4 = (double*)d; double* dg = dq4 + 4 * nb; double* da = dg + 3 * nb; double* dqq = da + 3 * nb; double* ddq = dqq + 6 * nb;
    double* dx = ddq + 6 * nb; double* dP = dx + 12 * nb; double* dod = dP + 144 * nb; uint8_t* dc = (uint8_t*)(dod + 13 * nb);
    bool ok = cudaMemcpy(dq4, quat, 32 * nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dg, gyro_local, 24 * nb, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(da, accel_local, 24 * nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dqq, q, 48 * nb, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(ddq, dq, 48 * nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dx, xhat, 96 * nb, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(dP, P, 1152 * nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(dc, contact, 2 * nb, cudaMemcpyHostToDevice) == cudaSuccess;
    int rc = ok ? mpc_b200_kf_update_device(p, m, B, dt, dq4, dg, da, dqq, ddq, dc, dx, dP, dod, nullptr) : MPC_B200_ECUDA;
    if (rc == MPC_B200_OK) {
        ok = cudaMemcpy(xhat, dx, 96 * nb, cudaMemcpyDeviceToHost) == cudaSuccess && cudaMemcpy(P, dP, 1152 * nb, cudaMemcpyDeviceToHost) == cudaSuccess &&
             (!odom || cudaMemcpy(odom, dod, 104 * nb, cudaMemcpyDeviceToHost) == cudaSuccess);
        if (!ok) rc = MPC_B200_ECUDA;
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"
