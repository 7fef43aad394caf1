// leg_b200.cu -- sm_100a kernels and C ABI of the leg kinematics around the force MPC:
// batched forward kinematics of the two point feet (the `feet` input of the solve), the swing-leg step of
// MPC::run (gait -> foot placement -> swing trajectory -> damped least-squares IK -> joint targets), and the
// mapping of the optimal ground-reaction forces to stance-leg joint torques (tau = -J' f).
//
// One thread per robot: every robot is ~1.5 kFLOP of dependent 3x3 algebra on ~150 bytes, so the kernels are
// bound by HBM traffic / latency, not by arithmetic.  Each CTA stages its contiguous slice of the
// instance-major arrays through shared memory with coalesced loads/stores (a thread reading its own
// 6-13 doubles directly would issue 48-104-byte-strided requests), then every thread works out of registers.
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/mpc_b200.h"
#include "leg_core.cuh"

using namespace mpcb200;

namespace {

constexpr int kThreads = 128;

// coalesced copy of `n` doubles global -> shared / shared -> global by the whole CTA
__device__ __forceinline__ void cta_load(double* dst, const double* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void cta_store(double* __restrict__ dst, const double* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

LegModel to_model(const mpc_b200_leg_model& m) {
    LegModel M;
    static_assert(sizeof(M.offset) == sizeof(m.offset) && sizeof(M.axis) == sizeof(m.axis), "layout");
    memcpy(M.offset, m.offset, sizeof(M.offset));
    memcpy(M.axis, m.axis, sizeof(M.axis));
    return M;
}
SwingParams to_swing(const mpc_b200_swing_params& p) {
    SwingParams S;
    S.dt = p.dt; S.swing_time = p.swing_time; S.stance_time = p.stance_time; S.gait_height = p.gait_height;
    S.p_rel_max = p.p_rel_max;
    for (int i = 0; i < 3; ++i) { S.foot_off_l[i] = p.foot_offset_left[i]; S.foot_off_r[i] = p.foot_offset_right[i]; }
    S.ik_tol = p.ik_tol; S.ik_dt = p.ik_dt; S.ik_damp = p.ik_damp; S.ik_max_iter = p.ik_max_iter;
    S.ik_mode = p.ik_mode;
    return S;
}

// ---- forward kinematics of both feet (reference include/pinocchio_kinematics.h:30-43,153-157) --------------------
__global__ void __launch_bounds__(kThreads)
leg_fk_kernel(const __grid_constant__ LegModel M, int B, const double* __restrict__ pos, const double* __restrict__ quat,
              const double* __restrict__ q, double* __restrict__ feet, double* __restrict__ jac) {
    __shared__ double s_in[kThreads * 13];    // pos 3 | quat 4 | q 6, array by array
    __shared__ double s_out[kThreads * 18];   // feet 6, then the Jacobians in a second pass
    const int first = blockIdx.x * kThreads, nb = min(kThreads, B - first), t = threadIdx.x;
    double* s_pos = s_in; double* s_quat = s_in + 3 * kThreads; double* s_q = s_in + 7 * kThreads;
    cta_load(s_pos, pos + 3 * (size_t)first, 3 * nb);
    cta_load(s_quat, quat + 4 * (size_t)first, 4 * nb);
    cta_load(s_q, q + 6 * (size_t)first, 6 * nb);
    __syncthreads();
    double J[18];
    if (t < nb) {
        double Rb[9];
        quat_to_rot(s_quat + 4 * t, Rb);
        for (int leg = 0; leg < 2; ++leg) leg_fk_world(M, leg, s_pos + 3 * t, Rb, s_q + 6 * t + 3 * leg, s_out + 6 * t + 3 * leg, jac ? J + 9 * leg : nullptr);
    }
    __syncthreads();
    cta_store(feet + 6 * (size_t)first, s_out, 6 * nb);
    if (jac) {
        __syncthreads();
        if (t < nb) for (int i = 0; i < 18; ++i) s_out[18 * t + i] = J[i];
        __syncthreads();
        cta_store(jac + 18 * (size_t)first, s_out, 18 * nb);
    }
}

// ---- swing-leg step of MPC::run (reference include/MPCController.h:61-75,106-175) ------------------------------------
__global__ void __launch_bounds__(kThreads)
swing_step_kernel(const __grid_constant__ LegModel M, const __grid_constant__ SwingParams P, int B, const double* __restrict__ pos,
                  const double* __restrict__ quat, const double* __restrict__ q, const double* __restrict__ des_vel,
                  const int32_t* __restrict__ iter, double* __restrict__ q_cmd, double* __restrict__ feet,
                  double* __restrict__ next_foot, int32_t* __restrict__ swing_leg, double* __restrict__ ik_err,
                  int32_t* __restrict__ ik_iters) {
    __shared__ double s_in[kThreads * 16];    // pos 3 | quat 4 | q 6 | des_vel 3
    __shared__ double s_out[kThreads * 9];    // feet 6 | next foot 3
    const int first = blockIdx.x * kThreads, nb = min(kThreads, B - first), t = threadIdx.x;
    double* s_pos = s_in; double* s_quat = s_in + 3 * kThreads; double* s_q = s_in + 7 * kThreads; double* s_dv = s_in + 13 * kThreads;
    cta_load(s_pos, pos + 3 * (size_t)first, 3 * nb);
    cta_load(s_quat, quat + 4 * (size_t)first, 4 * nb);
    cta_load(s_q, q + 6 * (size_t)first, 6 * nb);
    cta_load(s_dv, des_vel + 3 * (size_t)first, 3 * nb);
    __syncthreads();
    if (t < nb) {
        const size_t b = (size_t)first + t;
        double Rb[9], qv[3], fin[3], nxt[3], phase, remain, err;
        int ls, rs;
        quat_to_rot(s_quat + 4 * t, Rb);
        gait_state(P, iter[b], ls, rs, phase, remain);
        const int leg = (ls == 1) ? 0 : 1;                       // the swing leg (include/MPCController.h:148-152)
        double* ft = s_out + 6 * t;
        for (int l = 0; l < 2; ++l) leg_fk_world(M, l, s_pos + 3 * t, Rb, s_q + 6 * t + 3 * l, ft + 3 * l, nullptr);
        foot_placement(P, s_pos + 3 * t, s_dv + 3 * t, remain, ls, fin);
        swing_next_position(P, ft + 3 * leg, fin, remain, nxt);
        for (int k = 0; k < 3; ++k) qv[k] = s_q[6 * t + 3 * leg + k];
        const int its = leg_ik_task(M, P, leg, s_pos + 3 * t, Rb, nxt, qv, err);
        for (int k = 0; k < 3; ++k) q_cmd[6 * b + 3 * leg + k] = qv[k];   // only the swing leg's targets are written (:164-174)
        for (int k = 0; k < 3; ++k) s_out[6 * kThreads + 3 * t + k] = nxt[k];
        if (swing_leg) swing_leg[b] = leg;
        if (ik_err) ik_err[b] = err;
        if (ik_iters) ik_iters[b] = its;
    }
    __syncthreads();
    if (feet) cta_store(feet + 6 * (size_t)first, s_out, 6 * nb);
    if (next_foot) cta_store(next_foot + 3 * (size_t)first, s_out + 6 * kThreads, 3 * nb);
}

// ---- PinocchioKinematics::inverseKinematics for a batch (reference include/pinocchio_kinematics.h:61-149) ----------------
// one thread per robot: target position of contact_{L,R}_Link -> joint angles, from the initial guess q_init
__global__ void __launch_bounds__(kThreads)
leg_ik_kernel(const __grid_constant__ LegModel M, const __grid_constant__ SwingParams P, int B, const double* __restrict__ pos,
              const double* __restrict__ quat, const int32_t* __restrict__ leg, const double* __restrict__ target,
              const double* __restrict__ q_init, double* __restrict__ q_out, double* __restrict__ ik_err,
              int32_t* __restrict__ ik_iters) {
    __shared__ double s_in[kThreads * 16];    // pos 3 | quat 4 | q 6 | target 3
    __shared__ double s_out[kThreads * 6];
    const int first = blockIdx.x * kThreads, nb = min(kThreads, B - first), t = threadIdx.x;
    double* s_pos = s_in; double* s_quat = s_in + 3 * kThreads; double* s_q = s_in + 7 * kThreads; double* s_tg = s_in + 13 * kThreads;
    cta_load(s_pos, pos + 3 * (size_t)first, 3 * nb);
    cta_load(s_quat, quat + 4 * (size_t)first, 4 * nb);
    cta_load(s_q, q_init + 6 * (size_t)first, 6 * nb);
    cta_load(s_tg, target + 3 * (size_t)first, 3 * nb);
    __syncthreads();
    if (t < nb) {
        const size_t b = (size_t)first + t;
        double Rb[9], qv[3], err;
        quat_to_rot(s_quat + 4 * t, Rb);
        const int l = leg[b] ? 1 : 0;
        for (int k = 0; k < 3; ++k) qv[k] = s_q[6 * t + 3 * l + k];
        const int its = leg_ik_task(M, P, l, s_pos + 3 * t, Rb, s_tg + 3 * t, qv, err);
        for (int k = 0; k < 6; ++k) s_out[6 * t + k] = s_q[6 * t + k];      // the other leg's joints pass through
        for (int k = 0; k < 3; ++k) s_out[6 * t + 3 * l + k] = qv[k];
        if (ik_err) ik_err[b] = err;
        if (ik_iters) ik_iters[b] = its;
    }
    __syncthreads();
    cta_store(q_out + 6 * (size_t)first, s_out, 6 * nb);
}

// ---- stance-leg torques from the optimal ground-reaction forces: tau = -J' f --------------------
__global__ void __launch_bounds__(kThreads)
grf_torque_kernel(const __grid_constant__ LegModel M, int B, const double* __restrict__ quat, const double* __restrict__ q,
                  const double* __restrict__ u0, double* __restrict__ tau) {
    __shared__ double s_in[kThreads * 16];    // quat 4 | q 6 | u0 6
    __shared__ double s_out[kThreads * 6];
    const int first = blockIdx.x * kThreads, nb = min(kThreads, B - first), t = threadIdx.x;
    double* s_quat = s_in; double* s_q = s_in + 4 * kThreads; double* s_u = s_in + 10 * kThreads;
    cta_load(s_quat, quat + 4 * (size_t)first, 4 * nb);
    cta_load(s_q, q + 6 * (size_t)first, 6 * nb);
    cta_load(s_u, u0 + 6 * (size_t)first, 6 * nb);
    __syncthreads();
    if (t < nb) {
        double Rb[9], p[3], J[9];
        const double zero[3] = {0.0, 0.0, 0.0};
        quat_to_rot(s_quat + 4 * t, Rb);
        for (int leg = 0; leg < 2; ++leg) {
            leg_fk_world(M, leg, zero, Rb, s_q + 6 * t + 3 * leg, p, J);   // the Jacobian does not depend on the base position
            grf_to_torque(J, s_u + 6 * t + 3 * leg, s_out + 6 * t + 3 * leg);   // a swing foot has f = 0 -> tau = 0
        }
    }
    __syncthreads();
    cta_store(tau + 6 * (size_t)first, s_out, 6 * nb);
}

int check_launch() { return cudaGetLastError() == cudaSuccess ? MPC_B200_OK : MPC_B200_ECUDA; }

// Device scratch of the host-buffer entry points (single robot / small batches through the C++ facade): one
// lazily grown allocation and one stream per device.  Single caller per device, like an engine.
struct HostScratch {
    unsigned char* d = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
};
HostScratch g_scratch[64];

int scratch_for(int device, size_t need, HostScratch** out) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ENODEV; }
    if (device < 0 || device >= n || device >= 64) return MPC_B200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return MPC_B200_ECUDA;
    HostScratch& h = g_scratch[device];
    if (!h.s && cudaStreamCreateWithFlags(&h.s, cudaStreamNonBlocking) != cudaSuccess) return MPC_B200_ECUDA;
    if (h.bytes < need) {
        if (h.d) cudaFree(h.d);
        h.d = nullptr; h.bytes = 0;
        size_t cap = need < 4096 ? 4096 : need;
        if (cudaMalloc((void**)&h.d, cap) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ENOMEM; }
        h.bytes = cap;
    }
    *out = &h;
    return MPC_B200_OK;
}
inline size_t up16(size_t v) { return (v + 15) & ~size_t(15); }

// bump allocator over the scratch: copies `bytes` from host when src != nullptr
struct Stage {
    HostScratch* h; size_t off = 0; bool ok = true;
    template <class T> T* put(const T* src, size_t count) {
        T* d = reinterpret_cast<T*>(h->d + off);
        off = up16(off + sizeof(T) * count);
        if (src && cudaMemcpyAsync(d, src, sizeof(T) * count, cudaMemcpyHostToDevice, h->s) != cudaSuccess) ok = false;
        return d;
    }
    template <class T> void get(T* dst, const T* d, size_t count) {
        if (dst && cudaMemcpyAsync(dst, d, sizeof(T) * count, cudaMemcpyDeviceToHost, h->s) != cudaSuccess) ok = false;
    }
};

}  // namespace

extern "C" {

int mpc_b200_leg_default_model(mpc_b200_leg_model* m) {
    if (!m) return MPC_B200_EINVAL;
    // reference include/MPCParam.h:13-38; the left leg mirrors y exactly as static_foot_offset_left does (:64-66)
    const double o[5][3] = {{0.05556, 0.105, -0.2602}, {-0.077, 0.02050, 0.0}, {-0.1500, -0.02050, -0.25981},
                            {0.145, 0.0, -0.2598}, {0.0, 0.0, -0.032}};
    for (int k = 0; k < 5; ++k) {
        for (int i = 0; i < 3; ++i) { m->offset[0][k][i] = o[k][i]; m->offset[1][k][i] = o[k][i]; }
        if (k < 3) m->offset[0][k][1] = -o[k][1];   // :65 negates the abad, hip and knee y offsets only
    }
    const double ax[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 1, 0}};   // not stated by the reference (external URDF)
    for (int l = 0; l < 2; ++l)
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 3; ++i) m->axis[l][k][i] = ax[k][i];
    return MPC_B200_OK;
}

int mpc_b200_swing_default_params(mpc_b200_swing_params* p) {
    if (!p) return MPC_B200_EINVAL;
    memset(p, 0, sizeof(*p));
    p->dt = 0.001f; p->swing_time = 0.5f; p->stance_time = 0.5f; p->gait_height = 0.1f;   // include/MPCParam.h:44-51
    p->p_rel_max = 0.3;                                                                    // include/MPCController.h:111
    mpc_b200_leg_model m;
    mpc_b200_leg_default_model(&m);
    for (int i = 0; i < 3; ++i) {   // include/MPCParam.h:64-73: sums of the link offsets in the reference's order
        double l = 0.0, r = 0.0;
        for (int k = 0; k < 5; ++k) { l += m.offset[0][k][i]; r += m.offset[1][k][i]; }
        p->foot_offset_left[i] = l; p->foot_offset_right[i] = r;
    }
    p->ik_tol = 1e-3; p->ik_dt = 1e-1; p->ik_damp = 1e-6; p->ik_max_iter = 10;                // include/pinocchio_kinematics.h:61,76-77
    return MPC_B200_OK;
}

int mpc_b200_leg_fk_device(const mpc_b200_leg_model* m, int B, const double* d_base_pos, const double* d_base_quat,
                           const double* d_q, double* d_feet, double* d_jac, void* stream) {
    if (!m || B < 1 || !d_base_pos || !d_base_quat || !d_q || !d_feet) return MPC_B200_EINVAL;
    leg_fk_kernel<<<(B + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(to_model(*m), B, d_base_pos, d_base_quat, d_q, d_feet, d_jac);
    return check_launch();
}

int mpc_b200_swing_step_device(const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, int B, const double* d_base_pos,
                               const double* d_base_quat, const double* d_q, const double* d_des_vel, const int32_t* d_iter,
                               double* d_q_cmd, double* d_feet, double* d_next_foot, int32_t* d_swing_leg, double* d_ik_err,
                               int32_t* d_ik_iters, void* stream) {
    if (!m || !p || B < 1 || !d_base_pos || !d_base_quat || !d_q || !d_des_vel || !d_iter || !d_q_cmd) return MPC_B200_EINVAL;
    if (p->ik_max_iter < 0 || p->ik_mode < 0 || p->ik_mode > 1 || !(p->swing_time > 0.0f) || !(p->swing_time + p->stance_time > 0.0f)) return MPC_B200_EINVAL;
    swing_step_kernel<<<(B + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(
        to_model(*m), to_swing(*p), B, d_base_pos, d_base_quat, d_q, d_des_vel, d_iter, d_q_cmd, d_feet, d_next_foot, d_swing_leg,
        d_ik_err, d_ik_iters);
    return check_launch();
}

int mpc_b200_leg_ik_device(const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, int B, const double* d_base_pos,
                           const double* d_base_quat, const int32_t* d_leg, const double* d_target, const double* d_q_init,
                           double* d_q_out, double* d_ik_err, int32_t* d_ik_iters, void* stream) {
    if (!m || !p || B < 1 || !d_base_pos || !d_base_quat || !d_leg || !d_target || !d_q_init || !d_q_out) return MPC_B200_EINVAL;
    if (p->ik_max_iter < 0 || p->ik_mode < 0 || p->ik_mode > 1) return MPC_B200_EINVAL;
    leg_ik_kernel<<<(B + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(
        to_model(*m), to_swing(*p), B, d_base_pos, d_base_quat, d_leg, d_target, d_q_init, d_q_out, d_ik_err, d_ik_iters);
    return check_launch();
}

int mpc_b200_grf_to_torque_device(const mpc_b200_leg_model* m, int B, const double* d_base_quat, const double* d_q,
                                  const double* d_u0, double* d_tau, void* stream) {
    if (!m || B < 1 || !d_base_quat || !d_q || !d_u0 || !d_tau) return MPC_B200_EINVAL;
    grf_torque_kernel<<<(B + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(to_model(*m), B, d_base_quat, d_q, d_u0, d_tau);
    return check_launch();
}


// ---- host-buffer variants (copies in, kernel, copies out, synchronise) -------------------------------------------------
int mpc_b200_leg_fk_host(int device, const mpc_b200_leg_model* m, int B, const double* base_pos, const double* base_quat,
                         const double* q, double* feet, double* jac) {
    if (!m || B < 1 || !base_pos || !base_quat || !q || !feet) return MPC_B200_EINVAL;
    HostScratch* h;
    const size_t nb = (size_t)B;
    int rc = scratch_for(device, 8 * nb * (3 + 4 + 6 + 6 + 18) + 256, &h);
    if (rc) return rc;
    Stage st{h};
    const double* dp = st.put(base_pos, 3 * nb); const double* dq4 = st.put(base_quat, 4 * nb); const double* dq = st.put(q, 6 * nb);
    double* df = st.put((const double*)nullptr, 6 * nb); double* dj = jac ? st.put((const double*)nullptr, 18 * nb) : nullptr;
    rc = mpc_b200_leg_fk_device(m, B, dp, dq4, dq, df, dj, h->s);
    if (rc) return rc;
    st.get(feet, df, 6 * nb); st.get(jac, dj, 18 * nb);
    if (!st.ok || cudaStreamSynchronize(h->s) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ECUDA; }
    return MPC_B200_OK;
}

int mpc_b200_swing_step_host(int device, const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, int B, const double* base_pos,
                             const double* base_quat, const double* q, const double* des_vel, const int32_t* iter, double* q_cmd,
                             double* feet, double* next_foot, int32_t* swing_leg, double* ik_err, int32_t* ik_iters) {
    if (!m || !p || B < 1 || !base_pos || !base_quat || !q || !des_vel || !iter || !q_cmd) return MPC_B200_EINVAL;
    HostScratch* h;
    const size_t nb = (size_t)B;
    int rc = scratch_for(device, 8 * nb * (3 + 4 + 6 + 3 + 6 + 6 + 3 + 1) + 4 * nb * 3 + 512, &h);
    if (rc) return rc;
    Stage st{h};
    const double* dp = st.put(base_pos, 3 * nb); const double* dq4 = st.put(base_quat, 4 * nb); const double* dq = st.put(q, 6 * nb);
    const double* dv = st.put(des_vel, 3 * nb); const int32_t* di = st.put(iter, nb);
    double* dc = st.put(q_cmd, 6 * nb);       // in/out: the stance leg's entries pass through unchanged
    double* df = st.put((const double*)nullptr, 6 * nb); double* dn = st.put((const double*)nullptr, 3 * nb);
    int32_t* dl = st.put((const int32_t*)nullptr, nb); double* de = st.put((const double*)nullptr, nb);
    int32_t* dit = st.put((const int32_t*)nullptr, nb);
    rc = mpc_b200_swing_step_device(m, p, B, dp, dq4, dq, dv, di, dc, df, dn, dl, de, dit, h->s);
    if (rc) return rc;
    st.get(q_cmd, dc, 6 * nb); st.get(feet, df, 6 * nb); st.get(next_foot, dn, 3 * nb);
    st.get(swing_leg, dl, nb); st.get(ik_err, de, nb); st.get(ik_iters, dit, nb);
    if (!st.ok || cudaStreamSynchronize(h->s) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ECUDA; }
    return MPC_B200_OK;
}

int mpc_b200_leg_ik_host(int device, const mpc_b200_leg_model* m, const mpc_b200_swing_params* p, int B, const double* base_pos,
                         const double* base_quat, const int32_t* leg, const double* target, const double* q_init, double* q_out,
                         double* ik_err, int32_t* ik_iters) {
    if (!m || !p || B < 1 || !base_pos || !base_quat || !leg || !target || !q_init || !q_out) return MPC_B200_EINVAL;
    HostScratch* h;
    const size_t nb = (size_t)B;
    int rc = scratch_for(device, 8 * nb * (3 + 4 + 3 + 6 + 6 + 1) + 4 * nb * 2 + 512, &h);
    if (rc) return rc;
    Stage st{h};
    const double* dp = st.put(base_pos, 3 * nb); const double* dq4 = st.put(base_quat, 4 * nb);
    const int32_t* dl = st.put(leg, nb); const double* dt = st.put(target, 3 * nb); const double* dq = st.put(q_init, 6 * nb);
    double* dout = st.put((const double*)nullptr, 6 * nb); double* de = st.put((const double*)nullptr, nb);
    int32_t* dit = st.put((const int32_t*)nullptr, nb);
    rc = mpc_b200_leg_ik_device(m, p, B, dp, dq4, dl, dt, dq, dout, de, dit, h->s);
    if (rc) return rc;
    st.get(q_out, dout, 6 * nb); st.get(ik_err, de, nb); st.get(ik_iters, dit, nb);
    if (!st.ok || cudaStreamSynchronize(h->s) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ECUDA; }
    return MPC_B200_OK;
}

int mpc_b200_grf_to_torque_host(int device, const mpc_b200_leg_model* m, int B, const double* base_quat, const double* q,
                                const double* u0, double* tau) {
    if (!m || B < 1 || !base_quat || !q || !u0 || !tau) return MPC_B200_EINVAL;
    HostScratch* h;
    const size_t nb = (size_t)B;
    int rc = scratch_for(device, 8 * nb * (4 + 6 + 6 + 6) + 256, &h);
    if (rc) return rc;
    Stage st{h};
    const double* dq4 = st.put(base_quat, 4 * nb); const double* dq = st.put(q, 6 * nb); const double* du = st.put(u0, 6 * nb);
    double* dt = st.put((const double*)nullptr, 6 * nb);
    rc = mpc_b200_grf_to_torque_device(m, B, dq4, dq, du, dt, h->s);
    if (rc) return rc;
    st.get(tau, dt, 6 * nb);
    if (!st.ok || cudaStreamSynchronize(h->s) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ECUDA; }
    return MPC_B200_OK;
}

}  // extern "C"
