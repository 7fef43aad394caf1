// mpc_b200.cu -- sm_100a kernels and the C ABI (include/mpc_b200.h) of the batched convex-MPC engine.
//
// Kernel map:
//   tron1_solve_kernel<N,WPI,IPC>   the hot path: one thread group (WPI warps) per robot instance,
//                                   IPC instances per CTA; inputs staged into shared memory with
//                                   1-D TMA bulk copies (cp.async.bulk + mbarrier), everything else
//                                   (model, condensing, packed Cholesky, active-face iterations,
//                                   ADMM fallback) lives in shared memory / registers; only the
//                                   forces, status and iteration count go back to HBM.
//   tron1_condense_kernel<N>        parity dump of H, f, A_aug, B_aug (tests only use it)
//   contact_schedule_kernel         MPC::calculateGait over the horizon, bit-exact
//   fp64_peak_kernel                DFMA-chain microbenchmark: the FP64 roofline denominator
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mpc_b200.h"
#include "tron1_core.cuh"
#include "tron1_params.h"

using namespace mpcb200;

// horizon 50 uses the tiled storage + tensor-core Cholesky (tron1_core.cuh: chol_tiled); the shorter horizons keep the
// packed triangle and the register-resident eliminations
#ifndef MPC_LANES
#define MPC_LANES 6          // engine-owned streams of the pipelined device entry (2, 3, 4, 6 measured: profiles/r2_pipelined_lanes.log)
#endif
#ifndef MPC_DYNAMIC
// direct class: 1 = persistent grid, groups pull instances from an atomic counter.  Built, parity
// green and MEASURED (profiles/r2_dynamic_vs_static.log): 7-10 % SLOWER than the static one-CTA-per-four-instances
// mapping at every batch size (B = 4096: 50.3 us against 45.5 us; B = 65536: 139.6 against 150.7 M solves/s).  The
// hardware CTA scheduler already refills a slot as soon as a CTA retires, so what the batch time is made of is two
// rounds of the ~20 us loaded single-instance latency, which instance-level pulling does not shorten, while the 8-byte
// asynchronous copies cost more issue slots than the three TMA bulk copies per CTA of the static path.  Default: static.
#define MPC_DYNAMIC 0
#endif
#ifndef MPC_N50_WPI_S
#define MPC_N50_WPI_S 8      // warps per instance of the single-stance class of horizon 50
#endif
#ifndef MPC_N10_MINB_L
#define MPC_N10_IPC_L 2      // double-support class of horizon 10: instances per CTA (2 warps each), resident CTAs per SM
#define MPC_N10_MINB_L 3
#endif
#ifndef MPC_N50_WPI_L
// double-support class of horizon 50: warps per instance, resident CTAs per SM.  16 warps x 1 CTA: the 379 KB factor slabs of
// the resident CTAs (148 x 379 KB = 56 MB) stay inside the 126 MB L2; 8 warps x 2 CTAs (112 MB of slabs) measured between
// 409 k and 757 k cycles per factorisation depending on how much of that the L2 kept (tools/microbench/chol_dmma_bench.cu).
#define MPC_N50_WPI_L 16
#define MPC_N50_MINB_L 1
#endif
#ifndef MPC_TILED_N50
#define MPC_TILED_N50 1
#endif
#ifndef MPC_TILED_N20_S
#define MPC_TILED_N20_S 0    // single-stance class of horizon 20 (60 variables) on the tiled DMMA Cholesky instead of the register elimination
#endif
#ifndef MPC_N20_WPI_S
#define MPC_N20_WPI_S 2
#define MPC_N20_IPC_S 2
#define MPC_N20_MINB_S 2
#endif
// Riccati direct class per horizon (0 = off): active-face solves as O(N) sweeps instead of a dense factorisation
#ifndef MPC_RIC_N50
#define MPC_RIC_N50 1
#endif
#ifndef MPC_RIC_N20
#define MPC_RIC_N20 0
#endif
// horizon 20: Riccati work type as the LIST-DRIVEN class (double support: 18 against 1.9 M solves/s for the packed Cholesky), the dense
// class behind it on a second list
#ifndef MPC_RIC_L_N20
#define MPC_RIC_L_N20 1
#endif
#ifndef MPC_RIC_N10
#define MPC_RIC_N10 0
#endif
#ifndef MPC_RIC_WPI
#define MPC_RIC_WPI 1
#endif
#ifndef MPC_RIC_IPC
#define MPC_RIC_IPC 1
#endif
#ifndef MPC_RIC_MINB
#define MPC_RIC_MINB 8
#endif
// Riccati class of horizon 50: the per-step gains (300 x 13 doubles at full capacity) live in global memory, in slabs drawn
// from a ring of free slabs sized for the resident CTAs (L2-resident working set) -- 1 = external, 0 = inside shared memory
#ifndef MPC_RIC_EXT_N50
#define MPC_RIC_EXT_N50 1
#endif
#define MPC_RIC_RING_DBL 4096     // doubles reserved in front of the slabs for the ring: head, tail, size, -, entries
#ifndef MPC_PLAIN_LOAD_GRID
#define MPC_PLAIN_LOAD_GRID 64   // direct class: grids of at most this many CTAs stage their inputs with per-thread asynchronous copies instead of TMA bulk copies
#endif
#ifndef MPC_N10_WPI_LAT
// latency class of the double-support instances of horizon 10 (a standing robot, BASELINE configs[0]): batches that leave
// SMs idle anyway (B <= number of SMs) run one instance per CTA on MPC_N10_WPI_LAT warps with the tiled tensor-core
// factorisation instead of two warps on the register elimination.  Host call B = 1, standing: 37.6 -> 30.4 us p50
// (profiles/r2_latency_class.log); at throughput batch sizes the two-warp class wins (same file).
#define MPC_N10_WPI_LAT 8
#endif
// storage rule: tiled 8x8 layout (DMMA Cholesky) for horizon 50 and for groups of >= 4 warps on the 60-variable class of horizon 10
template <int N, int NC, bool AINL = true, int WPI = 1, bool RIC = false>
using SolveWork = Tron1Work<N, NC, AINL, !RIC && ((N == 50 && MPC_TILED_N50 != 0) || (N == 10 && NC == 60 && WPI >= 4) || (N == 20 && NC == 60 && MPC_TILED_N20_S != 0)), RIC>;

// ------------------------------------------------------------------------------------------------
// thread group = WPI warps cooperating on one instance
template <int WPI>
struct GrpCuda {
    static constexpr int kThreads = 32 * WPI;
    int t, gid;
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int size() const { return 32 * WPI; }
    __device__ __forceinline__ void sync() const {
        if (WPI == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(32 * WPI) : "memory");
    }
    // group-wide OR of a predicate (a barrier as well): one vote for a warp, one reducing named barrier otherwise
    __device__ __forceinline__ bool any(bool p) const {
        if (WPI == 1) return __any_sync(0xffffffffu, p);
        int r;
        asm volatile("{\n.reg .pred p, q;\nsetp.ne.s32 q, %1, 0;\nbar.red.or.pred p, %2, %3, q;\nselp.s32 %0, 1, 0, p;\n}"
                     : "=r"(r) : "r"((int)p), "r"(gid + 1), "n"(32 * WPI) : "memory");
        return r != 0;
    }
};

// ---- 1-D TMA bulk copy global -> shared, completion on an mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

#if defined(MPC_PHASE_TIMING)
__device__ unsigned long long g_cta_trace[4 * 16384];   // per CTA: start ns, end ns, SM id, inputs-arrived ns (profiling build only)
__device__ unsigned long long g_cta_trace2[4 * 16384];  // per CTA: copies issued, schedule evaluated, CTA barrier passed, (unused)
extern "C" int mpc_b200_debug_cta_trace2(unsigned long long* out, int n) {
    return cudaMemcpyFromSymbol(out, g_cta_trace2, sizeof(unsigned long long) * 4 * n) == cudaSuccess ? 0 : -3;
}
extern "C" int mpc_b200_debug_cta_trace(unsigned long long* out, int n) {
    return cudaMemcpyFromSymbol(out, g_cta_trace, sizeof(unsigned long long) * 4 * n) == cudaSuccess ? 0 : -3;
}
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ unsigned long long g_phase_cycles[16];
extern "C" int mpc_b200_debug_phase_cycles(unsigned long long* out, int reset) {
    if (out && cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return -3;
    if (reset) {
        unsigned long long z[16] = {};
        if (cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)) != cudaSuccess) return -3;
    }
    return 0;
}
#endif

template <int N, int IPC>
struct CtaStage {
    static constexpr int XR = 13 * (N + 1);
    alignas(16) double xr[IPC * XR];
    alignas(16) double x0[IPC * 13 + 1];
    alignas(16) double feet[IPC * 6 * N];   // sized for per-step feet
    alignas(8) uint64_t bar;
};

// Capacity routing: NC = 3N (one stance foot per step, the reference's alternating gait) keeps the
// per-instance workspace small enough for 12-16 resident instances per SM; instances that need more
// (double support) are appended to an overflow list and solved by the NC = 6N instantiation, which
// runs INDIRECT (list-driven, grid-stride, plain loads) right after.  No host synchronisation.
// DYNAMIC (direct class only): the grid is PERSISTENT -- at most MINB CTAs per SM -- and every thread group pulls its next
// instance from an atomic counter (ovf_count[2]; ovf_count[3] counts finished CTAs, the last one clears both), so a group
// whose instance needs three active-face iterations no longer holds three finished neighbours' slots, and the batch
// does not run in whole waves of CTAs.  A group's inputs arrive by 8-byte asynchronous copies (cp.async / LDGSTS; instance
// slices are only 8-byte aligned, so the 16-byte bulk copies of the static path cannot address a single instance), and
// the copy of the NEXT instance is issued as soon as the current one's staged inputs are dead (after setup_instance),
// i.e. it overlaps the whole elimination.
template <int N, int NC, int WPI, int IPC, int MINB, bool INDIRECT, bool AINL = true, bool DYNAMIC = false, bool RIC = false>
__global__ void __launch_bounds__(32 * WPI * IPC, MINB)
tron1_solve_kernel(const __grid_constant__ Tron1Const P, int B, const double* __restrict__ x0,
                   const double* __restrict__ xref, const double* __restrict__ feet,
                   const uint8_t* __restrict__ contact, const int32_t* __restrict__ iter,
                   double* __restrict__ forces, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                   int32_t* __restrict__ ovf_list, int32_t* __restrict__ ovf_count, double* __restrict__ ext_A,
                   const double* __restrict__ cmd_oy, const double* __restrict__ cmd_vx, int first_step_only) {
    // cmd_oy/cmd_vx != nullptr: controller-shaped call -- x_ref is generated here from the staged state and
    // the commanded yaw rate / forward velocity (include/mpcQP.h:74-97) instead of being read from HBM;
    // first_step_only: write u_0 (6 doubles per instance, include/mpcQP.h:118) instead of the whole horizon
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Stage = CtaStage<N, IPC>;
    using Work = SolveWork<N, NC, AINL, WPI, RIC>;
    Stage& st = *reinterpret_cast<Stage*>(smem_raw);
    constexpr size_t stage_bytes = (sizeof(Stage) + 15) & ~size_t(15);
    Work* works = reinterpret_cast<Work*>(smem_raw + stage_bytes);
    const int fstride = (P.per_step_feet && P.ltv) ? 6 * N : 6;
    constexpr int XR = Stage::XR;

#if defined(MPC_PHASE_TIMING)
    if (!INDIRECT && threadIdx.x == 0 && blockIdx.x < 16384) {
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_cta_trace[4 * blockIdx.x] = gtime_ns(); g_cta_trace[4 * blockIdx.x + 2] = smid; g_cta_trace[4 * blockIdx.x + 1] = 0;
    }
#endif
    GrpCuda<WPI> g;
    g.t = threadIdx.x % (32 * WPI);
    g.gid = threadIdx.x / (32 * WPI);
    Work& S = works[g.gid];
    S.x0 = st.x0 + g.gid * 13;
    S.feet = st.feet + g.gid * fstride;
    S.Aext = (Work::AINL || RIC) ? nullptr : ext_A + ((size_t)blockIdx.x * IPC + g.gid) * Work::ASZ;   // one slab per resident group
    const double* xr_s = st.xr + g.gid * XR;
    if constexpr (RIC) {
        // the staged inputs are dead once the instance is set up: the adjoint scratch of the gradient passes reuses them
        static_assert(IPC == 1 && sizeof(Stage) >= sizeof(double) * 18 * (N + 1) + 16, "adjoint scratch inside the input staging area");
        S.adjx = st.xr;
    }

    auto load_contact = [&](int b, int it0 = 0, bool have_it = false) {   // fills S.contact, returns the compact size 3 * stance foot-steps
        if (contact) {
            for (int s = g.t; s < 2 * N; s += g.size()) S.contact[s] = contact[(size_t)b * 2 * N + s] ? 1 : 0;
        } else {
            if (!have_it) it0 = iter[b];
            for (int k = g.t; k < N; k += g.size()) {
                int l, r;
                gait_contact(P, it0 < 0 ? it0 : it0 + k * P.gait_mpc_step, l, r);
                S.contact[2 * k] = (int8_t)l;
                S.contact[2 * k + 1] = (int8_t)r;
            }
        }
        g.sync();
        if constexpr (WPI == 1 && 2 * N <= 32) {   // one foot-step per lane: count by ballot
            return 3 * __popc(__ballot_sync(0xffffffffu, g.t < 2 * N && S.contact[g.t]));
        } else {
            int c = 0;
            for (int s = 0; s < 2 * N; ++s) c += S.contact[s];
            return 3 * c;
        }
    };
    auto finish = [&](int b) {
        int its = 0;
        if (cmd_oy) {   // build the reference in shared memory (the staged x0 is already visible to this group)
            make_reference(S.x0, cmd_oy[b], cmd_vx[b], P.Ts, N, const_cast<double*>(xr_s), g);
            g.sync();
        }
#if defined(MPC_PHASE_TIMING)
        if (g.t == 0) { for (int i = 0; i < 16; ++i) S.prof[i] = 0; S.t_last = clock64(); }
        g.sync();
#endif
        // Riccati work type: a further active-face iteration needs the inputs again (plain loads: rare); the staging area has
        // been the adjoint scratch of the gradient pass in between
        auto restage = [&]() {
            if constexpr (RIC) {
                g.sync();
                double* sx = st.xr + g.gid * XR; double* s0 = st.x0 + g.gid * 13; double* sf = st.feet + g.gid * fstride;
                for (int i = g.t; i < 13; i += g.size()) s0[i] = x0[(size_t)b * 13 + i];
                for (int i = g.t; i < fstride; i += g.size()) sf[i] = feet[(size_t)b * fstride + i];
                if (cmd_oy) {
                    g.sync();
                    make_reference(s0, cmd_oy[b], cmd_vx[b], P.Ts, N, sx, g);
                } else {
                    for (int i = g.t; i < XR; i += g.size()) sx[i] = xref[(size_t)b * XR + i];
                }
                g.sync();
            }
        };
        int code = solve_instance<Work>(P, S, xr_s, g, its, false, NoHook(), restage);
#if defined(MPC_PHASE_TIMING)
        MPC_TICK(S, g, 13);
        if (g.t == 0) for (int i = 0; i < 16; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)S.prof[i]);
#endif
        if (RIC && code == ST_DEFER) {
            // Riccati class: the active-face iteration did not certify -> the dense class solves the instance from scratch.
            // As the direct class it appends to the overflow list; as the list-driven class it appends to the SECOND list
            // (ext_A carries it: this instantiation keeps its gains in shared memory), counters ovf_count[2..3]
            if (g.t == 0) {
                int32_t* dl = INDIRECT ? reinterpret_cast<int32_t*>(ext_A) : ovf_list;
                int32_t* dc = INDIRECT ? ovf_count + 2 : ovf_count;
                if (dl) dl[atomicAdd(dc, 1)] = b;
                else if (status) status[b] = ST_FAILED;
            }
            return;
        }
        if (first_step_only) {
            if (g.t < 6) forces[(size_t)b * 6 + g.t] = S.up()[g.t];
        } else {
            double* out = forces + (size_t)b * 6 * N;
            for (int i = g.t; i < 6 * N; i += g.size()) out[i] = S.up()[i];
        }
        if (g.t == 0) {
            if (status) status[b] = code;
            if (iters) iters[b] = its;
        }
#if defined(MPC_PHASE_TIMING)
        if (!INDIRECT && g.t == 0 && blockIdx.x < 16384) atomicMax(&g_cta_trace[4 * blockIdx.x + 1], gtime_ns());
#endif
    };

    if constexpr (!INDIRECT && DYNAMIC) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the overflow grid may queue up behind us
        double* sx = st.xr + g.gid * XR;
        double* s0 = st.x0 + g.gid * 13;
        double* sf = st.feet + g.gid * fstride;
        int32_t* next = ovf_count + 2;
        int* s_next = reinterpret_cast<int*>(&st.bar) + (g.gid & 1);   // the static path's mbarrier slot: 2 ints (IPC <= 2 when WPI > 1)
        auto claim = [&]() -> int {
            int b;
            if constexpr (WPI == 1) {
                b = 0;
                if (g.t == 0) b = atomicAdd(next, 1);
                b = __shfl_sync(0xffffffffu, b, 0);
            } else {
                static_assert(WPI == 1 || IPC <= 2, "claim slot");
                g.sync();                                  // the previous value has been read by every thread
                if (g.t == 0) *s_next = atomicAdd(next, 1);
                g.sync();
                b = *s_next;
            }
            return b;
        };
        auto cp8 = [](double* dst, const double* src) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
        };
        auto fetch = [&](int b) {      // asynchronous: completes at the next wait_all
            if (b >= B) return;
            if (!cmd_oy) for (int i = g.t; i < XR; i += g.size()) cp8(sx + i, xref + (size_t)b * XR + i);
            for (int i = g.t; i < 13; i += g.size()) cp8(s0 + i, x0 + (size_t)b * 13 + i);
            for (int i = g.t; i < fstride; i += g.size()) cp8(sf + i, feet + (size_t)b * fstride + i);
        };
        // schedule word(s) of an instance, fetched when it is claimed and consumed when it is started
        auto sched = [&](int b) -> int {
            if (b >= B) return 0;
            if (contact) { int v = 0; for (int s = g.t, k = 0; s < 2 * N; s += g.size(), ++k) v |= (contact[(size_t)b * 2 * N + s] ? 1 : 0) << k; return v; }
            return iter[b];
        };
        auto put_contact = [&](int word) {   // fills S.contact from the schedule word, returns the compact size
            if (contact) {
                for (int s = g.t, k = 0; s < 2 * N; s += g.size(), ++k) S.contact[s] = (word >> k) & 1;
            } else {
                for (int k = g.t; k < N; k += g.size()) {
                    int l, r;
                    gait_contact(P, word < 0 ? word : word + k * P.gait_mpc_step, l, r);
                    S.contact[2 * k] = (int8_t)l;
                    S.contact[2 * k + 1] = (int8_t)r;
                }
            }
            g.sync();
            if constexpr (WPI == 1 && 2 * N <= 32) {
                return 3 * __popc(__ballot_sync(0xffffffffu, g.t < 2 * N && S.contact[g.t]));
            } else {
                int c = 0;
                for (int s = 0; s < 2 * N; ++s) c += S.contact[s];
                return 3 * c;
            }
        };
        int b = claim();
        fetch(b);
        int word = sched(b);
        while (b < B) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            g.sync();
            const int nc = put_contact(word);
            int b_next = B, word_next = 0;
            bool claimed = false;
            auto prefetch_next = [&]() {
                g.sync();                      // every thread of the group is done with the staged inputs
                b_next = claim();
                fetch(b_next);
                word_next = sched(b_next);
                claimed = true;
            };
            if (nc > NC) {     // does not fit this capacity class: hand over to the large instantiation
                if (g.t == 0) {
                    if (ovf_list) ovf_list[atomicAdd(ovf_count, 1)] = b;
                    else if (status) status[b] = ST_FAILED;
                }
                prefetch_next();
            } else {
                int its = 0;
                if (cmd_oy) {
                    make_reference(S.x0, cmd_oy[b], cmd_vx[b], P.Ts, N, sx, g);
                    g.sync();
                }
                const int code = solve_instance<Work>(P, S, sx, g, its, false, prefetch_next);
                if (first_step_only) {
                    if (g.t < 6) forces[(size_t)b * 6 + g.t] = S.up()[g.t];
                } else {
                    double* out = forces + (size_t)b * 6 * N;
                    for (int i = g.t; i < 6 * N; i += g.size()) out[i] = S.up()[i];
                }
                if (g.t == 0) {
                    if (status) status[b] = code;
                    if (iters) iters[b] = its;
                }
                if (!claimed) prefetch_next();
            }
            b = b_next;
            word = word_next;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&ovf_count[3], 1) == (int)gridDim.x - 1) { __threadfence(); ovf_count[2] = 0; ovf_count[3] = 0; }
        }
    } else if constexpr (!INDIRECT) {
        const int first = blockIdx.x * IPC;
        const int valid = min(IPC, B - first);
        // ---- stage this CTA's inputs: TMA bulk copies when the CTA's slice is 16-byte aligned ----------
        const double* gx = xref + (size_t)first * XR;
        const double* g0 = x0 + (size_t)first * 13;
        const double* gf = feet + (size_t)first * fstride;
        const uint32_t bx = cmd_oy ? 0u : (uint32_t)(valid * XR * sizeof(double));
        const uint32_t b0 = (uint32_t)(valid * 13 * sizeof(double));
        const uint32_t bf = (uint32_t)(valid * fstride * sizeof(double));
        // grids that leave SMs idle (a single robot, a handful of robots) are latency-bound on the arrival of the inputs: 8-byte
        // asynchronous copies issued by every thread arrive sooner than the bulk copies (host call, B = 1, pinned buffers: -2.4 us p50,
        // profiles/r2_latency_class.log)
        const bool bulk = (((cmd_oy ? (uintptr_t)0 : (uintptr_t)gx) | (uintptr_t)g0 | (uintptr_t)gf | bx | b0 | bf) & 15) == 0 &&
                          gridDim.x > MPC_PLAIN_LOAD_GRID;
        const bool mine = g.gid < valid;
        const int b = first + g.gid;
        if (bulk) {
            if (threadIdx.x == 0) {
                mbar_init(&st.bar, 1);
                mbar_expect_tx(&st.bar, bx + b0 + bf);
                if (bx) bulk_g2s(st.xr, gx, bx, &st.bar);
                bulk_g2s(st.x0, g0, b0, &st.bar);
                bulk_g2s(st.feet, gf, bf, &st.bar);
            }
        } else {
            auto cp8 = [](double* dst, const double* src) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
            };
            if (!cmd_oy) for (int i = threadIdx.x; i < valid * XR; i += blockDim.x) cp8(st.xr + i, gx + i);
            for (int i = threadIdx.x; i < valid * 13; i += blockDim.x) cp8(st.x0 + i, g0 + i);
            for (int i = threadIdx.x; i < valid * fstride; i += blockDim.x) cp8(st.feet + i, gf + i);
        }
#if defined(MPC_PHASE_TIMING)
        if (threadIdx.x == 0 && blockIdx.x < 16384) g_cta_trace2[4 * blockIdx.x] = gtime_ns();
#endif
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the overflow grid may queue up behind us
        // the contact schedule is evaluated while the copies are in flight
        // (the gait clock is read AFTER the copies have been issued: both kinds of copy are asynchronous, so one round trip covers
        // everything; reading it first delays the issue of the bulk copies by that round trip whenever the compiler schedules its
        // first use early -- measured: asynchronous controller-shaped host entry 97 -> 76 M solves/s)
        int nc = 0;
        if (mine) nc = load_contact(b);
#if defined(MPC_PHASE_TIMING)
        if (threadIdx.x == 0 && blockIdx.x < 16384) g_cta_trace2[4 * blockIdx.x + 1] = gtime_ns();
#endif
        if (!bulk) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();   // barrier init / asynchronous copies of every thread visible
#if defined(MPC_PHASE_TIMING)
        if (threadIdx.x == 0 && blockIdx.x < 16384) g_cta_trace2[4 * blockIdx.x + 2] = gtime_ns();
#endif
        if (bulk) mbar_wait(&st.bar, 0);
#if defined(MPC_PHASE_TIMING)
        if (threadIdx.x == 0 && blockIdx.x < 16384) g_cta_trace[4 * blockIdx.x + 3] = gtime_ns();
#endif
        if (!mine) return;
        if (nc > NC) {     // does not fit this capacity class: hand over to the large instantiation
            if (g.t == 0) {
                if (ovf_list) ovf_list[atomicAdd(ovf_count, 1)] = b;
                else if (status) status[b] = ST_FAILED;
            }
            return;
        }
        if constexpr (RIC && !AINL) {
            // gain slab for the lifetime of this instance: pop a free slab index from the ring (entries are taken with an exchange and
            // returned with a compare-and-swap, so a slot is never read before it was refilled nor refilled before it was taken)
            static_assert(IPC == 1, "one instance per CTA owns the slab");
            int32_t* ring = reinterpret_cast<int32_t*>(ext_A);
            int slab = 0;
            if (g.t == 0) {
                const unsigned h = atomicAdd(reinterpret_cast<unsigned*>(ring), 1u) % (unsigned)ring[2];
                int v;
                do { v = atomicExch(ring + 4 + h, -1); } while (v < 0);
                slab = v;
                st.bar = (uint64_t)v;             // the staging barrier is dead: broadcast slot
            }
            g.sync();
            slab = (int)st.bar;
            S.Aext = ext_A + MPC_RIC_RING_DBL + (size_t)slab * Work::ASZ;
            finish(b);
            g.sync();
            if (g.t == 0) {
                const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(ring) + 1, 1u) % (unsigned)ring[2];
                while (atomicCAS(ring + 4 + t, -1, slab) != -1) { }
            }
        } else
        finish(b);
    } else {
        // launched with programmatic stream serialisation: this grid may start while the DIRECT grid
        // drains; everything it reads is produced by that grid, so wait for its completion first
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // ovf_count[0] = list length, ovf_count[1] = CTAs that have read it; the last reader clears both
        // so the next call starts from zero without a memset (calls of one engine never overlap)
        // ovf_list == nullptr: the host has established that EVERY instance belongs to this class (e.g. a standing
        // robot): the list is the identity and no direct pass ran
        __shared__ int s_count;
        if (ovf_list) {
            if (threadIdx.x == 0) {
                // the read of the length must be performed before this CTA's ticket becomes visible: the CTA that
                // draws the last ticket clears the length, and a relaxed load could otherwise be ordered after it
                s_count = *reinterpret_cast<volatile int32_t*>(ovf_count);
                __threadfence();
                if (atomicAdd(&ovf_count[1], 1) == (int)gridDim.x - 1) {
                    __threadfence();
                    ovf_count[0] = 0; ovf_count[1] = 0;
                }
            }
            __syncthreads();
        }
        const int count = ovf_list ? s_count : B;
        for (int slot = blockIdx.x * IPC + g.gid; slot < count; slot += gridDim.x * IPC) {
            const int b = ovf_list ? ovf_list[slot] : slot;
            double* sx = st.xr + g.gid * XR;
            double* s0 = st.x0 + g.gid * 13;
            double* sf = st.feet + g.gid * fstride;
            // latency class (a single robot): the gait clock is requested together with the inputs -- one round trip instead of two
            // when they are in host memory (-1 us).  Not in the two-warp throughput class: it sits at its 168-register cap and one
            // more live value costs 80 bytes of spills and 4 % throughput (measured)
            constexpr bool kPre = (N == 10 && WPI >= 4);
            const bool pre = kPre && !contact;
            int it0 = 0;
            if constexpr (kPre) { if (pre) it0 = iter[b]; }
            if (!cmd_oy) for (int i = g.t; i < XR; i += g.size()) sx[i] = xref[(size_t)b * XR + i];
            for (int i = g.t; i < 13; i += g.size()) s0[i] = x0[(size_t)b * 13 + i];
            for (int i = g.t; i < fstride; i += g.size()) sf[i] = feet[(size_t)b * fstride + i];
            load_contact(b, it0, pre);
            finish(b);
            g.sync();
        }
    }
}

// ---- mpcQP reference generator for a batch (include/mpcQP.h:74-97) -------------------------------------
struct GrpThread {   // a single thread acting as a whole group
    static constexpr int kThreads = 1;
    __device__ __forceinline__ int tid() const { return 0; }
    __device__ __forceinline__ int size() const { return 1; }
    __device__ __forceinline__ void sync() const {}
};
__global__ void tron1_reference_kernel(const __grid_constant__ Tron1Const P, int B, int N, const double* __restrict__ x0,
                                       const double* __restrict__ oy, const double* __restrict__ vx, double* __restrict__ xr) {
    const int XR = 13 * (N + 1);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)B * XR; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / XR), r = (int)(idx % XR), i = r / 13, c = r % 13;
        const double* x = x0 + (size_t)b * 13;
        const double t = (double)i * P.Ts;
        double v = x[c];
        if (c == 2) v = x[2] + t * oy[b];
        else if (c == 3) v = x[3] + t * vx[b];
        else if (c == 9) v = (i == 0) ? x[9] : vx[b];
        else if (c == 12) v = -9.8;
        xr[idx] = v;
    }
}

// ---- closed-loop rollout (BASELINE configs[4]): state resident in shared memory for all steps -----------
template <int N, int IPC>
struct RollStage {
    static constexpr int XR = 13 * (N + 1);
    alignas(16) double xr[IPC * XR];
    alignas(16) double x[IPC * 14];
    alignas(16) double feet[IPC * 6];
};

template <int N, int NC, int WPI, int IPC, int MINB, bool STANDING>
__global__ void __launch_bounds__(32 * WPI * IPC, MINB)
tron1_rollout_kernel(const __grid_constant__ Tron1Const P, int B, int steps, double* __restrict__ x,
                     const double* __restrict__ oy, const double* __restrict__ vx, const int32_t* __restrict__ iter0,
                     double* __restrict__ u_traj, int32_t* __restrict__ uncert, int32_t* __restrict__ iters_total) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Stage = RollStage<N, IPC>;
    using Work = Tron1Work<N, NC>;
    Stage& st = *reinterpret_cast<Stage*>(smem_raw);
    constexpr size_t stage_bytes = (sizeof(Stage) + 15) & ~size_t(15);
    Work* works = reinterpret_cast<Work*>(smem_raw + stage_bytes);
    GrpCuda<WPI> g;
    g.t = threadIdx.x % (32 * WPI);
    g.gid = threadIdx.x / (32 * WPI);
    const int b = blockIdx.x * IPC + g.gid;
    if (b >= B) return;
    const int it0 = iter0[b];
    if ((it0 < 0) != STANDING) return;   // the other capacity class handles this instance
    Work& S = works[g.gid];
    S.Aext = nullptr;
    double* xs = st.x + g.gid * 14;
    double* fs = st.feet + g.gid * 6;
    double* xr = st.xr + g.gid * Stage::XR;
    S.x0 = xs;
    S.feet = fs;
    for (int i = g.t; i < 13; i += g.size()) xs[i] = x[(size_t)b * 13 + i];
    const double oyb = oy[b], vxb = vx[b];
    g.sync();
    int bad = 0, tot = 0;
    for (int s = 0; s < steps; ++s) {
        if (g.t == 0) nominal_feet(xs, P.foot_off_l, P.foot_off_r, fs);
        make_reference(xs, oyb, vxb, P.Ts, N, xr, g);
        for (int k = g.t; k < N; k += g.size()) {
            int l, r;
            gait_contact(P, it0 < 0 ? it0 : it0 + (s + k) * P.gait_mpc_step, l, r);
            S.contact[2 * k] = (int8_t)l;
            S.contact[2 * k + 1] = (int8_t)r;
        }
        g.sync();
        int its = 0;
        const int code = solve_instance<Work>(P, S, xr, g, its, s > 0);
        if (u_traj && g.t < 6) u_traj[((size_t)b * steps + s) * 6 + g.t] = S.up()[g.t];
        bad += code != 0;
        tot += its;
        integrate_state<Work>(P, S, xs, g);
    }
    for (int i = g.t; i < 13; i += g.size()) x[(size_t)b * 13 + i] = xs[i];
    if (g.t == 0) {
        if (uncert) uncert[b] = bad;
        if (iters_total) iters_total[b] = tot;
    }
}

// ---- parity dump: one warp per instance -------------------------------------------------------------
template <int N, bool AINL>
__global__ void __launch_bounds__(32)
tron1_condense_kernel(const __grid_constant__ Tron1Const P, int B, const double* __restrict__ x0,
                      const double* __restrict__ xref, const double* __restrict__ feet,
                      double* __restrict__ H, double* __restrict__ f, double* __restrict__ A_aug,
                      double* __restrict__ B_aug, double* __restrict__ ext_A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Work = Tron1Work<N, 6 * N, AINL>;
    Work& S = *reinterpret_cast<Work*>(smem_raw);
    const int b = blockIdx.x;
    if (b >= B) return;
    S.Aext = AINL ? nullptr : ext_A + (size_t)b * Work::PKN;
    GrpCuda<1> g;
    g.t = threadIdx.x;
    g.gid = 0;
    const int fstride = (P.per_step_feet && P.ltv) ? 6 * N : 6;
    const double* xr = xref + (size_t)b * 13 * (N + 1);
    S.x0 = x0 + (size_t)b * 13;
    S.feet = feet + (size_t)b * fstride;
    for (int s = g.t; s < 2 * N; s += 32) S.contact[s] = 1;
    g.sync();
    setup_instance<Work>(P, S, xr, g);
    build_hessian<Work>(P, S, 0.0, false, g);
    constexpr int n = 6 * N, p = 13 * (N + 1);
    if (H) {
        double* Hb = H + (size_t)b * n * n;
        for (int idx = g.t; idx < n * n; idx += 32) {
            int i = idx % n, j = idx / n;
            Hb[idx] = i >= j ? S.Ap()[Work::pk(i, j)] : S.Ap()[Work::pk(j, i)];
        }
    }
    if (f) for (int i = g.t; i < n; i += 32) f[(size_t)b * n + i] = S.f[i];
    if (A_aug) {
        double* Ab = A_aug + (size_t)b * p * 13;
        for (int idx = g.t; idx < p * 13; idx += 32) {
            int row = idx % p, c = idx / p;
            Ab[idx] = a_aug_entry<Work>(P, S, row / 13, row % 13, c);
        }
    }
    if (B_aug) {
        double* Bb = B_aug + (size_t)b * p * n;
        for (int idx = g.t; idx < p * n; idx += 32) {
            int row = idx % p, c = idx / p;
            Bb[idx] = b_aug_entry<Work>(P, S, row / 13, c / 6, row % 13, c % 6);
        }
    }
}

__global__ void contact_schedule_kernel(const __grid_constant__ Tron1Const P, int B, int N,
                                        const int32_t* __restrict__ iter, uint8_t* __restrict__ contact) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * N) return;
    int b = idx / N, k = idx % N;
    int it0 = iter[b], l, r;
    gait_contact(P, it0 < 0 ? it0 : it0 + k * P.gait_mpc_step, l, r);
    contact[2 * (size_t)idx] = (uint8_t)l;
    contact[2 * (size_t)idx + 1] = (uint8_t)r;
}

// ---- FP64 peak: 8 independent DFMA chains per thread, register resident -------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int reps, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------
struct mpc_b200_engine {
    int device = 0, N = 0, max_batch = 0;
    mpc_b200_tron1_params params;
    Tron1Const C;
    cudaStream_t stream = nullptr;   // used by the host-buffer entry point
    // device staging for the host-buffer entry point
    double *d_x0 = nullptr, *d_xref = nullptr, *d_feet = nullptr, *d_forces = nullptr;
    uint8_t* d_contact = nullptr;
    int32_t *d_iter = nullptr, *d_status = nullptr, *d_iters = nullptr;
    int32_t* d_ovf_list2 = nullptr;                          // second overflow list (instances the list-driven Riccati class hands to the dense class)
    int32_t *d_ovf_list = nullptr, *d_ovf_count = nullptr;   // per slot: {overflow length, readers, next instance, finished CTAs}
    double *d_oy = nullptr, *d_vx = nullptr, *d_u0 = nullptr; // controller-shaped entry: commands in, first-step force out
    static constexpr int kPipe = 8;                          // streams used by the host-buffer entry
    static constexpr int kSmallB = 64;                       // packed single-copy path below this batch size
    static constexpr int kZeroCopyB = 8;                     // below this the kernels read/write the pinned staging directly
    cudaStream_t pipe[kPipe] = {};
    unsigned char *h_small = nullptr, *d_small = nullptr;    // pinned / device staging of the packed path
    // pipelined device entry (mpc_b200_tron1_solve_device_pipelined): consecutive independent batches run on alternating
    // engine-owned streams so that the CTAs of batch k+1 fill the slots the stragglers of batch k leave idle
    static constexpr int kLanes = MPC_LANES;
    cudaStream_t lane[kLanes] = {};
    cudaEvent_t lane_in[kLanes] = {}, lane_done[kLanes] = {};
    bool lane_busy[kLanes] = {};
    unsigned lane_next = 0;
    // asynchronous host entry (mpc_b200_tron1_solve_host_async): per-lane device result buffers, allocated on first use
    double* d_lane_forces = nullptr;
    int32_t *d_lane_status = nullptr, *d_lane_iters = nullptr;
    double* d_condense_ws = nullptr;                         // horizon-50 parity dump: packed 300 x 300 workspace, grown on demand
    size_t condense_ws_bytes = 0;
    double* d_extA = nullptr;                                // global-memory factor slabs (N = 50 double support)
    int extA_slabs = 0;
    double* d_ricK = nullptr;                                // Riccati class of horizon 50: ring of free slabs + gain slabs
    cudaEvent_t extA_free = nullptr;                         // recorded behind every kernel that uses the slabs: the next user
                                                             // (possibly on another pipe stream) waits on it
    size_t small_bytes = 0;
    int host_mode = MPC_B200_HOST_AUTO;                      // how the host-buffer entry points move data
    int last_host_path = 0;                                  // 1 = zero-copy, 0 = staged (for tests / bench)
    int num_sms = 148;
    int64_t launches = 0;
    std::string err;
};

static int set_err(mpc_b200_engine* e, int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (e) {
        e->err = what;
        if (ce != cudaSuccess) { e->err += ": "; e->err += cudaGetErrorString(ce); }
    }
    return code;
}
#define CU(e, call)                                                        \
    do {                                                                   \
        cudaError_t ce_ = (call);                                          \
        if (ce_ != cudaSuccess) return set_err((e), MPC_B200_ECUDA, #call, ce_); \
    } while (0)

template <int N, int NC, int IPC>
static size_t solve_smem_bytes() {
    return ((sizeof(CtaStage<N, IPC>) + 15) & ~size_t(15)) + sizeof(Tron1Work<N, NC>) * IPC;
}

// small class (direct, TMA-staged) followed by the large class (indirect, overflow list)
// RIC_L: the list-driven class is the Riccati work type (gains in shared memory); the instances it does not certify go to a second
// list and a third, dense list-driven launch <WPI_D, IPC_D, MINB_D> (always launched: it also resets the second list's counters)
template <int N, int WPI_S, int IPC_S, int MINB_S, int WPI_L, int IPC_L, bool AINL_L = true, int MINB_L = 1, bool RIC_S = false, bool AINL_S = true,
          bool RIC_L = false, int WPI_D = 2, int IPC_D = 2, int MINB_D = 1>
static int launch_solve(mpc_b200_engine* e, int B, const double* x0, const double* xref, const double* feet,
                        const uint8_t* contact, const int32_t* iter, double* forces, int32_t* status,
                        int32_t* iters, cudaStream_t s, int32_t* ovf_list, int32_t* ovf_count,
                        const double* cmd_oy, const double* cmd_vx, int first_only, int cls_hint) {
    // cls_hint (host entry points have seen the schedule): 0 = no instance needs the large class, 2 = every instance does,
    // 1 = mixed / unknown (device entry point)
    // RIC_S: the direct class is the Riccati work type with full capacity (every contact pattern), the list-driven dense class
    // behind it takes the instances whose active-face iteration did not certify -- both kernels always run
    constexpr int NC_S = RIC_S ? 6 * N : 3 * N;
    auto ks = tron1_solve_kernel<N, NC_S, WPI_S, IPC_S, MINB_S, false, AINL_S, (MPC_DYNAMIC != 0) && !RIC_S, RIC_S>;
    auto kl = tron1_solve_kernel<N, 6 * N, WPI_L, IPC_L, MINB_L, true, AINL_L, false, RIC_L>;
    auto kd = tron1_solve_kernel<N, 6 * N, WPI_D, IPC_D, MINB_D, true, true>;
    static_assert(!RIC_L || (AINL_L && MPC_DYNAMIC == 0), "the second overflow list uses the counters of the dynamic scheduler");
    const size_t smem_d = ((sizeof(CtaStage<N, IPC_D>) + 15) & ~size_t(15)) + sizeof(SolveWork<N, 6 * N, true, WPI_D>) * IPC_D;
    int32_t* ovf2 = RIC_L ? e->d_ovf_list2 + (ovf_list - e->d_ovf_list) : nullptr;
    auto launch_dense_behind = [&]() -> int {     // third launch: the dense class over the second list
        int grid_d = (B + IPC_D - 1) / IPC_D;
        if (grid_d > e->num_sms * (MINB_D > 2 ? MINB_D : 2)) grid_d = e->num_sms * (MINB_D > 2 ? MINB_D : 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid_d); cfg.blockDim = dim3(32 * WPI_D * IPC_D); cfg.dynamicSmemBytes = smem_d; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CU(e, cudaLaunchKernelEx(&cfg, kd, e->C, B, x0, xref, feet, contact, iter, forces, status, iters, ovf2, ovf_count + 2,
                                 (double*)nullptr, cmd_oy, cmd_vx, first_only));
        CU(e, cudaGetLastError());
        e->launches += 1;
        return MPC_B200_OK;
    };
    const size_t smem_s = ((sizeof(CtaStage<N, IPC_S>) + 15) & ~size_t(15)) + sizeof(SolveWork<N, NC_S, AINL_S, WPI_S, RIC_S>) * IPC_S;
    if (RIC_S) cls_hint = 1;
    const size_t smem_l = ((sizeof(CtaStage<N, IPC_L>) + 15) & ~size_t(15)) + sizeof(SolveWork<N, 6 * N, AINL_L, WPI_L, RIC_L>) * IPC_L;
    static std::atomic<bool> configured[64];      // per device; setting the attribute twice is harmless, so a lost race is too
    if (!configured[e->device & 63].load(std::memory_order_acquire)) {
        CU(e, cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        CU(e, cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
        if (RIC_L) CU(e, cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
        configured[e->device & 63].store(true, std::memory_order_release);
    }
    // the large class of horizon 50 keeps its factors in ONE set of global slabs indexed by CTA: two such kernels must
    // never overlap, whatever streams they are on (the chunk-pipelined host path uses several)
    if (!AINL_L) CU(e, cudaStreamWaitEvent(s, e->extA_free, 0));
    if (cls_hint == 2) {   // the host entry point has looked at the schedule: every instance is double support
        int grid = (B + IPC_L - 1) / IPC_L;
        if (grid > e->num_sms * (MINB_L > 2 ? MINB_L : 2)) grid = e->num_sms * (MINB_L > 2 ? MINB_L : 2);
        if (!AINL_L && grid > e->num_sms * MINB_L) grid = e->num_sms * MINB_L;     // slab class: only the resident CTAs (L2 working set)
        if (!AINL_L && grid * IPC_L > e->extA_slabs) grid = e->extA_slabs / IPC_L;
        kl<<<grid, 32 * WPI_L * IPC_L, smem_l, s>>>(e->C, B, x0, xref, feet, contact, iter, forces, status, iters, nullptr, ovf_count,
                                                  RIC_L ? reinterpret_cast<double*>(ovf2) : (AINL_L ? nullptr : e->d_extA), cmd_oy, cmd_vx, first_only);
        CU(e, cudaGetLastError());
        if (!AINL_L) CU(e, cudaEventRecord(e->extA_free, s));
        e->launches += 1;
        if (RIC_L) return launch_dense_behind();
        return MPC_B200_OK;
    }
    int grid_s = (B + IPC_S - 1) / IPC_S;
    if (MPC_DYNAMIC != 0 && grid_s > e->num_sms * MINB_S) grid_s = e->num_sms * MINB_S;   // persistent: the groups pull instances
    ks<<<grid_s, 32 * WPI_S * IPC_S, smem_s, s>>>(e->C, B, x0, xref, feet, contact, iter, forces, status,
                                                                 iters, ovf_list, ovf_count, (RIC_S && !AINL_S) ? e->d_ricK : nullptr, cmd_oy, cmd_vx, first_only);
    CU(e, cudaGetLastError());
    if (cls_hint == 0) {   // the host entry point has looked at the schedule: no instance can overflow, the list stays empty
        e->launches += 1;
        return MPC_B200_OK;
    }
    int grid_l = (B + IPC_L - 1) / IPC_L;
    if (grid_l > e->num_sms * (MINB_L > 2 ? MINB_L : 2)) grid_l = e->num_sms * (MINB_L > 2 ? MINB_L : 2);
    if (!AINL_L && grid_l > e->num_sms * MINB_L) grid_l = e->num_sms * MINB_L;
    if (!AINL_L && grid_l * IPC_L > e->extA_slabs) grid_l = e->extA_slabs / IPC_L;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid_l); cfg.blockDim = dim3(32 * WPI_L * IPC_L); cfg.dynamicSmemBytes = smem_l; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        double* ext = RIC_L ? reinterpret_cast<double*>(ovf2) : (AINL_L ? nullptr : e->d_extA);
        CU(e, cudaLaunchKernelEx(&cfg, kl, e->C, B, x0, xref, feet, contact, iter, forces, status, iters, ovf_list, ovf_count, ext,
                                 cmd_oy, cmd_vx, first_only));
    }
    CU(e, cudaGetLastError());
    if (!AINL_L) CU(e, cudaEventRecord(e->extA_free, s));
    e->launches += 2;
    if (RIC_L) return launch_dense_behind();
    return MPC_B200_OK;
}

// compiled configurations per horizon
static int dispatch_solve(mpc_b200_engine* e, int B, const double* x0, const double* xref, const double* feet,
                          const uint8_t* contact, const int32_t* iter, double* forces, int32_t* status,
                          int32_t* iters, cudaStream_t s, int slot = 0, int list_offset = 0,
                          const double* cmd_oy = nullptr, const double* cmd_vx = nullptr, int first_only = 0, int cls_hint = 1) {
    if (B + (list_offset % e->max_batch) > e->max_batch || B > e->max_batch) return set_err(e, MPC_B200_ECAPACITY, "solve: B > max_batch");
    int32_t* ol = e->d_ovf_list + list_offset;
    int32_t* oc = e->d_ovf_count + 4 * slot;
    switch (e->N) {
        case 10:
#if MPC_RIC_N10
            return launch_solve<10, MPC_RIC_WPI, MPC_RIC_IPC, MPC_RIC_MINB, 2, MPC_N10_IPC_L, true, MPC_N10_MINB_L, true>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
#endif
            if (B <= e->num_sms && cls_hint != 0)     // latency class for the double-support instances (see MPC_N10_WPI_LAT)
                return launch_solve<10, 1, 4, 4, MPC_N10_WPI_LAT, 1, true, 1>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
            return launch_solve<10, 1, 4, 4, 2, MPC_N10_IPC_L, true, MPC_N10_MINB_L>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
        case 20:
#if MPC_RIC_N20
            return launch_solve<20, MPC_RIC_WPI, MPC_RIC_IPC, MPC_RIC_MINB, 2, 2, true, 1, true>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
#endif
#if MPC_RIC_L_N20
            return launch_solve<20, MPC_N20_WPI_S, MPC_N20_IPC_S, MPC_N20_MINB_S, 1, 1, true, 8, false, true, true, 2, 2, 1>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
#endif
            return launch_solve<20, MPC_N20_WPI_S, MPC_N20_IPC_S, MPC_N20_MINB_S, 2, 2>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
        case 50:
#if MPC_RIC_N50
            return launch_solve<50, MPC_RIC_WPI, MPC_RIC_IPC, MPC_RIC_MINB, MPC_N50_WPI_L, 1, false, MPC_N50_MINB_L, true, MPC_RIC_EXT_N50 == 0>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
#endif
            return launch_solve<50, MPC_N50_WPI_S, 1, 1, MPC_N50_WPI_L, 1, false, MPC_N50_MINB_L>(e, B, x0, xref, feet, contact, iter, forces, status, iters, s, ol, oc, cmd_oy, cmd_vx, first_only, cls_hint);
        default: return set_err(e, MPC_B200_EINVAL, "unsupported horizon");
    }
}

template <int N, int NC, int WPI, int IPC, int MINB, bool STANDING>
static int launch_rollout_one(mpc_b200_engine* e, int B, int steps, double* x, const double* oy, const double* vx,
                              const int32_t* iter0, double* u_traj, int32_t* uncert, int32_t* iters, cudaStream_t s) {
    auto k = tron1_rollout_kernel<N, NC, WPI, IPC, MINB, STANDING>;
    const size_t smem = ((sizeof(RollStage<N, IPC>) + 15) & ~size_t(15)) + sizeof(Tron1Work<N, NC>) * IPC;
    CU(e, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(B + IPC - 1) / IPC, 32 * WPI * IPC, smem, s>>>(e->C, B, steps, x, oy, vx, iter0, u_traj, uncert, iters);
    CU(e, cudaGetLastError());
    e->launches++;
    return MPC_B200_OK;
}

extern "C" {

int mpc_b200_version(void) { return 100; }

const char* mpc_b200_strerror(int code) {
    switch (code) {
        case MPC_B200_OK: return "ok";
        case MPC_B200_EINVAL: return "invalid argument";
        case MPC_B200_ENODEV: return "no usable CUDA device (sm_100a required; there is no CPU fallback)";
        case MPC_B200_ECUDA: return "CUDA runtime error";
        case MPC_B200_ENOMEM: return "out of memory";
        case MPC_B200_ECAPACITY: return "batch exceeds engine capacity";
        default: return "unknown error";
    }
}

int mpc_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mpc_b200_tron1_default_params(mpc_b200_tron1_params* p) {
    if (!p) return MPC_B200_EINVAL;
    tron1_default_params(*p);
    return MPC_B200_OK;
}

int mpc_b200_create(const mpc_b200_tron1_params* p, int horizon, int max_batch, int device, mpc_b200_engine** out) {
    if (!p || !out || max_batch < 1) return MPC_B200_EINVAL;
    if (horizon != 10 && horizon != 20 && horizon != 50) return MPC_B200_EINVAL;
    *out = nullptr;
    int ndev = mpc_b200_device_count();
    if (device < 0 || device >= ndev) return MPC_B200_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return MPC_B200_ENODEV;
    Tron1Const C;
    if (make_tron1_const(*p, C)) return MPC_B200_EINVAL;
    if (C.per_step_feet && !C.ltv) return MPC_B200_EINVAL;
    mpc_b200_engine* e = new (std::nothrow) mpc_b200_engine();
    if (!e) return MPC_B200_ENOMEM;
    e->device = device; e->N = horizon; e->max_batch = max_batch; e->params = *p; e->C = C;
    const int N = horizon;
    const size_t fstride = (C.per_step_feet && C.ltv) ? 6 * (size_t)N : 6;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&e->d_x0, sizeof(double) * 13 * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_xref, sizeof(double) * 13 * (N + 1) * (size_t)max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_feet, sizeof(double) * fstride * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_forces, sizeof(double) * 6 * N * (size_t)max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_contact, (size_t)2 * N * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_iter, sizeof(int32_t) * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_status, sizeof(int32_t) * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_iters, sizeof(int32_t) * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_ovf_list, sizeof(int32_t) * (size_t)max_batch * (1 + mpc_b200_engine::kLanes)) == cudaSuccess &&
              cudaMalloc(&e->d_ovf_list2, sizeof(int32_t) * (size_t)max_batch * (1 + mpc_b200_engine::kLanes)) == cudaSuccess &&
              cudaMalloc(&e->d_ovf_count, 4 * (mpc_b200_engine::kPipe + 1 + mpc_b200_engine::kLanes) * sizeof(int32_t)) == cudaSuccess &&
              cudaMemset(e->d_ovf_count, 0, 4 * (mpc_b200_engine::kPipe + 1 + mpc_b200_engine::kLanes) * sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&e->d_oy, sizeof(double) * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_vx, sizeof(double) * max_batch) == cudaSuccess &&
              cudaMalloc(&e->d_u0, sizeof(double) * 6 * max_batch) == cudaSuccess;
    for (int i = 0; ok && i < mpc_b200_engine::kPipe; ++i)
        ok = cudaStreamCreateWithFlags(&e->pipe[i], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < mpc_b200_engine::kLanes; ++i)
        ok = cudaStreamCreateWithFlags(&e->lane[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->lane_in[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->lane_done[i], cudaEventDisableTiming) == cudaSuccess;
    // packed staging of the small-batch path: inputs and outputs of kSmallB instances, 16-byte aligned segments
    e->small_bytes = (size_t)mpc_b200_engine::kSmallB * (sizeof(double) * (13 + 13 * (N + 1) + fstride + 2 + 6 * N) + 2 * N + 16) + 1024;
    ok = ok && cudaHostAlloc((void**)&e->h_small, e->small_bytes, cudaHostAllocMapped) == cudaSuccess &&
         cudaMalloc((void**)&e->d_small, e->small_bytes) == cudaSuccess;
    e->num_sms = prop.multiProcessorCount;
    if (ok && horizon == 50) {   // double-support factor (301*302/2 doubles = 364 KB) does not fit shared memory
        e->extA_slabs = e->num_sms * 2;
        if (e->extA_slabs > max_batch) e->extA_slabs = max_batch;
        ok = cudaMalloc(&e->d_extA, sizeof(double) * SolveWork<50, 300, false>::ASZ * (size_t)e->extA_slabs) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->extA_free, cudaEventDisableTiming) == cudaSuccess;
    }
#if MPC_RIC_N50 && MPC_RIC_EXT_N50
    if (ok && horizon == 50) {
        const int pool = e->num_sms * 8;            // >= resident CTAs of the Riccati class (fewer would only make CTAs wait for a slab)
        const size_t slab = SolveWork<50, 300, false, 1, true>::ASZ;
        std::vector<int32_t> ring(4 + pool);
        ring[0] = 0; ring[1] = 0; ring[2] = pool; ring[3] = 0;
        for (int i = 0; i < pool; ++i) ring[4 + i] = i;
        static_assert(sizeof(int32_t) * (4 + 148 * 8 * 2) <= sizeof(double) * MPC_RIC_RING_DBL, "ring area");
        ok = (size_t)(4 + pool) * sizeof(int32_t) <= sizeof(double) * MPC_RIC_RING_DBL &&
             cudaMalloc(&e->d_ricK, sizeof(double) * (MPC_RIC_RING_DBL + slab * (size_t)pool)) == cudaSuccess &&
             cudaMemcpy(e->d_ricK, ring.data(), ring.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
    }
#endif
    if (!ok) {
        cudaGetLastError();
        mpc_b200_destroy(e);
        return MPC_B200_ENOMEM;
    }
    *out = e;
    return MPC_B200_OK;
}

int mpc_b200_destroy(mpc_b200_engine* e) {
    if (!e) return MPC_B200_EINVAL;
    cudaSetDevice(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    cudaFree(e->d_x0); cudaFree(e->d_xref); cudaFree(e->d_feet); cudaFree(e->d_forces);
    cudaFree(e->d_contact); cudaFree(e->d_iter); cudaFree(e->d_status); cudaFree(e->d_iters);
    cudaFree(e->d_ovf_list); cudaFree(e->d_ovf_list2); cudaFree(e->d_ovf_count);
    cudaFree(e->d_oy); cudaFree(e->d_vx); cudaFree(e->d_u0);
    for (int i = 0; i < mpc_b200_engine::kPipe; ++i) if (e->pipe[i]) { cudaStreamSynchronize(e->pipe[i]); cudaStreamDestroy(e->pipe[i]); }
    if (e->h_small) cudaFreeHost(e->h_small);
    cudaFree(e->d_small);
    cudaFree(e->d_extA);
    cudaFree(e->d_ricK);
    cudaFree(e->d_condense_ws);
    cudaFree(e->d_lane_forces); cudaFree(e->d_lane_status); cudaFree(e->d_lane_iters);
    for (int i = 0; i < mpc_b200_engine::kLanes; ++i) {
        if (e->lane[i]) { cudaStreamSynchronize(e->lane[i]); cudaStreamDestroy(e->lane[i]); }
        if (e->lane_in[i]) cudaEventDestroy(e->lane_in[i]);
        if (e->lane_done[i]) cudaEventDestroy(e->lane_done[i]);
    }
    if (e->extA_free) cudaEventDestroy(e->extA_free);
    delete e;
    return MPC_B200_OK;
}

const char* mpc_b200_last_error(const mpc_b200_engine* e) { return e ? e->err.c_str() : ""; }
int mpc_b200_set_host_mode(mpc_b200_engine* e, int mode) {
    if (!e || mode < MPC_B200_HOST_AUTO || mode > MPC_B200_HOST_ZEROCOPY) return set_err(e, MPC_B200_EINVAL, "set_host_mode: bad mode");
    e->host_mode = mode;
    return MPC_B200_OK;
}
int mpc_b200_last_host_path(const mpc_b200_engine* e) { return e ? e->last_host_path : 0; }

int mpc_b200_pin_host_buffer(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return MPC_B200_EINVAL;
    cudaError_t ce = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (ce == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return MPC_B200_OK; }
    if (ce != cudaSuccess) { cudaGetLastError(); return ce == cudaErrorMemoryAllocation ? MPC_B200_ENOMEM : MPC_B200_ECUDA; }
    return MPC_B200_OK;
}
int mpc_b200_unpin_host_buffer(void* ptr) {
    if (!ptr) return MPC_B200_EINVAL;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return MPC_B200_ECUDA; }
    return MPC_B200_OK;
}
int64_t mpc_b200_launch_count(const mpc_b200_engine* e) { return e ? e->launches : 0; }

int mpc_b200_contact_schedule_device(mpc_b200_engine* e, int B, const int32_t* d_iter, uint8_t* d_contact, void* stream) {
    if (!e || !d_iter || !d_contact || B < 1) return set_err(e, MPC_B200_EINVAL, "contact_schedule: bad argument");
    CU(e, cudaSetDevice(e->device));
    const int total = B * e->N;
    contact_schedule_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->C, B, e->N, d_iter, d_contact);
    CU(e, cudaGetLastError());
    e->launches++;
    return MPC_B200_OK;
}

int mpc_b200_tron1_solve_device(mpc_b200_engine* e, int B, const double* d_x0, const double* d_x_ref,
                                const double* d_feet, const uint8_t* d_contact, const int32_t* d_iter,
                                double* d_forces, int32_t* d_status, int32_t* d_iters, void* stream) {
    if (!e || !d_x0 || !d_x_ref || !d_feet || !d_forces || B < 1) return set_err(e, MPC_B200_EINVAL, "solve: bad argument");
    if ((d_contact == nullptr) == (d_iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "solve: pass exactly one of contact / iter");
    if (((uintptr_t)d_x0 | (uintptr_t)d_x_ref | (uintptr_t)d_feet | (uintptr_t)d_forces) & 7)
        return set_err(e, MPC_B200_EINVAL, "solve: pointers must be 8-byte aligned");
    CU(e, cudaSetDevice(e->device));
    return dispatch_solve(e, B, d_x0, d_x_ref, d_feet, d_contact, d_iter, d_forces, d_status, d_iters, (cudaStream_t)stream);
}

int mpc_b200_tron1_solve_device_pipelined(mpc_b200_engine* e, int B, const double* d_x0, const double* d_x_ref,
                                          const double* d_feet, const uint8_t* d_contact, const int32_t* d_iter,
                                          double* d_forces, int32_t* d_status, int32_t* d_iters, void* stream) {
    if (!e || !d_x0 || !d_x_ref || !d_feet || !d_forces || B < 1) return set_err(e, MPC_B200_EINVAL, "solve_pipelined: bad argument");
    if ((d_contact == nullptr) == (d_iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "solve_pipelined: pass exactly one of contact / iter");
    if (((uintptr_t)d_x0 | (uintptr_t)d_x_ref | (uintptr_t)d_feet | (uintptr_t)d_forces) & 7)
        return set_err(e, MPC_B200_EINVAL, "solve_pipelined: pointers must be 8-byte aligned");
    CU(e, cudaSetDevice(e->device));
    const int l = (int)(e->lane_next++ % mpc_b200_engine::kLanes);
    // ordered after everything already queued on the caller's stream (its inputs are ready) and, by stream order of the
    // lane, after the batch that used this lane before; NOT ordered against the other lanes: that is the overlap
    CU(e, cudaEventRecord(e->lane_in[l], (cudaStream_t)stream));
    CU(e, cudaStreamWaitEvent(e->lane[l], e->lane_in[l], 0));
    const int rc = dispatch_solve(e, B, d_x0, d_x_ref, d_feet, d_contact, d_iter, d_forces, d_status, d_iters, e->lane[l],
                                  mpc_b200_engine::kPipe + 1 + l, (1 + l) * e->max_batch);
    if (rc) return rc;
    e->lane_busy[l] = true;
    return MPC_B200_OK;
}

int mpc_b200_join(mpc_b200_engine* e, void* stream) {
    if (!e) return MPC_B200_EINVAL;
    CU(e, cudaSetDevice(e->device));
    for (int l = 0; l < mpc_b200_engine::kLanes; ++l) {
        if (!e->lane_busy[l]) continue;
        CU(e, cudaEventRecord(e->lane_done[l], e->lane[l]));
        CU(e, cudaStreamWaitEvent((cudaStream_t)stream, e->lane_done[l], 0));
        e->lane_busy[l] = false;
    }
    return MPC_B200_OK;
}

namespace {
// packed layout of the small-batch path: every segment starts 16-byte aligned (TMA bulk copies)
struct SmallLayout {
    size_t x0, xref, feet, oy, vx, sched, in_bytes, forces, u0, status, iters, total;
};
inline size_t up16(size_t v) { return (v + 15) & ~size_t(15); }
SmallLayout small_layout(int B, int N, size_t fstride, bool has_contact, bool cmd) {
    SmallLayout L;
    size_t o = 0;
    L.x0 = o; o = up16(o + sizeof(double) * 13 * B);
    L.xref = o; if (!cmd) o = up16(o + sizeof(double) * 13 * (N + 1) * B);
    L.feet = o; o = up16(o + sizeof(double) * fstride * B);
    L.oy = o; if (cmd) o = up16(o + sizeof(double) * B);
    L.vx = o; if (cmd) o = up16(o + sizeof(double) * B);
    L.sched = o; o = up16(o + (has_contact ? (size_t)2 * N * B : sizeof(int32_t) * B));
    L.in_bytes = o;
    L.forces = o; if (!cmd) o = up16(o + sizeof(double) * 6 * N * B);
    L.u0 = o; if (cmd) o = up16(o + sizeof(double) * 6 * B);
    L.status = o; o = up16(o + sizeof(int32_t) * B);
    L.iters = o; o = up16(o + sizeof(int32_t) * B);
    L.total = o;
    return L;
}
}  // namespace

// Device view of a host pointer the GPU can address directly (cudaHostAlloc / cudaHostRegister / managed
// memory under UVA); nullptr for pageable memory.
static void* device_view(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost && a.type != cudaMemoryTypeManaged) return nullptr;
    return a.devicePointer;
}

// Host entry points see the schedule in host memory: when it is cheap to prove that no instance needs the large
// capacity class (double support), the list-driven second kernel is not launched at all (2 us per call).
// A gait clock >= 0 alternates single support (gait_contact); a negative one means standing on both feet.  When
// EVERY instance is double support (a standing robot) only the large-class kernel runs, over the identity list.
// returns 0 = no instance needs it, 2 = every instance needs it, 1 = mixed or not worth scanning
static int schedule_large_class(const mpc_b200_engine* e, int B, const uint8_t* contact, const int32_t* iter) {
    const int N = e->N;
    int n_large = 0;
    if (iter) {
        if (B > 8192) return 1;                          // not worth scanning: keep the general two-kernel launch
        for (int b = 0; b < B; ++b) n_large += iter[b] < 0;
    } else {
        if ((size_t)B * 2 * N > 32768) return 1;
        for (int b = 0; b < B; ++b) {
            int c = 0;
            for (int s = 0; s < 2 * N; ++s) c += contact[(size_t)b * 2 * N + s] ? 1 : 0;
            n_large += c > N;
        }
    }
    return n_large == 0 ? 0 : (n_large == B ? 2 : 1);
}

// shared implementation of the two host-buffer entry points.
//   cmd == false: x_ref in, full horizon forces out (forces_out [B][N][6])
//   cmd == true : (omega_yaw, velocity_x) in -> reference generated on the device (include/mpcQP.h:74-97),
//                 first-step forces out (forces_out [B][6]), i.e. u = U_opt.col(0) (include/mpcQP.h:118)
static int solve_host_impl(mpc_b200_engine* e, int B, const double* x0, const double* x_ref, const double* oy,
                           const double* vx, const double* feet, const uint8_t* contact, const int32_t* iter,
                           double* forces_out, int32_t* status, int32_t* iters, bool cmd) {
    const int N = e->N;
    const size_t fstride = (e->C.per_step_feet && e->C.ltv) ? 6 * (size_t)N : 6;
    const size_t XR = 13 * (size_t)(N + 1);
    e->last_host_path = 0;
    const int cls = schedule_large_class(e, B, contact, iter);   // 0 / 2: one of the two kernels of every dispatch is skipped
    if (e->host_mode != MPC_B200_HOST_STAGED) {
        // ---- zero-copy path: when every caller buffer is pinned (device-addressable), the solve kernel
        // reads its inputs straight from host memory (each CTA's slice arrives by TMA bulk copies over
        // PCIe, exactly once) and writes forces / status straight back: the H2D and D2H transfers happen
        // inside the kernel, CTA by CTA, overlapped with the solves of the other resident CTAs; no
        // cudaMemcpy calls, no staging, one stream synchronise.
        const void* a0 = device_view(x0);
        const void* a1 = cmd ? device_view(oy) : device_view(x_ref);
        const void* a2 = cmd ? device_view(vx) : a1;
        const void* a3 = device_view(feet);
        const void* a4 = contact ? device_view(contact) : device_view(iter);
        void* o0 = device_view(forces_out);
        void* o1 = status ? device_view(status) : nullptr;
        void* o2 = iters ? device_view(iters) : nullptr;
        if (a0 && a1 && a2 && a3 && a4 && o0 && (!status || o1) && (!iters || o2)) {
            // Results are written by the kernel too.  SM-issued stores to host memory are slow on these hosts (12 GB/s alone,
            // 5 GB/s per GPU with eight GPUs busy, against 51 GB/s for SM-issued reads: tools/microbench/host_read_probe.cu,
            // profiles/r2_host_bw_probe_8gpu.log), but inside ONE kernel they overlap the reads (PCIe is full duplex) and the
            // solves.  Bringing the results home with the copy engine instead was measured three ways at B = 4096 and lost
            // every time against the 26.3 M solves/s of this path: chunk kernels on concurrent streams 19.4 M, chunk kernels back
            // to back with event-ordered copies 16.7 M (chunking breaks the overlap of later CTAs' reads with earlier CTAs'
            // solves), one kernel + one copy behind it 21.5 M (the copy is serial time).
            cudaStream_t s = e->stream;
            int rc = dispatch_solve(e, B, (const double*)a0, cmd ? nullptr : (const double*)a1, (const double*)a3,
                                    contact ? (const uint8_t*)a4 : nullptr, contact ? nullptr : (const int32_t*)a4,
                                    (double*)o0, (int32_t*)o1, (int32_t*)o2, s, 0, 0,
                                    cmd ? (const double*)a1 : nullptr, cmd ? (const double*)a2 : nullptr, cmd ? 1 : 0, cls);
            if (rc) return rc;
            CU(e, cudaStreamSynchronize(s));
            e->last_host_path = 1;
            return MPC_B200_OK;
        }
        if (e->host_mode == MPC_B200_HOST_ZEROCOPY)
            return set_err(e, MPC_B200_EINVAL, "host buffers are not device-addressable (pin them, or use MPC_B200_HOST_AUTO/STAGED)");
    }
    if (B <= mpc_b200_engine::kSmallB) {
        // ---- latency path: pack on the host, ONE H2D, kernels, ONE D2H ------------------------------------
        const SmallLayout L = small_layout(B, N, fstride, contact != nullptr, cmd);
        unsigned char* h = e->h_small;
        // tiny batches: zero-copy -- the kernels read the inputs from, and write the results to, the mapped
        // pinned staging buffer over PCIe (UVA: same pointer on both sides), which removes two memcpy launches
        // from the single-solve latency path
        const bool zc = B <= mpc_b200_engine::kZeroCopyB;
        unsigned char* d = zc ? h : e->d_small;
        cudaStream_t s = e->stream;
        memcpy(h + L.x0, x0, sizeof(double) * 13 * B);
        if (!cmd) memcpy(h + L.xref, x_ref, sizeof(double) * XR * B);
        memcpy(h + L.feet, feet, sizeof(double) * fstride * B);
        if (cmd) { memcpy(h + L.oy, oy, sizeof(double) * B); memcpy(h + L.vx, vx, sizeof(double) * B); }
        if (contact) memcpy(h + L.sched, contact, (size_t)2 * N * B); else memcpy(h + L.sched, iter, sizeof(int32_t) * B);
        if (!zc) CU(e, cudaMemcpyAsync(d, h, L.in_bytes, cudaMemcpyHostToDevice, s));
        // controller-shaped call: the solve kernel builds x_ref itself and writes u_0 only (no extra launches)
        int rc = dispatch_solve(e, B, (const double*)(d + L.x0), cmd ? nullptr : (const double*)(d + L.xref), (const double*)(d + L.feet),
                                contact ? (const uint8_t*)(d + L.sched) : nullptr, contact ? nullptr : (const int32_t*)(d + L.sched),
                                cmd ? (double*)(d + L.u0) : (double*)(d + L.forces), (int32_t*)(d + L.status), (int32_t*)(d + L.iters), s,
                                0, 0, cmd ? (const double*)(d + L.oy) : nullptr, cmd ? (const double*)(d + L.vx) : nullptr, cmd ? 1 : 0, cls);
        if (rc) return rc;
        if (!zc) CU(e, cudaMemcpyAsync(h + L.in_bytes, d + L.in_bytes, L.total - L.in_bytes, cudaMemcpyDeviceToHost, s));
        CU(e, cudaStreamSynchronize(s));
        if (cmd) memcpy(forces_out, h + L.u0, sizeof(double) * 6 * B);
        else memcpy(forces_out, h + L.forces, sizeof(double) * 6 * N * B);
        if (status) memcpy(status, h + L.status, sizeof(int32_t) * B);
        if (iters) memcpy(iters, h + L.iters, sizeof(int32_t) * B);
        return MPC_B200_OK;
    }
    // ---- throughput path: chunks pipelined over kPipe streams (H2D / solve / D2H of different chunks overlap) --
    // chunk so that each H2D is ~1.25 MB: on PCIe Gen5 hosts mid-size pinned copies (4-20 MB) were measured
    // at half the rate of 1 MB copies (tools/pcie_probe.py), and copies of later chunks overlap the solve of
    // earlier ones; never more chunks than streams, never chunks smaller than 512 instances
    const size_t bytes = (size_t)B * (cmd ? 172 + 56 : sizeof(double) * (13 + XR + fstride + 6 * (size_t)N));
    int nchunk = (int)((bytes + (size_t)1250000 - 1) / (size_t)1250000);
    if (cmd) nchunk = B >= 16384 ? 4 : (B >= 2048 ? 2 : 1);   // tiny transfers: split only to overlap copies with the solve
    if (nchunk > mpc_b200_engine::kPipe) nchunk = mpc_b200_engine::kPipe;
    if (nchunk > B / 512) nchunk = B / 512;
    if (nchunk < 1) nchunk = 1;
    int chunk = (B + nchunk - 1) / nchunk;
    chunk = (chunk + 3) & ~3;   // chunk starts stay multiples of 4: CTA slices remain 16-byte aligned
    // an error half way must not leave copies of earlier chunks in flight into the caller's buffers
    struct Drain {
        mpc_b200_engine* e; bool armed;
        ~Drain() { if (armed) for (int i = 0; i < mpc_b200_engine::kPipe; ++i) cudaStreamSynchronize(e->pipe[i]); }
    } drain{e, true};
    for (int c = 0, first = 0; first < B; ++c, first += chunk) {
        const int nb = (B - first < chunk) ? (B - first) : chunk;
        cudaStream_t s = e->pipe[c % mpc_b200_engine::kPipe];
        const size_t f = first;
        CU(e, cudaMemcpyAsync(e->d_x0 + 13 * f, x0 + 13 * f, sizeof(double) * 13 * nb, cudaMemcpyHostToDevice, s));
        if (!cmd) CU(e, cudaMemcpyAsync(e->d_xref + XR * f, x_ref + XR * f, sizeof(double) * XR * nb, cudaMemcpyHostToDevice, s));
        CU(e, cudaMemcpyAsync(e->d_feet + fstride * f, feet + fstride * f, sizeof(double) * fstride * nb, cudaMemcpyHostToDevice, s));
        if (contact) CU(e, cudaMemcpyAsync(e->d_contact + 2 * N * f, contact + 2 * N * f, (size_t)2 * N * nb, cudaMemcpyHostToDevice, s));
        else CU(e, cudaMemcpyAsync(e->d_iter + f, iter + f, sizeof(int32_t) * nb, cudaMemcpyHostToDevice, s));
        if (cmd) {
            CU(e, cudaMemcpyAsync(e->d_oy + f, oy + f, sizeof(double) * nb, cudaMemcpyHostToDevice, s));
            CU(e, cudaMemcpyAsync(e->d_vx + f, vx + f, sizeof(double) * nb, cudaMemcpyHostToDevice, s));
        }
        int rc = dispatch_solve(e, nb, e->d_x0 + 13 * f, cmd ? nullptr : e->d_xref + XR * f, e->d_feet + fstride * f,
                                contact ? e->d_contact + 2 * N * f : nullptr, contact ? nullptr : e->d_iter + f,
                                cmd ? e->d_u0 + 6 * f : e->d_forces + 6 * N * f, e->d_status + f, e->d_iters + f, s,
                                1 + c % mpc_b200_engine::kPipe, first, cmd ? e->d_oy + f : nullptr, cmd ? e->d_vx + f : nullptr, cmd ? 1 : 0, cls);
        if (rc) return rc;
        if (cmd) {
            CU(e, cudaMemcpyAsync(forces_out + 6 * f, e->d_u0 + 6 * f, sizeof(double) * 6 * nb, cudaMemcpyDeviceToHost, s));
        } else {
            CU(e, cudaMemcpyAsync(forces_out + 6 * N * f, e->d_forces + 6 * N * f, sizeof(double) * 6 * N * nb, cudaMemcpyDeviceToHost, s));
        }
        if (status) CU(e, cudaMemcpyAsync(status + f, e->d_status + f, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
        if (iters) CU(e, cudaMemcpyAsync(iters + f, e->d_iters + f, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < mpc_b200_engine::kPipe; ++i) CU(e, cudaStreamSynchronize(e->pipe[i]));
    drain.armed = false;
    return MPC_B200_OK;
}

int mpc_b200_tron1_solve_host(mpc_b200_engine* e, int B, const double* x0, const double* x_ref, const double* feet,
                              const uint8_t* contact, const int32_t* iter, double* forces, int32_t* status, int32_t* iters) {
    if (!e || !x0 || !x_ref || !feet || !forces || B < 1) return set_err(e, MPC_B200_EINVAL, "solve_host: bad argument");
    if ((contact == nullptr) == (iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "solve_host: pass exactly one of contact / iter");
    if (B > e->max_batch) return set_err(e, MPC_B200_ECAPACITY, "solve_host: B > max_batch");
    CU(e, cudaSetDevice(e->device));
    return solve_host_impl(e, B, x0, x_ref, nullptr, nullptr, feet, contact, iter, forces, status, iters, false);
}

int mpc_b200_tron1_solve_host_multi(mpc_b200_engine* const* engines, int G, int B, const double* x0, const double* x_ref,
                                    const double* feet, const uint8_t* contact, const int32_t* iter, double* forces,
                                    int32_t* status, int32_t* iters) {
    if (!engines || G < 1 || G > 64 || B < 1 || !x0 || !x_ref || !feet || !forces) return MPC_B200_EINVAL;
    if ((contact == nullptr) == (iter == nullptr)) return MPC_B200_EINVAL;
    for (int g = 0; g < G; ++g) {
        if (!engines[g]) return MPC_B200_EINVAL;
        if (engines[g]->N != engines[0]->N || engines[g]->C.per_step_feet != engines[0]->C.per_step_feet)
            return set_err(engines[g], MPC_B200_EINVAL, "solve_host_multi: engines differ in horizon or feet layout");
        for (int h = 0; h < g; ++h)
            if (engines[h]->device == engines[g]->device) return set_err(engines[g], MPC_B200_EINVAL, "solve_host_multi: two engines on one device");
    }
    const int N = engines[0]->N;
    const size_t fstride = (engines[0]->C.per_step_feet && engines[0]->C.ltv) ? 6 * (size_t)N : 6;
    const size_t XR = 13 * (size_t)(N + 1);
    const int base = B / G, extra = B % G;
    std::vector<int> rc((size_t)G, MPC_B200_OK);
    auto work = [&](int g) {
        const int cnt = base + (g < extra ? 1 : 0);
        const size_t f = (size_t)g * base + (size_t)(g < extra ? g : extra);
        if (cnt == 0) return;
        rc[g] = mpc_b200_tron1_solve_host(engines[g], cnt, x0 + 13 * f, x_ref + XR * f, feet + fstride * f,
                                          contact ? contact + 2 * (size_t)N * f : nullptr, iter ? iter + f : nullptr,
                                          forces + 6 * (size_t)N * f, status ? status + f : nullptr, iters ? iters + f : nullptr);
    };
    // one host thread per GPU (the calling thread takes GPU 0): each blocks in its own stream synchronise
    std::vector<std::thread> th;
    th.reserve((size_t)G);
    for (int g = 1; g < G; ++g) th.emplace_back(work, g);
    work(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g) if (rc[g]) return rc[g];
    return MPC_B200_OK;
}

// shared body of the two asynchronous host entries (cmd: controller-shaped, x_ref generated on the device, u0 out)
static int solve_host_async_impl(mpc_b200_engine* e, int B, const double* x0, const double* x_ref, const double* oy, const double* vx,
                                 const double* feet, const uint8_t* contact, const int32_t* iter, double* out, int32_t* status,
                                 int32_t* iters, bool cmd) {
    const int N = e->N;
    const void* a0 = device_view(x0);
    const void* a1 = cmd ? device_view(oy) : device_view(x_ref);
    const void* a2 = cmd ? device_view(vx) : a1;
    const void* a3 = device_view(feet);
    const void* a4 = contact ? device_view(contact) : device_view(iter);
    void* o0 = device_view(out);
    void* o1 = status ? device_view(status) : nullptr;
    void* o2 = iters ? device_view(iters) : nullptr;
    if (!a0 || !a1 || !a2 || !a3 || !a4 || !o0 || (status && !o1) || (iters && !o2))
        return set_err(e, MPC_B200_EINVAL, "asynchronous host entry: every buffer must be pinned (cudaHostAlloc / mpc_b200_pin_host_buffer)");
    const int l = (int)(e->lane_next++ % mpc_b200_engine::kLanes);
    const int cls = schedule_large_class(e, B, contact, iter);
    if (cmd) {
        // 56 result bytes per instance: written by the kernel straight into the pinned host arrays
        const int rc = dispatch_solve(e, B, (const double*)a0, nullptr, (const double*)a3, contact ? (const uint8_t*)a4 : nullptr,
                                      contact ? nullptr : (const int32_t*)a4, (double*)o0, (int32_t*)o1, (int32_t*)o2, e->lane[l],
                                      mpc_b200_engine::kPipe + 1 + l, (1 + l) * e->max_batch, (const double*)a1, (const double*)a2, 1, cls);
        if (rc) return rc;
        e->lane_busy[l] = true;
        return MPC_B200_OK;
    }
    if (!e->d_lane_forces) {
        const size_t L = mpc_b200_engine::kLanes, mb = (size_t)e->max_batch;
        if (cudaMalloc(&e->d_lane_forces, sizeof(double) * 6 * N * mb * L) != cudaSuccess ||
            cudaMalloc(&e->d_lane_status, sizeof(int32_t) * mb * L) != cudaSuccess ||
            cudaMalloc(&e->d_lane_iters, sizeof(int32_t) * mb * L) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(e->d_lane_forces); cudaFree(e->d_lane_status); cudaFree(e->d_lane_iters);
            e->d_lane_forces = nullptr; e->d_lane_status = e->d_lane_iters = nullptr;
            return set_err(e, MPC_B200_ENOMEM, "asynchronous host entry: result buffers");
        }
    }
    const size_t mb = (size_t)e->max_batch;
    double* dF = e->d_lane_forces + 6 * (size_t)N * mb * l;
    int32_t* dS = e->d_lane_status + mb * l;
    int32_t* dI = e->d_lane_iters + mb * l;
    // inputs: read by the kernel straight from the pinned host arrays (TMA bulk copies over PCIe).  Results: device buffers
    // of this lane, then the copy engine (SM-issued stores to host memory are 4x slower than the copy engine on the measured
    // hosts); lanes overlap, so one lane's copy-back runs under another lane's reads -- PCIe is full duplex.
    const int rc = dispatch_solve(e, B, (const double*)a0, (const double*)a1, (const double*)a3, contact ? (const uint8_t*)a4 : nullptr,
                                  contact ? nullptr : (const int32_t*)a4, dF, dS, dI, e->lane[l], mpc_b200_engine::kPipe + 1 + l,
                                  (1 + l) * e->max_batch, nullptr, nullptr, 0, cls);
    if (rc) return rc;
    e->lane_busy[l] = true;
    CU(e, cudaMemcpyAsync(out, dF, sizeof(double) * 6 * N * (size_t)B, cudaMemcpyDeviceToHost, e->lane[l]));
    if (status) CU(e, cudaMemcpyAsync(status, dS, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, e->lane[l]));
    if (iters) CU(e, cudaMemcpyAsync(iters, dI, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, e->lane[l]));
    return MPC_B200_OK;
}

int mpc_b200_tron1_solve_host_async(mpc_b200_engine* e, int B, const double* x0, const double* x_ref, const double* feet,
                                    const uint8_t* contact, const int32_t* iter, double* forces, int32_t* status, int32_t* iters) {
    if (!e || !x0 || !x_ref || !feet || !forces || B < 1) return set_err(e, MPC_B200_EINVAL, "solve_host_async: bad argument");
    if ((contact == nullptr) == (iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "solve_host_async: pass exactly one of contact / iter");
    if (B > e->max_batch) return set_err(e, MPC_B200_ECAPACITY, "solve_host_async: B > max_batch");
    CU(e, cudaSetDevice(e->device));
    return solve_host_async_impl(e, B, x0, x_ref, nullptr, nullptr, feet, contact, iter, forces, status, iters, false);
}

int mpc_b200_tron1_control_host_async(mpc_b200_engine* e, int B, const double* x0, const double* omega_yaw, const double* velocity_x,
                                      const double* feet, const uint8_t* contact, const int32_t* iter, double* u0, int32_t* status,
                                      int32_t* iters) {
    if (!e || !x0 || !omega_yaw || !velocity_x || !feet || !u0 || B < 1) return set_err(e, MPC_B200_EINVAL, "control_host_async: bad argument");
    if ((contact == nullptr) == (iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "control_host_async: pass exactly one of contact / iter");
    if (B > e->max_batch) return set_err(e, MPC_B200_ECAPACITY, "control_host_async: B > max_batch");
    CU(e, cudaSetDevice(e->device));
    return solve_host_async_impl(e, B, x0, nullptr, omega_yaw, velocity_x, feet, contact, iter, u0, status, iters, true);
}

int mpc_b200_wait(mpc_b200_engine* e) {
    if (!e) return MPC_B200_EINVAL;
    CU(e, cudaSetDevice(e->device));
    for (int l = 0; l < mpc_b200_engine::kLanes; ++l) {
        if (!e->lane_busy[l]) continue;
        CU(e, cudaStreamSynchronize(e->lane[l]));
        e->lane_busy[l] = false;
    }
    return MPC_B200_OK;
}

int mpc_b200_tron1_control_host(mpc_b200_engine* e, int B, const double* x0, const double* omega_yaw, const double* velocity_x,
                                const double* feet, const uint8_t* contact, const int32_t* iter, double* u0, int32_t* status,
                                int32_t* iters) {
    if (!e || !x0 || !omega_yaw || !velocity_x || !feet || !u0 || B < 1) return set_err(e, MPC_B200_EINVAL, "control_host: bad argument");
    if ((contact == nullptr) == (iter == nullptr)) return set_err(e, MPC_B200_EINVAL, "control_host: pass exactly one of contact / iter");
    if (B > e->max_batch) return set_err(e, MPC_B200_ECAPACITY, "control_host: B > max_batch");
    CU(e, cudaSetDevice(e->device));
    return solve_host_impl(e, B, x0, nullptr, omega_yaw, velocity_x, feet, contact, iter, u0, status, iters, true);
}

int mpc_b200_tron1_condense_device(mpc_b200_engine* e, int B, const double* d_x0, const double* d_x_ref,
                                   const double* d_feet, double* d_H, double* d_f, double* d_A_aug,
                                   double* d_B_aug, void* stream) {
    if (!e || !d_x0 || !d_x_ref || !d_feet || B < 1) return set_err(e, MPC_B200_EINVAL, "condense: bad argument");
    CU(e, cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (e->N == 10) {
        auto k = tron1_condense_kernel<10, true>;
        size_t smem = sizeof(Tron1Work<10, 60, true>);
        CU(e, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 32, smem, s>>>(e->C, B, d_x0, d_x_ref, d_feet, d_H, d_f, d_A_aug, d_B_aug, nullptr);
    } else if (e->N == 20) {
        auto k = tron1_condense_kernel<20, true>;
        size_t smem = sizeof(Tron1Work<20, 120, true>);
        CU(e, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 32, smem, s>>>(e->C, B, d_x0, d_x_ref, d_feet, d_H, d_f, d_A_aug, d_B_aug, nullptr);
    } else if (e->N == 50) {
        // the 300 x 300 packed factor lives in a temporary global workspace (parity dump only)
        using W50 = Tron1Work<50, 300, false>;
        const size_t need = sizeof(double) * W50::PKN * (size_t)B;
        if (need > e->condense_ws_bytes) {      // kept by the engine: no allocation on repeated calls
            CU(e, cudaStreamSynchronize(s));
            cudaFree(e->d_condense_ws); e->d_condense_ws = nullptr; e->condense_ws_bytes = 0;
            if (cudaMalloc(&e->d_condense_ws, need) != cudaSuccess) { cudaGetLastError(); return set_err(e, MPC_B200_ENOMEM, "condense: workspace"); }
            e->condense_ws_bytes = need;
        }
        double* ws = e->d_condense_ws;
        auto k = tron1_condense_kernel<50, false>;
        size_t smem = sizeof(W50);
        cudaError_t ce = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce == cudaSuccess) {
            k<<<B, 32, smem, s>>>(e->C, B, d_x0, d_x_ref, d_feet, d_H, d_f, d_A_aug, d_B_aug, ws);
            ce = cudaGetLastError();
        }
        if (ce != cudaSuccess) return set_err(e, MPC_B200_ECUDA, "condense<50>", ce);
    } else return set_err(e, MPC_B200_EINVAL, "unsupported horizon");
    CU(e, cudaGetLastError());
    e->launches++;
    return MPC_B200_OK;
}

int mpc_b200_tron1_reference_device(mpc_b200_engine* e, int B, const double* d_x0, const double* d_omega_yaw,
                                    const double* d_velocity_x, double* d_x_ref, void* stream) {
    if (!e || !d_x0 || !d_omega_yaw || !d_velocity_x || !d_x_ref || B < 1) return set_err(e, MPC_B200_EINVAL, "reference: bad argument");
    CU(e, cudaSetDevice(e->device));
    const size_t total = (size_t)B * 13 * (e->N + 1);
    int grid = (int)((total + 255) / 256);
    if (grid > e->num_sms * 16) grid = e->num_sms * 16;
    tron1_reference_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e->C, B, e->N, d_x0, d_omega_yaw, d_velocity_x, d_x_ref);
    CU(e, cudaGetLastError());
    e->launches++;
    return MPC_B200_OK;
}

int mpc_b200_tron1_rollout_device(mpc_b200_engine* e, int B, int steps, double* d_x, const double* d_omega_yaw,
                                  const double* d_velocity_x, const int32_t* d_iter0, double* d_u_traj,
                                  int32_t* d_uncertified, int32_t* d_iters, void* stream) {
    if (!e || !d_x || !d_omega_yaw || !d_velocity_x || !d_iter0 || B < 1 || steps < 1)
        return set_err(e, MPC_B200_EINVAL, "rollout: bad argument");
    if (e->C.per_step_feet) return set_err(e, MPC_B200_EINVAL, "rollout: per_step_feet engines are not supported (nominal feet)");
    CU(e, cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (e->N == 10) {
        rc = launch_rollout_one<10, 30, 1, 4, 4, false>(e, B, steps, d_x, d_omega_yaw, d_velocity_x, d_iter0, d_u_traj, d_uncertified, d_iters, s);
        if (rc) return rc;
        return launch_rollout_one<10, 60, 2, 2, 1, true>(e, B, steps, d_x, d_omega_yaw, d_velocity_x, d_iter0, d_u_traj, d_uncertified, d_iters, s);
    } else if (e->N == 20) {
        rc = launch_rollout_one<20, 60, 2, 2, 2, false>(e, B, steps, d_x, d_omega_yaw, d_velocity_x, d_iter0, d_u_traj, d_uncertified, d_iters, s);
        if (rc) return rc;
        return launch_rollout_one<20, 120, 2, 2, 1, true>(e, B, steps, d_x, d_omega_yaw, d_velocity_x, d_iter0, d_u_traj, d_uncertified, d_iters, s);
    }
    return set_err(e, MPC_B200_EINVAL, "unsupported horizon");
}

int mpc_b200_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return MPC_B200_EINVAL;
    if (device < 0 || device >= mpc_b200_device_count()) return MPC_B200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return MPC_B200_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MPC_B200_ECUDA;
    double* d = nullptr;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, reps = 2000;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return MPC_B200_ENOMEM;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0.0;
    for (int trial = 0; trial < 5; ++trial) {
        cudaEventRecord(a);
        fp64_peak_kernel<<<blocks, threads>>>(d, reps, 0.999999, 1e-9);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return MPC_B200_ECUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double flops = 2.0 * 8 * 16 * (double)reps * blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (trial > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *tflops = best;
    return MPC_B200_OK;
}

}  // extern "C"
