// urgym_kernels.cuh -- sm_100a kernels of the batched UR5e reach simulator (C ABI: urgym_api.cu).
//
// Layout in HBM (per handle, one cudaMalloc): structure-of-arrays planes of 16-byte groups, one env per thread,
// so a warp reads/writes 512 contiguous bytes per plane:
//   qa  float4[N]  q0..q3                           rw every step
//   qb  float4[N]  q4, q5, elapsed (int bits), episode return       rw every step
//   ld4 float4[N]  link_dist 0..3 ; ld1 float[N] link_dist 4        rw every step (not Ori)
//   e4[g] float4[N], e2 float2[N] / e1 float[N]   episode constants E (goal, obstacle poses): read every step,
//                                                   written only at reset
//   va float4[N], vb float2[N]                      ReachDyn.velocity as left by the previous episode (quirk Q4)
// Caller-facing arrays (actions [N,6], obs [N,D], achieved [N,G]) are row-major as the reference's numpy arrays
// are; a block stages its 128-row tile in shared memory so that global traffic is 16-byte vectorised and coalesced.
//
// The step kernel is one fused pass: action -> joints -> obstacle motion -> FK -> collision / link distances ->
// observation -> success / reward / TimeLimit -> statistics -> terminal observation -> auto-reset of finished envs
// (done lanes are compacted with a warp ballot so that the rejection-sampling loop runs in dense warps).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/urgym_b200.h"
#include "ur5e_model_data.h"
#include "urgym_env.cuh"

using namespace urgym;

#define URGYM_BLOCK 128
#define URGYM_STAT_SLOTS 64
#define URGYM_RETURN_SCALE 65536.0f     /* episode returns are summed in 2^-16 fixed point: order-independent */

// The model constants (~1.3 KB) travel as a __grid_constant__ kernel parameter: they sit in the constant bank like
// __constant__ data would, without a global symbol shared between translation units.

// ------------------------------------------------------------------------------------------------ state planes
struct StateView {
    float4 *qa, *qb, *ld4;
    float *ld1;
    float4 *e4[4];
    float2 *e2;
    float *e1;
    float4 *va;
    float2 *vb;
};

template <int TASK> __device__ __forceinline__ void load_E(const StateView &st, int64_t i, float *E) {
    constexpr int EW = Traits<TASK>::EW, NF = EW / 4, TAIL = EW % 4;
#pragma unroll
    for (int g = 0; g < NF; g++) {
        float4 v = st.e4[g][i];
        E[4 * g] = v.x; E[4 * g + 1] = v.y; E[4 * g + 2] = v.z; E[4 * g + 3] = v.w;
    }
    if (TAIL == 2) { float2 v = st.e2[i]; E[4 * NF] = v.x; E[4 * NF + 1] = v.y; }
    if (TAIL == 1) { E[4 * NF] = st.e1[i]; }
}
template <int TASK> __device__ __forceinline__ void store_E(const StateView &st, int64_t i, const float *E) {
    constexpr int EW = Traits<TASK>::EW, NF = EW / 4, TAIL = EW % 4;
#pragma unroll
    for (int g = 0; g < NF; g++) st.e4[g][i] = make_float4(E[4 * g], E[4 * g + 1], E[4 * g + 2], E[4 * g + 3]);
    if (TAIL == 2) st.e2[i] = make_float2(E[4 * NF], E[4 * NF + 1]);
    if (TAIL == 1) st.e1[i] = E[4 * NF];
}
template <int TASK> __device__ __forceinline__ void load_dyn(const StateView &st, int64_t i, EnvState &s) {
    float4 a = st.qa[i], b = st.qb[i];
    s.q[0] = a.x; s.q[1] = a.y; s.q[2] = a.z; s.q[3] = a.w; s.q[4] = b.x; s.q[5] = b.y;
    s.elapsed = __float_as_int(b.z);
    s.ep_ret = b.w;
    if (Traits<TASK>::HAS_OBST) {
        float4 l = st.ld4[i];
        s.ld[0] = l.x; s.ld[1] = l.y; s.ld[2] = l.z; s.ld[3] = l.w; s.ld[4] = st.ld1[i];
    } else {
#pragma unroll
        for (int k = 0; k < 5; k++) s.ld[k] = 0.0f;
    }
}
template <int TASK> __device__ __forceinline__ void store_dyn(const StateView &st, int64_t i, const EnvState &s) {
    st.qa[i] = make_float4(s.q[0], s.q[1], s.q[2], s.q[3]);
    st.qb[i] = make_float4(s.q[4], s.q[5], __int_as_float(s.elapsed), s.ep_ret);
    if (Traits<TASK>::HAS_OBST) {
        st.ld4[i] = make_float4(s.ld[0], s.ld[1], s.ld[2], s.ld[3]);
        st.ld1[i] = s.ld[4];
    }
}

struct StepArgs {
    StateView st;
    int64_t n, offset;
    uint2 key;
    uint32_t event;
    int autoreset;
    const float *actions;
    float *obs, *ach, *des, *rew;
    uint8_t *term, *trunc, *succ;
    float *tobs, *tach;
    unsigned long long *stats;
    const float4 *hull;
};

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// stage the hull vertices in shared memory (hull mode only): every lane of a warp walks the same vertex list, so
// the support-function loop reads shared memory as a broadcast
template <int GEOM> __device__ __forceinline__ const float4 *stage_hull(const float4 *g, float4 *s) {
    if (GEOM != GEOM_HULL) return nullptr;
    for (int i = threadIdx.x; i < UR5E_NUM_HULL_VERTS; i += blockDim.x) s[i] = g[i];
    return s;
}
template <int TASK, int GEOM> constexpr size_t step_smem_bytes() {
    return (size_t)URGYM_BLOCK * (Traits<TASK>::OBS + 6) * sizeof(float) + URGYM_BLOCK * sizeof(int) +
           (GEOM == GEOM_HULL ? (size_t)UR5E_NUM_HULL_VERTS * sizeof(float4) : 0);
}

// ------------------------------------------------------------------------------------------------ step
template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_step_kernel(const __grid_constant__ ModelConst c_model, const StepArgs A) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS, G = TT::GOAL, B = URGYM_BLOCK;
    extern __shared__ float4 smem4[];
    float *s_obs = reinterpret_cast<float *>(smem4);          // [B][D]
    float *s_act = s_obs + B * D;                             // [B][6]
    int *s_list = reinterpret_cast<int *>(s_act + B * 6);     // [B] rows that finished this step
    float4 *s_hull = reinterpret_cast<float4 *>(s_list + B);
    __shared__ int s_ndone;
    __shared__ unsigned long long s_stats[URGYM_STATS_COUNT];

    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * B;
    const int rows = (A.n - base) < B ? (int)(A.n - base) : B;
    if (tid < URGYM_STATS_COUNT) s_stats[tid] = 0ull;
    if (tid == 0) s_ndone = 0;

    // action tile: 16-byte vectorised, coalesced
    const float *gact = A.actions + base * 6;
    if (rows == B && aligned16(gact)) {
        const float4 *g4 = reinterpret_cast<const float4 *>(gact);
        float4 *s4 = reinterpret_cast<float4 *>(s_act);
        for (int k = tid; k < B * 6 / 4; k += B) s4[k] = __ldcs(g4 + k);
    } else {
        for (int k = tid; k < rows * 6; k += B) s_act[k] = gact[k];
    }
    const float4 *hv = stage_hull<GEOM>(A.hull, s_hull);
    __syncthreads();

    const bool active = tid < rows;
    const int64_t i = base + tid;
    float *row = s_obs + tid * D;
    bool will_reset = false;
    if (active) {
        EnvState s;
        StepOut o;
        float vel[6];
        load_dyn<TASK>(A.st, i, s);
        load_E<TASK>(A.st, i, s.E);
        const int len_before = s.elapsed;
        env_step<TASK, GEOM>(c_model, s, s_act + tid * 6, hv, row, o, vel);
        A.rew[i] = o.reward;
        A.term[i] = o.terminated ? 1 : 0;
        A.trunc[i] = o.truncated ? 1 : 0;
        A.succ[i] = o.success ? 1 : 0;
        const bool done = o.terminated || o.truncated;
        if (done) {
            atomicAdd(&s_stats[0], 1ull);
            atomicAdd(&s_stats[1], (unsigned long long)__float2ll_rn(s.ep_ret * URGYM_RETURN_SCALE));
            atomicAdd(&s_stats[2], (unsigned long long)(len_before + 1));
            if (o.success) atomicAdd(&s_stats[3], 1ull);
            if (o.collision) atomicAdd(&s_stats[4], 1ull);
            if (o.truncated && !o.terminated) atomicAdd(&s_stats[5], 1ull);
        }
        will_reset = done && A.autoreset;
        if (!will_reset) store_dyn<TASK>(A.st, i, s);
    }
    // compaction of the finished rows: warp ballot, one shared-memory atomic per warp
    {
        const unsigned m = __ballot_sync(0xffffffffu, will_reset);
        if (m) {
            const int lane = tid & 31;
            int pos = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) pos = atomicAdd(&s_ndone, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (will_reset) s_list[pos + __popc(m & ((1u << lane) - 1u))] = tid;
        }
    }
    __syncthreads();
    const int nd = s_ndone;
    if (nd > 0) {
        // terminal observations (DummyVecEnv's info["terminal_observation"]): whole rows, coalesced per row
        if (A.tobs) {
            for (int k = tid; k < nd * D; k += B) {
                const int r = s_list[k / D], c = k % D;
                A.tobs[(base + r) * D + c] = s_obs[r * D + c];
            }
        }
        if (A.tach) {
            for (int k = tid; k < nd * G; k += B) {
                const int r = s_list[k / G], c = k % G;
                A.tach[(base + r) * G + c] = s_obs[r * D + c];
            }
        }
        __syncthreads();
        // auto-reset: thread t takes the t-th finished row, so the rejection loops run in dense warps
        if (tid < nd) {
            const int r = s_list[tid];
            const int64_t gi = base + r;
            const uint64_t genv = (uint64_t)(A.offset + gi);
            EnvState ns;
            ResetStream rs;
            rs.key = A.key; rs.episode = A.event; rs.env_lo = (uint32_t)genv; rs.env_hi = (uint32_t)(genv >> 32);
            rs.bpi = TT::BPI; rs.iter = 0;
            float *nrow = s_obs + r * D;
            if (TT::DYN) {      // the previous episode's last velocity stays visible until the next step (quirk Q4)
                A.st.va[gi] = make_float4(nrow[24], nrow[25], nrow[26], nrow[27]);
                A.st.vb[gi] = make_float2(nrow[28], nrow[29]);
            }
            const int iters = env_reset<TASK, GEOM>(c_model, ns, rs, hv, nrow);
            store_dyn<TASK>(A.st, gi, ns);
            store_E<TASK>(A.st, gi, ns.E);
            atomicAdd(&s_stats[7], (unsigned long long)iters);
        }
        __syncthreads();
    }
    // observation tile -> global, 16-byte vectorised
    float *gobs = A.obs + base * D;
    if (rows == B && aligned16(gobs)) {
        float4 *g4 = reinterpret_cast<float4 *>(gobs);
        const float4 *s4 = reinterpret_cast<const float4 *>(s_obs);
        for (int k = tid; k < B * D / 4; k += B) __stcs(g4 + k, s4[k]);
    } else {
        for (int k = tid; k < rows * D; k += B) gobs[k] = s_obs[k];
    }
    if (A.ach) {        // achieved_goal = ee position (+ Euler) = first G observation columns
        float *g = A.ach + base * G;
        for (int k = tid; k < rows * G; k += B) g[k] = s_obs[(k / G) * D + (k % G)];
    }
    if (A.des) {        // desired_goal = goal = observation columns 12..12+G
        float *g = A.des + base * G;
        for (int k = tid; k < rows * G; k += B) g[k] = s_obs[(k / G) * D + 12 + (k % G)];
    }
    if (tid == 0) atomicAdd(&s_stats[6], (unsigned long long)rows);
    __syncthreads();
    if (tid < URGYM_STATS_COUNT && s_stats[tid] != 0ull)
        atomicAdd(&A.stats[(blockIdx.x % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT + tid], s_stats[tid]);
}

// ------------------------------------------------------------------------------------------------ reset / observe / refresh
struct AuxArgs {
    StateView st;
    int64_t n, offset;
    uint2 key;
    uint32_t event;
    const uint8_t *mask;
    float *obs, *ach, *des;
    uint8_t *collision;
    unsigned long long *stats;
    const float4 *hull;
};

template <int TASK> __device__ __forceinline__ void write_rows(const AuxArgs &A, int64_t i, const float *row) {
    constexpr int D = Traits<TASK>::OBS, G = Traits<TASK>::GOAL;
    if (A.obs) for (int k = 0; k < D; k++) A.obs[i * D + k] = row[k];
    if (A.ach) for (int k = 0; k < G; k++) A.ach[i * G + k] = row[k];
    if (A.des) for (int k = 0; k < G; k++) A.des[i * G + k] = row[12 + k];
}

template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_reset_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    typedef Traits<TASK> TT;
    extern __shared__ float4 smem4[];
    const float4 *hv = stage_hull<GEOM>(A.hull, smem4);
    if (GEOM == GEOM_HULL) __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n || (A.mask && !A.mask[i])) return;
    float row[TT::OBS];
    EnvState s;
    if (TT::DYN) {
        // ReachDyn.velocity survives the reset (quirk Q4): carry the value the previous episode ended with
        load_dyn<TASK>(A.st, i, s);
        load_E<TASK>(A.st, i, s.E);
        float vel[6];
        if (s.elapsed == 0) {
            float4 a = A.st.va[i]; float2 b = A.st.vb[i];
            vel[0] = a.x; vel[1] = a.y; vel[2] = a.z; vel[3] = a.w; vel[4] = b.x; vel[5] = b.y;
        } else {
            Quat qs; float3 axis; float angle; float tw[6];
            dyn_twist(&s.E[6], &s.E[12], tw, qs, axis, angle);
            for (int k = 0; k < 6; k++) vel[k] = s.elapsed <= 25 ? tw[k] : 0.0f;
        }
        A.st.va[i] = make_float4(vel[0], vel[1], vel[2], vel[3]);
        A.st.vb[i] = make_float2(vel[4], vel[5]);
        for (int k = 0; k < 6; k++) row[24 + k] = vel[k];
    }
    const uint64_t genv = (uint64_t)(A.offset + i);
    ResetStream rs;
    rs.key = A.key; rs.episode = A.event; rs.env_lo = (uint32_t)genv; rs.env_hi = (uint32_t)(genv >> 32);
    rs.bpi = TT::BPI; rs.iter = 0;
    const int iters = env_reset<TASK, GEOM>(c_model, s, rs, hv, row);
    store_dyn<TASK>(A.st, i, s);
    store_E<TASK>(A.st, i, s.E);
    write_rows<TASK>(A, i, row);
    atomicAdd(&A.stats[(blockIdx.x % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT + 7], (unsigned long long)iters);
}

template <int TASK>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_observe_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    typedef Traits<TASK> TT;
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    EnvState s;
    load_dyn<TASK>(A.st, i, s);
    load_E<TASK>(A.st, i, s.E);
    float stale[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (TT::DYN) {
        float4 a = A.st.va[i]; float2 b = A.st.vb[i];
        stale[0] = a.x; stale[1] = a.y; stale[2] = a.z; stale[3] = a.w; stale[4] = b.x; stale[5] = b.y;
    }
    float row[TT::OBS];
    env_observe<TASK, GEOM_CAPSULE>(c_model, s, stale, row);
    write_rows<TASK>(A, i, row);
}

template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_refresh_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    extern __shared__ float4 smem4[];
    const float4 *hv = stage_hull<GEOM>(A.hull, smem4);
    if (GEOM == GEOM_HULL) __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    EnvState s;
    load_dyn<TASK>(A.st, i, s);
    load_E<TASK>(A.st, i, s.E);
    const bool coll = env_refresh<TASK, GEOM>(c_model, s, hv);
    store_dyn<TASK>(A.st, i, s);
    if (A.collision) A.collision[i] = coll ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ launchers
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + URGYM_BLOCK - 1) / URGYM_BLOCK); }

template <int TASK, int GEOM> cudaError_t launch_step(const ModelConst &M, const StepArgs &A, cudaStream_t s) {
    const size_t smem = step_smem_bytes<TASK, GEOM>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(urgym_step_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    urgym_step_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, smem, s>>>(M, A);
    return cudaGetLastError();
}
template <int TASK, int GEOM> cudaError_t launch_reset(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    const size_t smem = GEOM == GEOM_HULL ? (size_t)UR5E_NUM_HULL_VERTS * sizeof(float4) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(urgym_reset_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    urgym_reset_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, smem, s>>>(M, A);
    return cudaGetLastError();
}
template <int TASK, int GEOM> cudaError_t launch_refresh(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    const size_t smem = GEOM == GEOM_HULL ? (size_t)UR5E_NUM_HULL_VERTS * sizeof(float4) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(urgym_refresh_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    urgym_refresh_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, smem, s>>>(M, A);
    return cudaGetLastError();
}


template <int TASK> cudaError_t launch_observe(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    urgym_observe_kernel<TASK><<<grid_for(A.n), URGYM_BLOCK, 0, s>>>(M, A);
    return cudaGetLastError();
}

// one translation unit per (task, geometry) instantiates these (urgym_inst.cu), urgym_api.cu calls them by table
typedef cudaError_t (*step_launcher_t)(const ModelConst &, const StepArgs &, cudaStream_t);
typedef cudaError_t (*aux_launcher_t)(const ModelConst &, const AuxArgs &, cudaStream_t);
#define URGYM_DECLARE_INST(T, G)                                                           \
    cudaError_t urgym_inst_step_##T##_##G(const ModelConst &, const StepArgs &, cudaStream_t);   \
    cudaError_t urgym_inst_reset_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);   \
    cudaError_t urgym_inst_refresh_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);
URGYM_DECLARE_INST(0, 0) URGYM_DECLARE_INST(1, 0) URGYM_DECLARE_INST(2, 0) URGYM_DECLARE_INST(3, 0)
URGYM_DECLARE_INST(0, 1) URGYM_DECLARE_INST(1, 1) URGYM_DECLARE_INST(2, 1) URGYM_DECLARE_INST(3, 1)
