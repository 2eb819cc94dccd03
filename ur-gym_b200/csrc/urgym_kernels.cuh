// urgym_kernels.cuh -- sm_100a kernels of the batched UR5e reach simulator (C ABI: urgym_api.cu).
//
// Layout in HBM (per handle, one cudaMalloc): structure-of-arrays planes, one env per thread; the planes that a reset
// rewrites env by env are 32-byte groups (one 256-bit access per lane, 1 KB of contiguous memory per warp and group):
//   q8  [2N] float4  q0..q3 | q4, q5, elapsed (int bits), episode return        rw every step
//   ld4 [N] float4, ld1 [N] float   link_dist = last_dist                         rw every step (not Ori)
//   h8[g] [2N] float4  hot words: the episode constants the step reads + the episode cache   read every step
//   e8[g] [2N] float4  episode constants E as the reference holds them             cold (reset, state access)
//   v8  [2N] float4  ReachDyn.velocity as left by the previous episode (quirk Q4)   reset only
// Caller-facing arrays (actions [N,6], obs [N,D], achieved [N,G]) are row-major as the reference's numpy arrays
// are; a warp stages its 32 observation rows in shared memory so that they leave as 16-byte coalesced stores.
//
// Per env step two kernels: urgym_step_kernel (action -> joints -> obstacle motion -> FK -> collision / link distances
// -> observation -> success / reward / TimeLimit -> statistics -> terminal rows, finished envs appended to a queue)
// and urgym_autoreset_kernel (the queued envs, dense: 32 per warp, one per lane).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/urgym_b200.h"
#include "ur5e_model_data.h"
#include "urgym_env.cuh"

using namespace urgym;

#define URGYM_BLOCK 128
// Threads per block of the capsule-geometry step kernel.  Its warps are independent (no block-wide barrier, private
// tiles), so a block is one warp: a slot frees as soon as its warp retires instead of waiting for the slowest of four
// (achieved occupancy 20.6 -> 21.5 of 24 warps, shorter tail; 128 -> 32 threads: -2.2 % kernel time, 64: -0.3 %, 256: +2.4 %).
#ifndef URGYM_STEP_BLOCK
#define URGYM_STEP_BLOCK 32
#endif
#define URGYM_MAX_CHAINS 8     /* independent step chains (env sub-ranges advanced on their own streams) */
#ifndef URGYM_STEP_MINBLOCKS
#define URGYM_STEP_MINBLOCKS (736 / URGYM_STEP_BLOCK)   /* 23 resident step-kernel warps per SM, 88 registers (24 warps at 80: +0.3 %; 22: +0.1 %; 20, 26: +2.5 %) */
#endif
#define URGYM_STAT_SLOTS 64
#define URGYM_RETURN_SCALE 65536.0f     /* episode returns are summed in 2^-16 fixed point: order-independent */

// The model constants (~3 KB) travel as a __grid_constant__ kernel parameter: they sit in the constant bank like
// __constant__ data would, without a global symbol shared between translation units.

// ------------------------------------------------------------------------------------------------ state planes
// Planes that the auto-reset kernel writes env by env are kept in 32-BYTE groups per env (two float4), so that a reset
// writes whole 32-byte sectors: the ECC-protected HBM can only write a partial sector by read-modify-write, and 18
// partial stores per reset were what that kernel spent its time waiting for.  A warp of the step kernel still moves
// 1 KB of contiguous memory per group.
struct StateView {
    float4 *q8;             // [2N]  q0..q3 | q4, q5, elapsed (int bits), episode return
    float4 *ld4;            // [N]   link_dist 0..3
    float *ld1;             // [N]   link_dist 4
    float4 *e8[3];          // [2N]  episode constants E, 8 words per group
    float4 *v8;             // [2N]  ReachDyn.velocity carried over a reset (6 words)
    float4 *h8[3];          // [2N]  hot words H = E[0..EH) ++ C, 8 words per group
};

// One 256-bit access per 32-byte group (LDG.E.256 / STG.E.256, new with sm_100): a warp reads or writes 1 KB of
// contiguous memory per instruction.  p points at the env's pair of float4 (32-byte aligned).
__device__ __forceinline__ void ld_group(const float4 *p, float4 &a, float4 &b) {
    // .cg: the planes are read once per step, no use keeping them in L1 (UR5DynReach -0.5 %, UR5StaReach -0.4 %)
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void st_group(float4 *p, float4 a, float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}
template <int TASK> __device__ __forceinline__ void load_E(const StateView &st, int64_t i, float *E) {
    constexpr int EW = Traits<TASK>::EW;
#pragma unroll
    for (int g = 0; 4 * g < EW; g++) {
        const float4 v = st.e8[g >> 1][2 * i + (g & 1)];
        E[4 * g] = v.x;
        if (4 * g + 1 < EW) E[4 * g + 1] = v.y;
        if (4 * g + 2 < EW) E[4 * g + 2] = v.z;
        if (4 * g + 3 < EW) E[4 * g + 3] = v.w;
    }
}
template <int TASK> __device__ __forceinline__ void store_E(const StateView &st, int64_t i, const float *E) {
    constexpr int EW = Traits<TASK>::EW, NG = (EW + 7) / 8;      // whole 32-byte groups are written (padding: zeros)
#pragma unroll
    for (int G = 0; G < NG; G++) {
        auto w = [&](int k) { return 8 * G + k < EW ? E[8 * G + k] : 0.0f; };
        st_group(st.e8[G] + 2 * i, make_float4(w(0), w(1), w(2), w(3)), make_float4(w(4), w(5), w(6), w(7)));
    }
}
// hot words H = E[0..EH) ++ C[0..CW)
template <int TASK> struct HotLayout {
    static constexpr int EH = Traits<TASK>::EH, CW = Traits<TASK>::CW, HW = EH + CW, NG = (HW + 7) / 8;
    static_assert(NG <= 3, "hot planes");
};
template <int TASK> __device__ __forceinline__ float &hot_word(EnvState &s, int w) {
    return w < HotLayout<TASK>::EH ? s.E[w] : s.C[w - HotLayout<TASK>::EH];
}
template <int TASK> __device__ __forceinline__ void load_hot(const StateView &st, int64_t i, EnvState &s) {
    typedef HotLayout<TASK> L;
#pragma unroll
    for (int G = 0; G < L::NG; G++) {
        float4 v[2];
        ld_group(st.h8[G] + 2 * i, v[0], v[1]);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int w = 8 * G + 4 * h;
            if (w < L::HW) hot_word<TASK>(s, w) = v[h].x;
            if (w + 1 < L::HW) hot_word<TASK>(s, w + 1) = v[h].y;
            if (w + 2 < L::HW) hot_word<TASK>(s, w + 2) = v[h].z;
            if (w + 3 < L::HW) hot_word<TASK>(s, w + 3) = v[h].w;
        }
    }
}
template <int TASK> __device__ __forceinline__ void store_hot(const StateView &st, int64_t i, EnvState &s) {
    typedef HotLayout<TASK> L;
#pragma unroll
    for (int G = 0; G < L::NG; G++) {
        auto w = [&](int k) { return 8 * G + k < L::HW ? hot_word<TASK>(s, 8 * G + k) : 0.0f; };
        st_group(st.h8[G] + 2 * i, make_float4(w(0), w(1), w(2), w(3)), make_float4(w(4), w(5), w(6), w(7)));
    }
}
__device__ __forceinline__ void load_vel(const StateView &st, int64_t i, float *vel) {
    const float4 a = st.v8[2 * i], b = st.v8[2 * i + 1];
    vel[0] = a.x; vel[1] = a.y; vel[2] = a.z; vel[3] = a.w; vel[4] = b.x; vel[5] = b.y;
}
__device__ __forceinline__ void store_vel(const StateView &st, int64_t i, const float *vel) {
    st_group(st.v8 + 2 * i, make_float4(vel[0], vel[1], vel[2], vel[3]), make_float4(vel[4], vel[5], 0.0f, 0.0f));
}
// link_dist = last_dist: read by the observation row and the reward, i.e. late in the step
template <int TASK> __device__ __forceinline__ void load_ld(const StateView &st, int64_t i, EnvState &s) {
    if (Traits<TASK>::HAS_OBST) {
        float4 l = __ldcg(st.ld4 + i);
        s.ld[0] = l.x; s.ld[1] = l.y; s.ld[2] = l.z; s.ld[3] = l.w; s.ld[4] = __ldcg(st.ld1 + i);
    } else {
#pragma unroll
        for (int k = 0; k < 5; k++) s.ld[k] = 0.0f;
    }
}
template <int TASK> __device__ __forceinline__ void load_dyn(const StateView &st, int64_t i, EnvState &s) {
    float4 a, b;
    ld_group(st.q8 + 2 * i, a, b);
    s.q[0] = a.x; s.q[1] = a.y; s.q[2] = a.z; s.q[3] = a.w; s.q[4] = b.x; s.q[5] = b.y;
    s.elapsed = __float_as_int(b.z);
    s.ep_ret = b.w;
    load_ld<TASK>(st, i, s);
}
template <int TASK> __device__ __forceinline__ void store_dyn(const StateView &st, int64_t i, const EnvState &s) {
    st_group(st.q8 + 2 * i, make_float4(s.q[0], s.q[1], s.q[2], s.q[3]),
             make_float4(s.q[4], s.q[5], __int_as_float(s.elapsed), s.ep_ret));
    if (Traits<TASK>::HAS_OBST) {
        st.ld4[i] = make_float4(s.ld[0], s.ld[1], s.ld[2], s.ld[3]);
        st.ld1[i] = s.ld[4];
    }
}

struct StepArgs {
    StateView st;
    int64_t n;
    const float *actions;
    float *obs, *ach, *des, *rew;
    float *tobs, *tach;         // terminal observation / achieved goal rows of the envs this step finishes (or NULL)
    uint8_t *term, *trunc, *succ;
    unsigned long long *stats;
    uint32_t *event;            // device-resident reset-event counters, one per chain (URGYM_MAX_CHAINS)
    int bump;                   // 0: none (the caller bumped), 1: this chain's counter, 2: all counters (a whole step)
    int chain;
    const float4 *hull;
    int *queue;                 // auto-reset queue of this chain: local indices of the envs this step finished (or NULL)
    unsigned *qcount;           // its entry count
};

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifndef URGYM_AUTORESET_BLOCK
#define URGYM_AUTORESET_BLOCK 32    /* one warp per block: the warps finish at very different times (rejection loops) */
#endif
#ifndef URGYM_HULL_BLOCK
#define URGYM_HULL_BLOCK 512
#endif
#ifndef URGYM_HULL_AUTORESET_BLOCK
#define URGYM_HULL_AUTORESET_BLOCK 128  /* every block stages the hull blob: 32 -> 568 us, 64 -> 377, 128 -> 254, 256 -> 264 (262144 Dyn envs) */
#endif
// Block sizes.  Capsule geometry: 128-thread blocks, six resident per SM.  Hull geometry: every block stages the 111 KB
// hull blob (vertices + adjacency, urgym_device.cuh) in shared memory, so one big block per SM shares one copy.
template <int GEOM> struct Blk {
    static constexpr bool HULL = URGYM_BASE(GEOM) == GEOM_HULL;
    static constexpr int STEP = HULL ? URGYM_HULL_BLOCK : URGYM_STEP_BLOCK;
    static constexpr int STEP_MINBLOCKS = HULL ? 1 : URGYM_STEP_MINBLOCKS;
    static constexpr int AUTORESET = HULL ? URGYM_HULL_AUTORESET_BLOCK : URGYM_AUTORESET_BLOCK;
};
// stage the hull blob in shared memory (hull mode only): the hill-climbing support function gathers from it per lane
template <int GEOM> __device__ __forceinline__ const float4 *stage_hull(const float4 *g, float4 *s) {
    if (URGYM_BASE(GEOM) != GEOM_HULL) return nullptr;
    for (int i = threadIdx.x; i < URGYM_HULL_BLOB_F4; i += blockDim.x) s[i] = g[i];
    return s;
}
// per-warp tile: 32 observation rows; the capsule pass's scratch column block [41][32] overlays it (capsule geometry only)
template <int TASK, int GEOM> struct TileFloats {
    static constexpr int value = (URGYM_BASE(GEOM) == GEOM_HULL || Traits<TASK>::OBS > URGYM_SCRATCH_FLOATS)
                                     ? Traits<TASK>::OBS : URGYM_SCRATCH_FLOATS;
};
// Hull geometry: the exact GJK tests that the cheap stages of a block's 512 envs leave over, compacted into three dense
// task lists in shared memory (hull_pass_phased).  A task is  env (bits 0..15) | link << 16 | (box or second link) << 20.
#define URGYM_HULL_QCAP 2048
struct HullTasks {
    int count[4];                       // [0] link vs cylinder caps / rims  [1] link vs table / track  [2] self pairs
    int list[3][URGYM_HULL_QCAP];
};
template <int TASK, int GEOM> constexpr size_t step_smem_bytes() {
    return (size_t)Blk<GEOM>::STEP * TileFloats<TASK, GEOM>::value * sizeof(float) +
           (URGYM_BASE(GEOM) == GEOM_HULL ? (size_t)URGYM_HULL_BLOB_F4 * sizeof(float4) + sizeof(HullTasks) : 0);
}
__device__ __forceinline__ void stat_add(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }   // RED.E.ADD.64 (result unused)

// ------------------------------------------------------------------------------------------------ hull geometry pass
// The robot pass of the hull-geometry step kernel (same tests and arithmetic as robot_pass_rolled, urgym_env.cuh).  One
// GJK per lane with data-dependent work kept 5 of 32 lanes busy: a few lanes of every warp need the cylinder's caps, the
// exact table / track test or an exact self-collision test while the others wait.  Here every lane first runs only the
// stages that all lanes run (the chain walk, GJK against the obstacle's axis segment, the capsule broad phases) and
// queues what is left; then the whole block works through the three task lists with one task per lane, recomputing the
// link pose from the env's joint angles (240 instructions against the thousands of a GJK); then every lane collects its
// results.  Scratch per env, in its observation row until env_step_finish overwrites it: [0..5) link-obstacle
// distances, [5] collision flag, [6..12) joint angles (env_step_begin), [12..18) obstacle centre and axis.
static __device__ __noinline__ bool hull_self_exact(const ModelConst &M, const float *qrow, int l, int l2, const float4 *hv) {
    Pose T;
    pose_identity(T);
    LinkShape<GEOM_HULL> a, b;
    for (int j = 1; j <= l; j++) {
        fk_advance(M, T, j - 1, qrow[j - 1]);
        if (j == l2) b.set(M, l2, T, hv);
    }
    a.set(M, l, T, hv);
    return a.link_hit_exact(M, l, l2, b);
}
template <int TASK, int GEOM>
__device__ __forceinline__ bool hull_pass_phased(const ModelConst &M, float *rows, int tid, const ObstW &O, const float4 *hv,
                                                 float *ee, float *dist, HullTasks &Q) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS;
    float *row = rows + tid * D;
    auto push = [&](int type, int code) {
        const int slot = atomicAdd(&Q.count[type], 1);
        if (slot < URGYM_HULL_QCAP) Q.list[type][slot] = code;
        return slot < URGYM_HULL_QCAP;          // a full list: the caller does the test itself
    };
    Pose T;
    pose_identity(T);
    LinkShape<GEOM_CAPSULE> c1, c2, c3;
    bool hit = false;
    float wb[5] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
#pragma unroll 1
    for (int l = 1; l < 7; l++) {
        fk_advance(M, T, l - 1, row[6 + l - 1]);
        LinkShape<GEOM_HULL> cur;
        cur.set(M, l, T, hv);
        if (l >= 2) {
            if (TT::HAS_OBST) {
                float d;
                if (!cur.obstacle_dist_side(M, l, O, d)) {
                    d = 3.0e38f;
                    if (!push(0, tid | (l << 16))) d = cur.obstacle_dist_caps(M, l, O);
                }
                row[l - 2] = d;
                if (URGYM_WB(GEOM)) {
                    const float w = fminf(cur.box_dist(M, l, 0), cur.box_dist(M, l, 1));
                    if (l == 2) wb[0] = w; else if (l == 3) wb[1] = w; else if (l == 4) wb[2] = w; else if (l == 5) wb[3] = w; else wb[4] = w;
                }
            }
#pragma unroll 1
            for (int box = 0; box < 2; box++)
                if (cur.cap.box_hit(M, l, box) && !push(1, tid | (l << 16) | (box << 20))) hit = hit || cur.box_hit_exact(M, l, box);
        }
        if (l >= 3 && cur.cap.link_hit(M, l, 1, c1) && !push(2, tid | (l << 16) | (1 << 20))) hit = hit || hull_self_exact(M, row + 6, l, 1, hv);
        if (l >= 4 && cur.cap.link_hit(M, l, 2, c2) && !push(2, tid | (l << 16) | (2 << 20))) hit = hit || hull_self_exact(M, row + 6, l, 2, hv);
        if (l >= 5 && cur.cap.link_hit(M, l, 3, c3) && !push(2, tid | (l << 16) | (3 << 20))) hit = hit || hull_self_exact(M, row + 6, l, 3, hv);
        if (l == 1) c1 = cur.cap; else if (l == 2) c2 = cur.cap; else if (l == 3) c3 = cur.cap;
    }
    const float3 e = euler_from_mat(T.R);
    ee[0] = T.p.x; ee[1] = T.p.y; ee[2] = T.p.z; ee[3] = e.x; ee[4] = e.y; ee[5] = e.z;
    row[5] = hit ? 1.0f : 0.0f;
    if (TT::HAS_OBST) { row[12] = O.c.x; row[13] = O.c.y; row[14] = O.c.z; row[15] = O.u.x; row[16] = O.u.y; row[17] = O.u.z; }
    __syncthreads();
    const int nthreads = blockDim.x;
    if (TT::HAS_OBST) {
        const int n0 = min(Q.count[0], URGYM_HULL_QCAP);
        for (int k = tid; k < n0; k += nthreads) {
            const int code = Q.list[0][k], en = code & 0xFFFF, l = (code >> 16) & 15;
            float *re = rows + en * D;
            ObstW Oe = obstacle_none();
            Oe.c = f3(re[12], re[13], re[14]); Oe.u = f3(re[15], re[16], re[17]);
            Pose Tl;
            fk_link(M, re + 6, l, Tl);
            LinkShape<GEOM_HULL> L;
            L.set(M, l, Tl, hv);
            re[l - 2] = L.obstacle_dist_caps(M, l, Oe);
        }
    }
    {
        const int n1 = min(Q.count[1], URGYM_HULL_QCAP);
        for (int k = tid; k < n1; k += nthreads) {
            const int code = Q.list[1][k], en = code & 0xFFFF, l = (code >> 16) & 15, box = (code >> 20) & 1;
            float *re = rows + en * D;
            Pose Tl;
            fk_link(M, re + 6, l, Tl);
            LinkShape<GEOM_HULL> L;
            L.set(M, l, Tl, hv);
            if (L.box_hit_exact(M, l, box)) re[5] = 1.0f;
        }
        const int n2 = min(Q.count[2], URGYM_HULL_QCAP);
        for (int k = tid; k < n2; k += nthreads) {
            const int code = Q.list[2][k], en = code & 0xFFFF, l = (code >> 16) & 15, l2 = (code >> 20) & 3;
            float *re = rows + en * D;
            if (hull_self_exact(M, re + 6, l, l2, hv)) re[5] = 1.0f;
        }
    }
    __syncthreads();
    hit = row[5] != 0.0f;
    if (TT::HAS_OBST) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float d = row[k];
            hit = hit || (d <= URGYM_COLLISION_MARGIN);     // keys[5] == 'obstacle'   pyb_setup.py:398-404
            dist[k] = URGYM_WB(GEOM) ? fminf(d, wb[k]) : d;
        }
    }
    return hit;
}

// ------------------------------------------------------------------------------------------------ step
// One env per thread; every warp owns a private 32-row tile of the observation array in shared memory and there is
// no block-wide synchronisation (apart from the hull staging in hull mode): warps start, run and retire on their own.
// Finished envs are NOT reset here: the step only raises their terminated / truncated flags, and the auto-reset
// kernel that follows in the stream handles them in dense form.
template <int TASK, int GEOM>
__global__ void __launch_bounds__(Blk<GEOM>::STEP, Blk<GEOM>::STEP_MINBLOCKS) urgym_step_kernel(const __grid_constant__ ModelConst c_model, const StepArgs A) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS, G = TT::GOAL, B = Blk<GEOM>::STEP, W = 32;
    constexpr int TF = TileFloats<TASK, GEOM>::value;
    constexpr bool HULL = URGYM_BASE(GEOM) == GEOM_HULL;
    extern __shared__ float4 smem4[];
    float *s_tiles = reinterpret_cast<float *>(smem4);        // [warps][32 * TF]: obs tile [32][D]
    float4 *s_hull = reinterpret_cast<float4 *>(s_tiles + B * TF);
    HullTasks *s_tasks = reinterpret_cast<HullTasks *>(s_hull + URGYM_HULL_BLOB_F4);      // hull geometry only

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4 *hv = stage_hull<GEOM>(A.hull, s_hull);
    if (HULL) {
        if (tid < 4) s_tasks->count[tid] = 0;
        __syncthreads();
    }

    // Env indices of a launch fit 32 bits (the launcher checks it).  With 32-bit index arithmetic every plane address is one
    // IMAD.WIDE.U32 instead of an IADD3 / IADD3.X pair and the loads issue ~8 instructions earlier: UR5OriReach -2.6 %,
    // UR5DynReach -0.5 %, UR5ObsReach -0.15 %; UR5StaReach +0.5 % (ptxas schedules it differently), which keeps 64 bits.
#ifdef URGYM_IDX32
    constexpr bool IDX32 = URGYM_IDX32 != 0;
#else
    constexpr bool IDX32 = !HULL && TASK != TASK_STA;
#endif
    int64_t wbase, i;
    int rows;
    if (IDX32) {
        const unsigned n32 = (unsigned)A.n, wb32 = blockIdx.x * (unsigned)B + (unsigned)(warp * W);
        if (wb32 >= n32) return;
        wbase = wb32;
        rows = (n32 - wb32) < (unsigned)W ? (int)(n32 - wb32) : W;
        i = wb32 + (unsigned)(lane < rows ? lane : rows - 1);
    } else {
        wbase = (int64_t)blockIdx.x * B + warp * W;               // first env of this warp
        if (!HULL && wbase >= A.n) return;                       // (hull geometry: every warp of the block takes part in its barriers)
        // (the capsule kernels keep the plain expressions: the extra selects of the hull form cost UR5StaReach 9 % -- ptxas
        // schedules the whole kernel differently around them)
        rows = HULL ? (wbase >= A.n ? 0 : ((A.n - wbase) < W ? (int)(A.n - wbase) : W)) : ((A.n - wbase) < W ? (int)(A.n - wbase) : W);
        i = HULL ? (lane < rows ? wbase + lane : (rows ? wbase + rows - 1 : A.n - 1)) : wbase + (lane < rows ? lane : rows - 1);
    }
    float *s_obs = s_tiles + warp * W * TF;
    float *s_scr = s_obs + lane;                              // capsule scratch: column `lane` of a [41][32] block
    // every lane runs the step (warp-level barriers inside); lanes past the end redo the last env and store nothing
    const bool live = lane < rows;
    EnvState s;
    StepOut o;
    float vel[6], act[6];
    {   // the lane's own action row: 24 contiguous bytes (the warp's three requests cover the same 768 bytes)
        const float *ga = A.actions + i * 6;
        if ((reinterpret_cast<uintptr_t>(A.actions) & 7u) == 0) {
            const float2 a0 = __ldg(reinterpret_cast<const float2 *>(ga)), a1 = __ldg(reinterpret_cast<const float2 *>(ga) + 1),
                         a2 = __ldg(reinterpret_cast<const float2 *>(ga) + 2);
            act[0] = a0.x; act[1] = a0.y; act[2] = a1.x; act[3] = a1.y; act[4] = a2.x; act[5] = a2.y;
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) act[k] = __ldg(ga + k);
        }
    }
    load_hot<TASK>(A.st, i, s);         // first: the obstacle pose is the first thing computed (UR5DynReach -0.3 %, others +-0)
    load_dyn<TASK>(A.st, i, s);
    // (after the loads have been issued: ahead of them, this branch and its reconvergence point delayed every warp's loads;
    // moved here: UR5ObsReach -2.5 %, UR5OriReach -1.0 %, UR5StaReach -0.5 %, UR5DynReach -0.4 % kernel time)
    if (blockIdx.x == 0 && tid == 0) {                          // this launch is reset event number event[chain] + 1
        if (A.bump == 1) A.event[A.chain] += 1u;
        if (A.bump == 2) {      // a whole step: one logical event for all chains (lagging chain counters catch up)
            uint32_t m = 0u;
            for (int c = 0; c < URGYM_MAX_CHAINS; c++) m = A.event[c] > m ? A.event[c] : m;
            for (int c = 0; c < URGYM_MAX_CHAINS; c++) A.event[c] = m + 1u;
        }
    }
    // auto-reset queue: one reservation per warp that finished envs, made as soon as the flags are known (env_step_finish
    // calls the hook ahead of the observation row and the reward); the reply is needed only at the very end
    unsigned done_mask = 0u, qbase = 0u;
    auto reserve = [&](bool finished) {
        done_mask = __ballot_sync(0xffffffffu, live && finished);
        // (inline PTX: around atomicAdd the compiler builds its own warp aggregation, whose shuffle waits for the reply at once)
        // (the count goes through a shuffle so that ptxas does not see a warp-uniform operand: around a uniform atomic it builds
        // its own warp aggregation, whose shuffle waits for the reply at once)
        // (`+ threadIdx.y`, which is 0: around an atomic whose address it can prove warp-uniform ptxas builds its own warp
        // aggregation, and the shuffle of that aggregation waits for the reply at once -- 2.7-4.2 % of a warp's life)
        if (A.queue && done_mask && lane == 0) qbase = atomicAdd(A.qcount + threadIdx.y, (unsigned)__popc(done_mask));
    };
    if (HULL) {
        float3 oe;
        float velv[6], ee[6], dist[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const ObstW O = env_step_begin<TASK, GEOM>(s, act, s_obs + lane * D, vel, velv, oe);
        const bool coll = hull_pass_phased<TASK, GEOM>(c_model, s_tiles, tid, O, hv, ee, dist, *s_tasks);
        env_step_finish<TASK>(s, s_obs + lane * D, ee, dist, coll, O, oe, velv, o, reserve);
        if (rows == 0) return;
    } else {
        env_step<TASK, GEOM>(c_model, s, act, hv, s_obs + lane * D, o, vel, s_scr, W, reserve);
    }
    if (live) {
        store_dyn<TASK>(A.st, i, s);
        A.rew[i] = o.reward;
        A.term[i] = o.terminated ? 1 : 0;
        A.trunc[i] = o.truncated ? 1 : 0;
        A.succ[i] = o.success ? 1 : 0;
        // Episode statistics.  With the auto-reset on they are taken by the auto-reset kernel, which has the finished envs 32
        // to a warp (one reduction and six atomics per warp there, against ~45 instructions at one or two lanes in 70 % of
        // the warps here); without it each finished env adds its own figures (no result is read back).
        unsigned long long *slot = A.stats + (size_t)((blockIdx.x * (B / W) + warp) % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT;
        if (!A.queue && (o.terminated || o.truncated)) {
            stat_add(slot + 0, 1ull);
            stat_add(slot + 1, (unsigned long long)__float2ll_rn(s.ep_ret * URGYM_RETURN_SCALE));
            stat_add(slot + 2, (unsigned long long)s.elapsed);
            if (o.success) stat_add(slot + 3, 1ull);
            if (o.collision) stat_add(slot + 4, 1ull);
            if (o.truncated && !o.terminated) stat_add(slot + 5, 1ull);
        }
        if (lane == 0) stat_add(slot + 6, (unsigned long long)rows);
    }
    __syncwarp();
    // observation tile -> global, 16-byte vectorised
    float *gobs = A.obs + wbase * D;
    if (rows == W && aligned16(gobs)) {
        float4 *g4 = reinterpret_cast<float4 *>(gobs);
        const float4 *s4 = reinterpret_cast<const float4 *>(s_obs);
#pragma unroll
        for (int k = 0; k < (W * D / 4 + W - 1) / W; k++)
            if (k * W + lane < W * D / 4) __stcs(g4 + k * W + lane, s4[k * W + lane]);
    } else {
        for (int k = lane; k < rows * D; k += W) gobs[k] = s_obs[k];
    }
    if (live) {  // achieved_goal = first G observation columns, desired_goal = columns 12..12+G of the own row
        const float *row = s_obs + lane * D;
        if (A.ach) {
            float *g = A.ach + (wbase + lane) * G;
#pragma unroll
            for (int k = 0; k < G; k++) g[k] = row[k];
        }
        if (A.des) {
            float *g = A.des + (wbase + lane) * G;
#pragma unroll
            for (int k = 0; k < G; k++) g[k] = row[12 + k];
        }
    }
    // finished envs: their rows are DummyVecEnv's terminal observation (the auto-reset kernel overwrites `obs`)
    if (done_mask && (A.tobs || A.tach)) {
        for (unsigned m = done_mask; m; m &= m - 1u) {
            const int r = __ffs(m) - 1;
            const float *row = s_obs + r * D;
            if (A.tobs) {
                float *g = A.tobs + (wbase + r) * D;
                if (lane < D) g[lane] = row[lane];
                if (lane + W < D) g[lane + W] = row[lane + W];
            }
            if (A.tach && lane < G) A.tach[(wbase + r) * G + lane] = row[lane];
        }
    }
    if (A.queue && done_mask) {
        qbase = __shfl_sync(0xffffffffu, qbase, 0);
        if ((done_mask >> lane) & 1u) A.queue[qbase + __popc(done_mask & ((1u << lane) - 1u))] = (int)(wbase + lane);
    }
}

// ------------------------------------------------------------------------------------------------ reset / auto-reset
struct AuxArgs {
    StateView st;
    int64_t n, offset;
    uint2 key;
    const uint8_t *mask, *mask2;    // reset env i when (mask ? mask[i] : 1) | (mask2 ? mask2[i] : 0); both NULL: all envs
    int autoreset;                  // 1: called after a step (keep the terminal observation, velocity from the obs row)
    float *obs, *ach, *des;
    float *tobs, *tach;             // auto-reset: terminal observation / achieved goal rows of the finished envs
    uint8_t *collision;
    unsigned long long *stats;
    const uint32_t *event;
    int chain;
    const float4 *hull;
    const int *queue;               // auto-reset kernel: the step kernel's queue of finished envs
    const uint8_t *f_term, *f_trunc, *f_succ;   // auto-reset kernel: the step's flag arrays (episode statistics are taken here)
    unsigned *qcount;               // [0] entries, [1] block tickets of the auto-reset kernel
};

#ifndef URGYM_RESET_GROUP
#define URGYM_RESET_GROUP 256       /* envs scanned by one warp of the reset kernel: ~11 finished envs at a 4 % done
                                       rate (measured on B200, Dyn 1 Mi envs: 128 -> 0.205, 256 -> 0.199, 512 -> 0.215 ms per step) */
#endif

// RobotTaskEnv.reset (core.py:263-273) of `cnt` envs listed in s_list (indices relative to gbase), by one warp, 32 at a
// time with one env per lane.  ReachDyn's rejection loop accepts only ~17.5 % of its draws on the cheap start-end
// distance rule, so that part of the search runs 8 iterations per env in parallel (8-lane groups, 4 envs per pass)
// before the lane-per-env phase evaluates the surviving iteration completely.  Returns this lane's share of the
// rejection iterations used.
template <int TASK, int GEOM>
__device__ __forceinline__ unsigned long long reset_listed(const ModelConst &c_model, const AuxArgs &A, const float4 *hv, int lane,
                                                           int64_t gbase, const int *s_list, int *s_k, float *s_rows, int cnt,
                                                           uint32_t event) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS, G = TT::GOAL, W = 32;
    // auto-reset: the rows still hold the final observation of the finished episodes -> terminal observation
    // (flattened over (row, column) so that several independent loads are in flight per lane)
    if (A.autoreset && (A.tobs || A.tach)) {
#pragma unroll 4
        for (int k = lane; k < cnt * D; k += W) {
            const int j = k / D, c = k - j * D;
            const int64_t i = gbase + s_list[j];
            const float v = A.obs[i * D + c];
            if (A.tobs) A.tobs[i * D + c] = v;
            if (A.tach && c < G) A.tach[i * G + c] = v;
        }
    }

    // ReachDyn: first iteration that passes the start-end distance rule, 4 candidate iterations per env at a time
    for (int j = lane; j < cnt; j += W) s_k[j] = 0;
    __syncwarp();
    if (TT::DYN) {
        // The 32 lanes test 32 (env, iteration) candidates per round, dealt out over the envs that are still searching: with u
        // of them left, env number r (in lane order) gets the lanes r, r + u, r + 2u, ... and tests its next iterations
        // there; the smallest passing one wins, an env without one moves its window on.  ~183 candidates are needed for
        // 32 envs (acceptance 17.5 %), so a batch takes ~8 rounds; fixed groups of 4 or 8 lanes per env took 16.2 / 12.9,
        // because every pass waited for its unluckiest env.
        for (int jb = 0; jb < cnt; jb += W) {
            const int j = jb + lane;
            const bool have = j < cnt;
            const uint64_t genv = have ? (uint64_t)(A.offset + gbase + s_list[j]) : 0ull;
            const uint32_t my_lo = (uint32_t)genv, my_hi = (uint32_t)(genv >> 32);
            bool found = !have;
            int next_k = 0, kstar = URGYM_MAX_RESET_ITERS - 1;
            for (;;) {
                const unsigned U = __ballot_sync(0xffffffffu, !found);
                if (U == 0u) break;
                const int u = __popc(U);
                const int r = lane % u, m = lane / u;
                const int src = (int)__fns(U, 0u, r + 1);               // lane of the r-th env still searching
                ResetStream rs;
                rs.key = A.key; rs.episode = event; rs.bpi = TT::BPI;
                rs.env_lo = __shfl_sync(0xffffffffu, my_lo, src); rs.env_hi = __shfl_sync(0xffffffffu, my_hi, src);
                const int it = __shfl_sync(0xffffffffu, next_k, src) + m;
                rs.iter = (uint32_t)it;
                const bool ok = it < URGYM_MAX_RESET_ITERS && dyn_pair_far_enough(rs);
                const unsigned B = __ballot_sync(0xffffffffu, ok);
                if (!found) {
                    const int mine = __popc(U & ((1u << lane) - 1u));      // this env's number among the searching ones
                    int tested = 0;
                    for (int t = mine; t < W; t += u, tested++) {
                        if ((B >> t) & 1u) { found = true; kstar = next_k + tested; break; }
                    }
                    if (!found) {
                        next_k += tested;
                        if (next_k >= URGYM_MAX_RESET_ITERS) found = true;      // give up: kstar = the last iteration
                    }
                }
            }
            if (have) s_k[j] = kstar;
        }
        __syncwarp();
    }

    // one env per lane: complete sample from iteration s_k on, neutral pose, link distances, first observation
    unsigned long long iters_total = 0ull;
    const bool take_stats = A.autoreset && A.f_term != nullptr;     // the finished episodes' figures (see the step kernel)
    long long st_ret = 0;
    unsigned st_len = 0u, st_cnt = 0u;                              // st_cnt: episodes | successes << 8 | collisions << 16 | truncations << 24
    for (int j0 = 0; j0 < cnt; j0 += W) {
        const int j = j0 + lane;
        if (j < cnt) {
            const int64_t i = gbase + s_list[j];
            const uint64_t genv = (uint64_t)(A.offset + i);
            float *row = s_rows + lane * D;
            if (take_stats) {
                const float4 tail = A.st.q8[2 * i + 1];            // q4, q5, elapsed, episode return as the step left them
                const unsigned term = A.f_term[i], trunc = A.f_trunc[i], succ = A.f_succ[i];
                st_ret += __float2ll_rn(tail.w * URGYM_RETURN_SCALE);
                st_len += (unsigned)__float_as_int(tail.z);
                st_cnt += 1u | ((succ ? 1u : 0u) << 8) | (((term && !succ) ? 1u : 0u) << 16) | (((trunc && !term) ? 1u : 0u) << 24);
            }
            EnvState s;
            if (TT::DYN) {      // ReachDyn.velocity survives the reset (quirk Q4): carry what the last episode left
                float vel[6];
                if (A.autoreset) {
#pragma unroll
                    for (int k = 0; k < 6; k++) vel[k] = A.obs[i * D + 24 + k];
                } else {
                    load_dyn<TASK>(A.st, i, s);
                    load_hot<TASK>(A.st, i, s);
                    if (s.elapsed == 0) {
                        load_vel(A.st, i, vel);
                    } else {
#pragma unroll
                        for (int k = 0; k < 6; k++) vel[k] = s.elapsed <= 25 ? s.C[8 + k] : 0.0f;
                    }
                }
                store_vel(A.st, i, vel);
#pragma unroll
                for (int k = 0; k < 6; k++) row[24 + k] = vel[k];
            }
            if (TASK == TASK_STA) load_E<TASK>(A.st, i, s.E);      // obstacle_end / obstacle_start survive a reset (reach.py:465-481)
            ResetStream rs;
            rs.key = A.key; rs.episode = event; rs.env_lo = (uint32_t)genv; rs.env_hi = (uint32_t)(genv >> 32);
            rs.bpi = TT::BPI; rs.iter = 0;
            iters_total += (unsigned long long)env_reset<TASK, GEOM>(c_model, s, rs, hv, row, s_k[j]);
            store_dyn<TASK>(A.st, i, s);
            store_E<TASK>(A.st, i, s.E);
            store_hot<TASK>(A.st, i, s);
        }
        __syncwarp();
        // new rows -> global, one row per iteration with the lanes along the row: a lane writing its own row would put
        // 32 different sectors into every store instruction (the memory pipeline throttled on exactly that)
        const int nrows = min(W, cnt - j0);
#pragma unroll 4
        for (int r = 0; r < nrows; r++) {
            const int64_t i = gbase + s_list[j0 + r];
            const float *row = s_rows + r * D;
            if (A.obs) {
                float *g = A.obs + i * D;
                if (lane < D) g[lane] = row[lane];
                if (lane + W < D) g[lane + W] = row[lane + W];
            }
            if (A.ach && lane < G) A.ach[i * G + lane] = row[lane];
            if (A.des && lane < G) A.des[i * G + lane] = row[12 + lane];
        }
        __syncwarp();
    }
    if (take_stats) {       // warp-uniform; cnt <= 32 on this path, so the packed counters cannot overflow
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_ret += __shfl_xor_sync(0xffffffffu, st_ret, o);
            st_len += __shfl_xor_sync(0xffffffffu, st_len, o);
            st_cnt += __shfl_xor_sync(0xffffffffu, st_cnt, o);
        }
        if (lane == 0) {
            unsigned long long *slot = A.stats + (size_t)(blockIdx.x % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT;
            stat_add(slot + 0, (unsigned long long)(st_cnt & 0xFFu));
            stat_add(slot + 1, (unsigned long long)st_ret);
            stat_add(slot + 2, (unsigned long long)st_len);
            if ((st_cnt >> 8) & 0xFFu) stat_add(slot + 3, (unsigned long long)((st_cnt >> 8) & 0xFFu));
            if ((st_cnt >> 16) & 0xFFu) stat_add(slot + 4, (unsigned long long)((st_cnt >> 16) & 0xFFu));
            if ((st_cnt >> 24) & 0xFFu) stat_add(slot + 5, (unsigned long long)((st_cnt >> 24) & 0xFFu));
        }
    }
    return iters_total;
}

// Reset of the envs selected by up to two byte masks: every warp scans URGYM_RESET_GROUP consecutive envs, compacts the
// selected ones with a warp scan and resets them (reset_listed).
template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_reset_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS, W = 32, NW = URGYM_BLOCK / 32;
    extern __shared__ float4 smem4[];
    float *s_rows_all = reinterpret_cast<float *>(smem4);                    // [NW][32][D] new observation rows
    int *s_list_all = reinterpret_cast<int *>(s_rows_all + NW * W * D);      // [NW][GROUP] selected envs (local index)
    int *s_k_all = s_list_all + NW * URGYM_RESET_GROUP;                      // [NW][GROUP] first iteration to evaluate
    float4 *s_hull = reinterpret_cast<float4 *>(s_k_all + NW * URGYM_RESET_GROUP);
    const float4 *hv = stage_hull<GEOM>(A.hull, s_hull);
    if (URGYM_BASE(GEOM) == GEOM_HULL) __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_rows = s_rows_all + warp * W * D;
    int *s_list = s_list_all + warp * URGYM_RESET_GROUP, *s_k = s_k_all + warp * URGYM_RESET_GROUP;
    const int64_t gbase = ((int64_t)blockIdx.x * NW + warp) * URGYM_RESET_GROUP;
    if (gbase >= A.n) return;
    const uint32_t event = A.event[A.chain];

    // compaction of the selected envs of this group: each lane reads the mask bytes of its GROUP/32 consecutive envs as
    // 32-bit words (the compiler merges them into one 16-byte load per mask array), a warp scan of the per-lane counts places the entries
    int cnt = 0;
    {
        constexpr int PER = URGYM_RESET_GROUP / W;
        const int64_t i0 = gbase + (int64_t)lane * PER;
        unsigned bits = 0;
        const bool vec_ok = (i0 + PER <= A.n) && ((gbase & 15) == 0) &&
                            (!A.mask || aligned16(A.mask)) && (!A.mask2 || aligned16(A.mask2));
        if (A.mask == nullptr && A.mask2 == nullptr) {
            for (int k = 0; k < PER; k++) if (i0 + k < A.n) bits |= 1u << k;
        } else if (vec_ok) {
            unsigned w[PER / 4];
#pragma unroll
            for (int k = 0; k < PER / 4; k++) {
                w[k] = 0u;
                if (A.mask) w[k] |= reinterpret_cast<const unsigned *>(A.mask + i0)[k];
                if (A.mask2) w[k] |= reinterpret_cast<const unsigned *>(A.mask2 + i0)[k];
            }
#pragma unroll
            for (int k = 0; k < PER; k++) if ((w[k >> 2] >> (8 * (k & 3))) & 0xFFu) bits |= 1u << k;
        } else {
            for (int k = 0; k < PER; k++) {
                const int64_t i = i0 + k;
                if (i < A.n && ((A.mask && A.mask[i]) || (A.mask2 && A.mask2[i]))) bits |= 1u << k;
            }
        }
        const int mine = __popc(bits);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < W; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        cnt = __shfl_sync(0xffffffffu, incl, W - 1);
        int pos = incl - mine;
        while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            s_list[pos++] = lane * PER + k;
        }
    }
    if (cnt == 0) return;
    __syncwarp();
    unsigned long long iters_total = reset_listed<TASK, GEOM>(c_model, A, hv, lane, gbase, s_list, s_k, s_rows, cnt, event);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) iters_total += __shfl_xor_sync(0xffffffffu, iters_total, o);
    if (lane == 0) atomicAdd(&A.stats[(blockIdx.x % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT + 7], iters_total);
}

// Auto-reset after a step: the step kernel has queued the (local) indices of the finished envs (StepArgs::queue), so the
// lanes are dense here: every warp takes 32 queue entries at a time.  The last block to finish empties the queue counter
// for the next step.
template <int TASK, int GEOM>
__global__ void __launch_bounds__(Blk<GEOM>::AUTORESET) urgym_autoreset_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    typedef Traits<TASK> TT;
    constexpr int D = TT::OBS, W = 32, NW = Blk<GEOM>::AUTORESET / 32;
    extern __shared__ float4 smem4[];
    float *s_rows_all = reinterpret_cast<float *>(smem4);                    // [NW][32][D] new observation rows
    int *s_list_all = reinterpret_cast<int *>(s_rows_all + NW * W * D);      // [NW][32]
    int *s_k_all = s_list_all + NW * W;                                      // [NW][32]
    float4 *s_hull = reinterpret_cast<float4 *>(s_k_all + NW * W);
    const float4 *hv = stage_hull<GEOM>(A.hull, s_hull);
    if (URGYM_BASE(GEOM) == GEOM_HULL) __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_rows = s_rows_all + warp * W * D;
    int *s_list = s_list_all + warp * W, *s_k = s_k_all + warp * W;
    const uint32_t event = A.event[A.chain];
    const int total = (int)*A.qcount;
    const int nwarps = gridDim.x * NW;
    unsigned long long iters_total = 0ull;
    for (int base = (blockIdx.x * NW + warp) * W; base < total; base += nwarps * W) {
        const int cnt = min(W, total - base);
        if (lane < cnt) s_list[lane] = A.queue[base + lane];
        __syncwarp();
        iters_total += reset_listed<TASK, GEOM>(c_model, A, hv, lane, 0, s_list, s_k, s_rows, cnt, event);
    }
    if ((blockIdx.x * NW + warp) * W < total) {        // warp-uniform: this warp reset something
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) iters_total += __shfl_xor_sync(0xffffffffu, iters_total, o);
        if (lane == 0) atomicAdd(&A.stats[(blockIdx.x % URGYM_STAT_SLOTS) * URGYM_STATS_COUNT + 7], iters_total);
    }
    // every block has read the counter before it takes its ticket; the last ticket holder clears both
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(A.qcount + 1, 1u);
        if (t == gridDim.x - 1) { A.qcount[0] = 0u; A.qcount[1] = 0u; }
    }
}
template <int TASK, int GEOM> constexpr size_t autoreset_smem_bytes() {
    return (size_t)Blk<GEOM>::AUTORESET * Traits<TASK>::OBS * sizeof(float) + 2 * Blk<GEOM>::AUTORESET * sizeof(int) +
           (URGYM_BASE(GEOM) == GEOM_HULL ? (size_t)URGYM_HULL_BLOB_F4 * sizeof(float4) : 0);
}
template <int TASK, int GEOM> constexpr size_t reset_smem_bytes() {
    return (size_t)URGYM_BLOCK * Traits<TASK>::OBS * sizeof(float) + 2 * (URGYM_BLOCK / 32) * URGYM_RESET_GROUP * sizeof(int) +
           (URGYM_BASE(GEOM) == GEOM_HULL ? (size_t)URGYM_HULL_BLOB_F4 * sizeof(float4) : 0);
}

// every chain's reset-event counter := max over the chains + add.  add = 1: a reset event of its own (explicit reset,
// host-buffer step); add = 0: equalise the counters (first node of a chained step graph), so that chains stepped unequally
// before can never draw from a (seed, env, event) position that another layout of the chains has already used.
static __global__ void urgym_bump_kernel(uint32_t *event, uint32_t add) {
    uint32_t m = 0u;
    for (int c = 0; c < URGYM_MAX_CHAINS; c++) m = event[c] > m ? event[c] : m;
    for (int c = 0; c < URGYM_MAX_CHAINS; c++) event[c] = m + add;
}

template <int TASK> __device__ __forceinline__ void write_rows(const AuxArgs &A, int64_t i, const float *row) {
    constexpr int D = Traits<TASK>::OBS, G = Traits<TASK>::GOAL;
    if (A.obs) for (int k = 0; k < D; k++) A.obs[i * D + k] = row[k];
    if (A.ach) for (int k = 0; k < G; k++) A.ach[i * G + k] = row[k];
    if (A.des) for (int k = 0; k < G; k++) A.des[i * G + k] = row[12 + k];
}

// (GEOM is a template parameter of the task-only kernels too: each (task, geometry) translation unit is compiled with its
// own floating-point flags, and a kernel shared between them would be one COMDAT copy built with whichever flags won)
template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_observe_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    typedef Traits<TASK> TT;
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    EnvState s;
    load_dyn<TASK>(A.st, i, s);
    load_hot<TASK>(A.st, i, s);
    float stale[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (TT::DYN) {
        load_vel(A.st, i, stale);
    }
    float row[TT::OBS];
    env_observe<TASK, GEOM_CAPSULE>(c_model, s, stale, row);
    write_rows<TASK>(A, i, row);
}

// rebuild the hot planes from the episode constants E (after urgym_set_state wrote a goal / obstacle field)
template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_derive_kernel(const AuxArgs A) {
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    EnvState s;
    load_E<TASK>(A.st, i, s.E);
    derive_cache<TASK>(s.E, s.C);
    store_hot<TASK>(A.st, i, s);
}

template <int TASK, int GEOM>
__global__ void __launch_bounds__(URGYM_BLOCK) urgym_refresh_kernel(const __grid_constant__ ModelConst c_model, const AuxArgs A) {
    extern __shared__ float4 smem4[];
    const float4 *hv = stage_hull<GEOM>(A.hull, smem4);
    if (URGYM_BASE(GEOM) == GEOM_HULL) __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    EnvState s;
    load_dyn<TASK>(A.st, i, s);
    load_hot<TASK>(A.st, i, s);
    float scratch[URGYM_SCRATCH_FLOATS];
    const bool coll = env_refresh<TASK, GEOM>(c_model, s, hv, scratch, 1);
    store_dyn<TASK>(A.st, i, s);
    if (A.collision) A.collision[i] = coll ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ launchers
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + URGYM_BLOCK - 1) / URGYM_BLOCK); }

template <int TASK, int GEOM> cudaError_t launch_step(const ModelConst &M, const StepArgs &A, cudaStream_t s) {
    constexpr int B = Blk<GEOM>::STEP;
    if (A.n <= 0 || A.n > 0x7FFFFFFFll) return cudaErrorInvalidValue;        // the kernel's env indices are 32-bit
    urgym_step_kernel<TASK, GEOM><<<(unsigned)((A.n + B - 1) / B), B, step_smem_bytes<TASK, GEOM>(), s>>>(M, A);
    return cudaGetLastError();
}
template <int TASK, int GEOM> cudaError_t launch_reset(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    const int64_t per_block = (int64_t)(URGYM_BLOCK / 32) * URGYM_RESET_GROUP;
    urgym_reset_kernel<TASK, GEOM><<<(unsigned)((A.n + per_block - 1) / per_block), URGYM_BLOCK, reset_smem_bytes<TASK, GEOM>(), s>>>(M, A);
    return cudaGetLastError();
}
template <int TASK, int GEOM> cudaError_t launch_autoreset(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    // enough warps for ~1/8 of the envs finishing in one step; beyond that the warps loop
    constexpr int AB = Blk<GEOM>::AUTORESET;
    const int64_t blocks = (A.n + 8 * AB - 1) / (8 * AB);
    // Highest launch priority: the kernel is short, latency-bound and sits on the critical path of its chain
    // (step -> auto-reset -> next step), so its blocks should not queue behind the thousands of step-kernel blocks that
    // other chains have in flight.
    static int prio_hi = 1;
    if (prio_hi == 1) { int lo = 0; if (cudaDeviceGetStreamPriorityRange(&lo, &prio_hi) != cudaSuccess) prio_hi = 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(AB);
    cfg.dynamicSmemBytes = autoreset_smem_bytes<TASK, GEOM>(); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority; attr[0].val.priority = prio_hi;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, urgym_autoreset_kernel<TASK, GEOM>, M, A);
}
template <int TASK, int GEOM> cudaError_t launch_refresh(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    const size_t smem = URGYM_BASE(GEOM) == GEOM_HULL ? (size_t)URGYM_HULL_BLOB_F4 * sizeof(float4) : 0;
    urgym_refresh_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, smem, s>>>(M, A);
    return cudaGetLastError();
}
// opt the (task, geometry) kernels in to the dynamic shared memory they need (once, at urgym_create: not legal
// while a stream is being captured into a CUDA graph)
template <int TASK, int GEOM> cudaError_t prepare_kernels(const ModelConst &, const AuxArgs &, cudaStream_t) {
    cudaError_t e = cudaFuncSetAttribute(urgym_step_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes<TASK, GEOM>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(urgym_reset_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reset_smem_bytes<TASK, GEOM>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(urgym_autoreset_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)autoreset_smem_bytes<TASK, GEOM>());
    if (e != cudaSuccess) return e;
    const size_t smem = URGYM_BASE(GEOM) == GEOM_HULL ? (size_t)URGYM_HULL_BLOB_F4 * sizeof(float4) : 0;
    if (smem) e = cudaFuncSetAttribute(urgym_refresh_kernel<TASK, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e;
}

template <int TASK, int GEOM> cudaError_t launch_observe(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    urgym_observe_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, 0, s>>>(M, A);
    return cudaGetLastError();
}

template <int TASK, int GEOM> cudaError_t launch_derive(const ModelConst &, const AuxArgs &A, cudaStream_t s) {
    urgym_derive_kernel<TASK, GEOM><<<grid_for(A.n), URGYM_BLOCK, 0, s>>>(A);
    return cudaGetLastError();
}

// one translation unit per (task, geometry) instantiates these (urgym_inst.cu), urgym_api.cu calls them by table
typedef cudaError_t (*step_launcher_t)(const ModelConst &, const StepArgs &, cudaStream_t);
typedef cudaError_t (*aux_launcher_t)(const ModelConst &, const AuxArgs &, cudaStream_t);
#define URGYM_DECLARE_INST(T, G)                                                           \
    cudaError_t urgym_inst_step_##T##_##G(const ModelConst &, const StepArgs &, cudaStream_t);   \
    cudaError_t urgym_inst_reset_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);   \
    cudaError_t urgym_inst_autoreset_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);   \
    cudaError_t urgym_inst_refresh_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);      \
    cudaError_t urgym_inst_observe_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);      \
    cudaError_t urgym_inst_derive_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);       \
    cudaError_t urgym_inst_prepare_##T##_##G(const ModelConst &, const AuxArgs &, cudaStream_t);
URGYM_DECLARE_INST(0, 0) URGYM_DECLARE_INST(1, 0) URGYM_DECLARE_INST(2, 0) URGYM_DECLARE_INST(3, 0)
URGYM_DECLARE_INST(0, 1) URGYM_DECLARE_INST(1, 1) URGYM_DECLARE_INST(2, 1) URGYM_DECLARE_INST(3, 1)
// link-distance mode "workbench" (GEOM | GEOM_WB); UR5OriReach has no link_dist
URGYM_DECLARE_INST(1, 2) URGYM_DECLARE_INST(2, 2) URGYM_DECLARE_INST(3, 2)
URGYM_DECLARE_INST(1, 3) URGYM_DECLARE_INST(2, 3) URGYM_DECLARE_INST(3, 3)
