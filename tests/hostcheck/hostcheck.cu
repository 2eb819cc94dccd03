// tests/hostcheck/hostcheck.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Instantiates the per-environment functions of ur-gym_b200/csrc/urgym_env.cuh (the __host__ __device__ code the
// CUDA kernels run, one env per thread) for the HOST, so that the CPU-only test tier (`pytest -m "not gpu"`) can
// check the product's own FP32 arithmetic against the FP64 oracle without a GPU.  Nothing in ur-gym_b200/ loads this
// library; the product has no CPU path (urgym_create fails without a CUDA device).
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../ur-gym_b200/csrc/urgym_env.cuh"
#include "../../ur-gym_b200/csrc/urgym_model.h"

using namespace urgym;

static ModelConst g_M;
static int g_ld_mode = 0;      // link-distance mode: selects the GEOM | GEOM_WB instantiations
static std::vector<float4> g_hull;
static bool g_init = false;
static void init() {
    if (g_init) return;
    build_model_const(g_M);
    g_hull.resize(URGYM_HULL_BLOB_F4);
    build_hull_blob(g_hull.data());
    g_init = true;
}

struct HcState {        // mirrors EnvState, API layout
    float *q;           // [n,6]
    int32_t *elapsed;   // [n]
    float *ep_ret;      // [n]
    float *ld;          // [n,5]
    float *E;           // [n,24] (first EW words used)
};
static void get(const HcState &S, int64_t i, EnvState &s) {
    for (int k = 0; k < 6; k++) s.q[k] = S.q[i * 6 + k];
    s.elapsed = S.elapsed[i]; s.ep_ret = S.ep_ret[i];
    for (int k = 0; k < 5; k++) s.ld[k] = S.ld[i * 5 + k];
    for (int k = 0; k < 24; k++) s.E[k] = S.E[i * 24 + k];
}
static void put(const HcState &S, int64_t i, const EnvState &s) {
    for (int k = 0; k < 6; k++) S.q[i * 6 + k] = s.q[k];
    S.elapsed[i] = s.elapsed; S.ep_ret[i] = s.ep_ret;
    for (int k = 0; k < 5; k++) S.ld[i * 5 + k] = s.ld[k];
    for (int k = 0; k < 24; k++) S.E[i * 24 + k] = s.E[k];
}

template <int TASK, int GEOM>
static void step_t(int64_t n, HcState S, const float *act, float *obs, float *rew, uint8_t *flags, float *vel) {
    constexpr int D = Traits<TASK>::OBS;
    for (int64_t i = 0; i < n; i++) {
        EnvState s; StepOut o; float v[6] = {0, 0, 0, 0, 0, 0};
        get(S, i, s);
        derive_cache<TASK>(s.E, s.C);       // what the reset / derive kernels leave in the hot planes
        float scratch[URGYM_SCRATCH_FLOATS];
        env_step<TASK, GEOM>(g_M, s, act + i * 6, g_hull.data(), obs + i * D, o, v, scratch, 1);
        put(S, i, s);
        rew[i] = o.reward;
        flags[i * 4] = o.terminated; flags[i * 4 + 1] = o.truncated; flags[i * 4 + 2] = o.success; flags[i * 4 + 3] = o.collision;
        for (int k = 0; k < 6; k++) vel[i * 6 + k] = v[k];
    }
}
template <int TASK, int GEOM>
static void reset_t(int64_t n, HcState S, uint64_t seed, uint32_t event, const int64_t *env_index, float *obs, int32_t *iters) {
    constexpr int D = Traits<TASK>::OBS;
    for (int64_t i = 0; i < n; i++) {
        EnvState s;
        get(S, i, s);           // ReachSta keeps obstacle_end / obstacle_start over a reset
        ResetStream rs;
        rs.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); rs.episode = event;
        rs.env_lo = (uint32_t)env_index[i]; rs.env_hi = (uint32_t)((uint64_t)env_index[i] >> 32);
        rs.bpi = Traits<TASK>::BPI; rs.iter = 0;
        iters[i] = env_reset<TASK, GEOM>(g_M, s, rs, g_hull.data(), obs + i * D);
        put(S, i, s);
    }
}
template <int TASK, int GEOM>
static void refresh_t(int64_t n, HcState S, uint8_t *coll) {
    for (int64_t i = 0; i < n; i++) {
        EnvState s;
        get(S, i, s);
        derive_cache<TASK>(s.E, s.C);
        float scratch[URGYM_SCRATCH_FLOATS];
        coll[i] = env_refresh<TASK, GEOM>(g_M, s, g_hull.data(), scratch, 1);
        put(S, i, s);
    }
}
template <int TASK>
static void observe_t(int64_t n, HcState S, const float *stale_vel, float *obs) {
    constexpr int D = Traits<TASK>::OBS;
    for (int64_t i = 0; i < n; i++) {
        EnvState s;
        get(S, i, s);
        derive_cache<TASK>(s.E, s.C);
        env_observe<TASK, GEOM_CAPSULE>(g_M, s, stale_vel + i * 6, obs + i * D);
    }
}

// built once per geometry (-DHC_GEOM=0 hull / 1 capsule) so that the slow-to-compile hull instantiations do not
// hold up the capsule ones; `geom` arguments must equal HC_GEOM
#ifndef HC_GEOM
#error "compile with -DHC_GEOM=0 or 1"
#endif
#define SWITCH(fn, ...)                                                                                       \
    do {                                                                                                      \
        if (geom != HC_GEOM) return -1;                                                                       \
        if (g_ld_mode && task != 0) {                                                                         \
            if (task == 1) fn<1, HC_GEOM | GEOM_WB>(__VA_ARGS__);                                             \
            else if (task == 2) fn<2, HC_GEOM | GEOM_WB>(__VA_ARGS__); else fn<3, HC_GEOM | GEOM_WB>(__VA_ARGS__); \
        } else if (task == 0) fn<0, HC_GEOM>(__VA_ARGS__); else if (task == 1) fn<1, HC_GEOM>(__VA_ARGS__);   \
        else if (task == 2) fn<2, HC_GEOM>(__VA_ARGS__); else fn<3, HC_GEOM>(__VA_ARGS__);                    \
    } while (0)

extern "C" {
int hc_step(int task, int geom, int64_t n, float *q, int32_t *elapsed, float *ep_ret, float *ld, float *E,
             const float *act, float *obs, float *rew, uint8_t *flags, float *vel) {
    init();
    HcState S = {q, elapsed, ep_ret, ld, E};
    SWITCH(step_t, n, S, act, obs, rew, flags, vel);
    return 0;
}
int hc_reset(int task, int geom, int64_t n, float *q, int32_t *elapsed, float *ep_ret, float *ld, float *E,
              uint64_t seed, uint32_t event, const int64_t *env_index, float *obs, int32_t *iters) {
    init();
    HcState S = {q, elapsed, ep_ret, ld, E};
    SWITCH(reset_t, n, S, seed, event, env_index, obs, iters);
    return 0;
}
int hc_refresh(int task, int geom, int64_t n, float *q, int32_t *elapsed, float *ep_ret, float *ld, float *E, uint8_t *coll) {
    init();
    HcState S = {q, elapsed, ep_ret, ld, E};
    SWITCH(refresh_t, n, S, coll);
    return 0;
}
void hc_observe(int task, int64_t n, float *q, int32_t *elapsed, float *ep_ret, float *ld, float *E,
                const float *stale_vel, float *obs) {
    init();
    HcState S = {q, elapsed, ep_ret, ld, E};
    if (task == 0) observe_t<0>(n, S, stale_vel, obs); else if (task == 1) observe_t<1>(n, S, stale_vel, obs);
    else if (task == 2) observe_t<2>(n, S, stale_vel, obs); else observe_t<3>(n, S, stale_vel, obs);
}
// FK only: link poses are not exported; EE pose (6) per env
void hc_ee_pose(int64_t n, const float *q, float *ee) {
    init();
    for (int64_t i = 0; i < n; i++) {
        ObstW O; O.c = f3(0, 0, 0); O.q.x = O.q.y = O.q.z = 0.0f; O.q.w = 1.0f; O.u = f3(0, 0, 1);
        float du[5];
        robot_pass<TASK_ORI, GEOM_CAPSULE>(g_M, q + i * 6, q + i * 6, O, nullptr, false, ee + i * 6, du, nullptr, 1);
    }
}
// experiments (tools/closed_loop_cpu.py): shift a scene box along z. box 0 = table, 1 = track
void hc_shift_box_z(int box, float dz) {
    init();
    g_M.box_c[box][2] += dz;
    g_M.box_k[box][0] += dz;
}
// experiments: obstacle cylinder core radius / half height / margin
void hc_set_obstacle(float r, float h, float margin) {
    init();
    g_M.obst_r = r; g_M.obst_h = h; g_M.obst_margin = margin;
}
void hc_set_ld_mode(int mode) { g_ld_mode = mode; }
void hc_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    uint4 r = philox4x32_10(make_uint4(c[0], c[1], c[2], c[3]), make_uint2(k[0], k[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
