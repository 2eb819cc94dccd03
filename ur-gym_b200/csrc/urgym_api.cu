// urgym_api.cu -- the C ABI declared in include/urgym_b200.h: handle, state pool, dispatch to the per-(task,
// geometry) kernel instantiations, state field access, statistics, host-buffer entry points.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <new>

#include "urgym_kernels.cuh"
#include "urgym_model.h"

// ------------------------------------------------------------------------------------------------ field access
// API layout (row-major [N,K], what the reference's numpy arrays look like) <-> planes
struct FieldArgs {
    StateView st;
    int64_t n;
    int field, task;
    void *ext;      // device pointer in API layout
    int to_state;   // 1: ext -> planes (set_state), 0: planes -> ext (get_state)
};

__device__ __forceinline__ float *e_word(const StateView &st, int task, int64_t i, int w) {
    (void)task;         // word w of E: 32-byte group w / 8, float4 (w % 8) / 4 of the env's pair, component w % 4
    return reinterpret_cast<float *>(&st.e8[w / 8][2 * i + (w % 8) / 4]) + (w % 4);
}

__global__ void __launch_bounds__(URGYM_BLOCK) urgym_field_kernel(const FieldArgs A) {
    const int64_t i = (int64_t)blockIdx.x * URGYM_BLOCK + threadIdx.x;
    if (i >= A.n) return;
    float *qa = reinterpret_cast<float *>(&A.st.q8[2 * i]), *qb = reinterpret_cast<float *>(&A.st.q8[2 * i + 1]);
    float *x = reinterpret_cast<float *>(A.ext);
    auto mv = [&](float *plane, float *e) { if (A.to_state) *plane = *e; else *e = *plane; };
    switch (A.field) {
    case URGYM_F_Q:
        for (int k = 0; k < 4; k++) mv(qa + k, x + i * 6 + k);
        for (int k = 0; k < 2; k++) mv(qb + k, x + i * 6 + 4 + k);
        break;
    case URGYM_F_ELAPSED: mv(qb + 2, x + i); break;            // int32 bit pattern
    case URGYM_F_EP_RETURN: mv(qb + 3, x + i); break;
    case URGYM_F_GOAL: {
        const int G = A.task == TASK_OBS ? 3 : 6;
        for (int k = 0; k < G; k++) mv(e_word(A.st, A.task, i, k), x + i * G + k);
        break;
    }
    case URGYM_F_OBSTACLE: {
        const int off = A.task == TASK_OBS ? 3 : 6;
        for (int k = 0; k < 6; k++) mv(e_word(A.st, A.task, i, off + k), x + i * 6 + k);
        break;
    }
    case URGYM_F_OBSTACLE_END:
        for (int k = 0; k < 6; k++) mv(e_word(A.st, A.task, i, 12 + k), x + i * 6 + k);
        break;
    case URGYM_F_OBSTACLE_START: {      // Sta: its own words; Dyn: obstacle_start is what URGYM_F_OBSTACLE holds
        const int off = A.task == TASK_STA ? 18 : 6;
        for (int k = 0; k < 6; k++) mv(e_word(A.st, A.task, i, off + k), x + i * 6 + k);
        break;
    }
    case URGYM_F_LINK_DIST: {
        float *l = reinterpret_cast<float *>(&A.st.ld4[i]);
        for (int k = 0; k < 4; k++) mv(l + k, x + i * 5 + k);
        mv(&A.st.ld1[i], x + i * 5 + 4);
        break;
    }
    case URGYM_F_VELOCITY: {
        float *a = reinterpret_cast<float *>(&A.st.v8[2 * i]), *b = reinterpret_cast<float *>(&A.st.v8[2 * i + 1]);
        for (int k = 0; k < 4; k++) mv(a + k, x + i * 6 + k);
        for (int k = 0; k < 2; k++) mv(b + k, x + i * 6 + 4 + k);
        break;
    }
    case URGYM_F_HOT:
        for (int k = 0; k < 24; k++)
            mv(reinterpret_cast<float *>(&A.st.h8[k / 8][2 * i + (k % 8) / 4]) + (k % 4), x + i * 24 + k);
        break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct urgym_env {
    int task, geom, device;     // geom: GEOM template value = geometry | link-distance mode << 1
    int64_t n, offset;
    uint64_t seed;
    uint32_t *d_event;      // device-resident reset-event counter: position of the counter-based reset stream.
                            // It lives on the device so that a captured CUDA graph of steps advances it on replay.
    int autoreset;
    StateView st;
    ModelConst model;
    void *pool;
    unsigned long long *stats;
    int *queue;             // auto-reset queue: local indices of the envs the last step finished (a chain uses the
                            // entries of its own env range)
    unsigned *qcount;       // per chain: [0] queue entries, [1] block tickets of the auto-reset kernel
    float4 *hull;
    int64_t launches;
    // staging for the host-buffer entry points
    cudaStream_t hstream;
    cudaStream_t cstream[3];        // chunk pipeline of the host-buffer entry point
    cudaEvent_t ev_fork, ev_join[3];
    void *dstage;
    size_t dstage_bytes;
    // asynchronous host-buffer step: two output slots, each with its own device staging and copy-out stream
    void *aslot[2];
    size_t aslot_bytes[2];
    cudaStream_t ostream[2];
    cudaEvent_t ev_chunk[2][16];     // step kernels of chunk c (slot s) done -> the copy-out stream may read them
    cudaEvent_t ev_out[2];           // copy-out of slot s complete (the next step into slot s waits for it)
    int aslot_busy[2];
    // kernel timing (urgym_profile_enable): event triples around the step and auto-reset kernels
    int profiling, prof_n;
    cudaEvent_t prof_ev[256][3];
    char err[512];
};

static char g_create_err[512] = "";

static int fail(urgym_env *h, int code, const char *fmt, const char *detail) {
    char *dst = h ? h->err : g_create_err;
    snprintf(dst, 512, fmt, detail ? detail : "");
    return code;
}
#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return fail(h, URGYM_ECUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

static int obs_dim(int task) { return task == 0 ? 18 : task == 1 ? 26 : task == 2 ? 29 : task == 3 ? 35 : -1; }
static int goal_dim(int task) { return task == 1 ? 3 : (task >= 0 && task <= 3 ? 6 : -1); }

extern "C" int urgym_obs_dim(int task) { return obs_dim(task); }
extern "C" int urgym_goal_dim(int task) { return goal_dim(task); }
extern "C" int64_t urgym_num_envs(const urgym_env_t *h) { return h ? h->n : 0; }
extern "C" const char *urgym_last_error(const urgym_env_t *h) { return h ? h->err : g_create_err; }
extern "C" int64_t urgym_launch_count(const urgym_env_t *h) { return h ? h->launches : 0; }

static inline uint2 key_of(uint64_t seed) { return make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); }

// [geom | ld_mode << 1][task]  (GEOM template value: bit 0 geometry, bit 1 link-distance mode "workbench")
static const step_launcher_t k_step[4][4] = {
    {urgym_inst_step_0_0, urgym_inst_step_1_0, urgym_inst_step_2_0, urgym_inst_step_3_0},
    {urgym_inst_step_0_1, urgym_inst_step_1_1, urgym_inst_step_2_1, urgym_inst_step_3_1},
    {nullptr, urgym_inst_step_1_2, urgym_inst_step_2_2, urgym_inst_step_3_2},
    {nullptr, urgym_inst_step_1_3, urgym_inst_step_2_3, urgym_inst_step_3_3}};
static const aux_launcher_t k_reset[4][4] = {
    {urgym_inst_reset_0_0, urgym_inst_reset_1_0, urgym_inst_reset_2_0, urgym_inst_reset_3_0},
    {urgym_inst_reset_0_1, urgym_inst_reset_1_1, urgym_inst_reset_2_1, urgym_inst_reset_3_1},
    {nullptr, urgym_inst_reset_1_2, urgym_inst_reset_2_2, urgym_inst_reset_3_2},
    {nullptr, urgym_inst_reset_1_3, urgym_inst_reset_2_3, urgym_inst_reset_3_3}};
static const aux_launcher_t k_autoreset[4][4] = {
    {urgym_inst_autoreset_0_0, urgym_inst_autoreset_1_0, urgym_inst_autoreset_2_0, urgym_inst_autoreset_3_0},
    {urgym_inst_autoreset_0_1, urgym_inst_autoreset_1_1, urgym_inst_autoreset_2_1, urgym_inst_autoreset_3_1},
    {nullptr, urgym_inst_autoreset_1_2, urgym_inst_autoreset_2_2, urgym_inst_autoreset_3_2},
    {nullptr, urgym_inst_autoreset_1_3, urgym_inst_autoreset_2_3, urgym_inst_autoreset_3_3}};
static const aux_launcher_t k_refresh[4][4] = {
    {urgym_inst_refresh_0_0, urgym_inst_refresh_1_0, urgym_inst_refresh_2_0, urgym_inst_refresh_3_0},
    {urgym_inst_refresh_0_1, urgym_inst_refresh_1_1, urgym_inst_refresh_2_1, urgym_inst_refresh_3_1},
    {nullptr, urgym_inst_refresh_1_2, urgym_inst_refresh_2_2, urgym_inst_refresh_3_2},
    {nullptr, urgym_inst_refresh_1_3, urgym_inst_refresh_2_3, urgym_inst_refresh_3_3}};
static const aux_launcher_t k_observe[4][4] = {
    {urgym_inst_observe_0_0, urgym_inst_observe_1_0, urgym_inst_observe_2_0, urgym_inst_observe_3_0},
    {urgym_inst_observe_0_1, urgym_inst_observe_1_1, urgym_inst_observe_2_1, urgym_inst_observe_3_1},
    {nullptr, urgym_inst_observe_1_2, urgym_inst_observe_2_2, urgym_inst_observe_3_2},
    {nullptr, urgym_inst_observe_1_3, urgym_inst_observe_2_3, urgym_inst_observe_3_3}};
static const aux_launcher_t k_derive[4][4] = {
    {urgym_inst_derive_0_0, urgym_inst_derive_1_0, urgym_inst_derive_2_0, urgym_inst_derive_3_0},
    {urgym_inst_derive_0_1, urgym_inst_derive_1_1, urgym_inst_derive_2_1, urgym_inst_derive_3_1},
    {nullptr, urgym_inst_derive_1_2, urgym_inst_derive_2_2, urgym_inst_derive_3_2},
    {nullptr, urgym_inst_derive_1_3, urgym_inst_derive_2_3, urgym_inst_derive_3_3}};

static const aux_launcher_t k_prepare[4][4] = {
    {urgym_inst_prepare_0_0, urgym_inst_prepare_1_0, urgym_inst_prepare_2_0, urgym_inst_prepare_3_0},
    {urgym_inst_prepare_0_1, urgym_inst_prepare_1_1, urgym_inst_prepare_2_1, urgym_inst_prepare_3_1},
    {nullptr, urgym_inst_prepare_1_2, urgym_inst_prepare_2_2, urgym_inst_prepare_3_2},
    {nullptr, urgym_inst_prepare_1_3, urgym_inst_prepare_2_3, urgym_inst_prepare_3_3}};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int urgym_create(urgym_env_t **out, int task, int geom, int64_t n_envs, int64_t env_index_offset,
                            uint64_t seed, int device) {
    urgym_env *h = nullptr;
    if (!out) return fail(nullptr, URGYM_EINVAL, "urgym_create: out is NULL%s", "");
    *out = nullptr;
    if (task < 0 || task > 3) return fail(nullptr, URGYM_EINVAL, "urgym_create: task must be 0..3%s", "");
    if (geom != URGYM_GEOM_HULL && geom != URGYM_GEOM_CAPSULE) return fail(nullptr, URGYM_EINVAL, "urgym_create: bad geom%s", "");
    if (n_envs <= 0 || env_index_offset < 0) return fail(nullptr, URGYM_EINVAL, "urgym_create: n_envs must be > 0 and offset >= 0%s", "");
    // env indices inside a handle are 32-bit (the kernels' index arithmetic; 2^31 envs would need 2 TB of state anyway); the
    // GLOBAL index offset + i of the reset stream stays 64-bit
    if (n_envs > 0x7FFFFFFFll) return fail(nullptr, URGYM_EINVAL, "urgym_create: n_envs must be < 2^31 per handle (shard over more handles)%s", "");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, URGYM_ENODEVICE, "urgym_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, URGYM_EINVAL, "urgym_create: device index out of range%s", "");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, URGYM_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, URGYM_ENODEVICE, "urgym_create: device is not sm_100 (built for sm_100a only)%s", "");
    h = new (std::nothrow) urgym_env();
    if (!h) return fail(nullptr, URGYM_ENOMEM, "urgym_create: host allocation failed%s", "");
    memset(h, 0, sizeof(*h));
    h->task = task; h->geom = geom; h->device = device; h->n = n_envs; h->offset = env_index_offset; h->seed = seed;
    h->autoreset = 1;
    int rc = URGYM_OK;
    do {
        if ((e = cudaSetDevice(device)) != cudaSuccess) { rc = URGYM_ECUDA; break; }
        build_model_const(h->model);
        // planes: 16-byte groups first, all 256-byte aligned
        const size_t n = (size_t)n_envs;
        const size_t p32 = align_up(n * 32, 256), p16 = align_up(n * 16, 256), p4 = align_up(n * 4, 256);
        const size_t total = p32 + p16 + p4 + 3 * p32 + p32 + 3 * p32 + p4 +
                             URGYM_STAT_SLOTS * URGYM_STATS_COUNT * 8 + 256 + 256;
        if ((e = cudaMalloc(&h->pool, total)) != cudaSuccess) { rc = URGYM_ENOMEM; break; }
        if ((e = cudaMemset(h->pool, 0, total)) != cudaSuccess) { rc = URGYM_ECUDA; break; }
        char *p = (char *)h->pool;
        h->st.q8 = (float4 *)p; p += p32;
        h->st.ld4 = (float4 *)p; p += p16;
        h->st.ld1 = (float *)p; p += p4;
        for (int g = 0; g < 3; g++) { h->st.e8[g] = (float4 *)p; p += p32; }
        h->st.v8 = (float4 *)p; p += p32;
        for (int g = 0; g < 3; g++) { h->st.h8[g] = (float4 *)p; p += p32; }
        h->queue = (int *)p; p += p4;
        h->stats = (unsigned long long *)p; p += URGYM_STAT_SLOTS * URGYM_STATS_COUNT * 8;
        h->d_event = (uint32_t *)p; p += 256;
        h->qcount = (unsigned *)p;
        {
            AuxArgs dummy;
            memset(&dummy, 0, sizeof(dummy));
            if ((e = k_prepare[geom][task](h->model, dummy, 0)) != cudaSuccess) { rc = URGYM_ECUDA; break; }
        }
        if ((geom & 1) == URGYM_GEOM_HULL) {
            static float4 hv[URGYM_HULL_BLOB_F4];
            build_hull_blob(hv);
            if ((e = cudaMalloc(&h->hull, sizeof(hv))) != cudaSuccess) { rc = URGYM_ENOMEM; break; }
            if ((e = cudaMemcpy(h->hull, hv, sizeof(hv), cudaMemcpyHostToDevice)) != cudaSuccess) { rc = URGYM_ECUDA; break; }
        }
        if ((e = cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking)) != cudaSuccess) { rc = URGYM_ECUDA; break; }
        for (int k = 0; k < 3 && e == cudaSuccess; k++) {
            e = cudaStreamCreateWithFlags(&h->cstream[k], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        if (e != cudaSuccess) { rc = URGYM_ECUDA; break; }
    } while (0);
    if (rc != URGYM_OK) {
        fail(nullptr, rc, "urgym_create: %s", cudaGetErrorString(e));
        if (h->hull) cudaFree(h->hull);
        if (h->pool) cudaFree(h->pool);
        delete h;
        return rc;
    }
    *out = h;
    return URGYM_OK;
}

extern "C" int urgym_destroy(urgym_env_t *h) {
    if (!h) return URGYM_EINVAL;
    cudaSetDevice(h->device);
    if (h->hstream) cudaStreamDestroy(h->hstream);
    for (int k = 0; k < 3; k++) {
        if (h->cstream[k]) cudaStreamDestroy(h->cstream[k]);
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->prof_ev[0][0])
        for (int k = 0; k < 256; k++)
            for (int j = 0; j < 3; j++) if (h->prof_ev[k][j]) cudaEventDestroy(h->prof_ev[k][j]);
    if (h->dstage) cudaFree(h->dstage);
    for (int k = 0; k < 2; k++) {
        if (h->aslot[k]) cudaFree(h->aslot[k]);
        if (h->ostream[k]) cudaStreamDestroy(h->ostream[k]);
        if (h->ev_out[k]) cudaEventDestroy(h->ev_out[k]);
        for (int c = 0; c < 16; c++) if (h->ev_chunk[k][c]) cudaEventDestroy(h->ev_chunk[k][c]);
    }
    if (h->hull) cudaFree(h->hull);
    if (h->pool) cudaFree(h->pool);
    delete h;
    return URGYM_OK;
}

extern "C" int urgym_set_autoreset(urgym_env_t *h, int enabled) {
    if (!h) return URGYM_EINVAL;
    h->autoreset = enabled ? 1 : 0;
    return URGYM_OK;
}
extern "C" int urgym_sync_events(urgym_env_t *h, void *stream) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    urgym_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->d_event, 0u);
    CK(cudaGetLastError());
    h->launches++;
    return URGYM_OK;
}
extern "C" int urgym_set_link_dist_mode(urgym_env_t *h, int mode) {
    if (!h) return URGYM_EINVAL;
    if (mode != URGYM_LD_OBSTACLE && mode != URGYM_LD_WORKBENCH) return fail(h, URGYM_EINVAL, "urgym_set_link_dist_mode: unknown mode%s", "");
    if (h->task == 0) return fail(h, URGYM_EUNSUPPORTED, "urgym_set_link_dist_mode: UR5OriReach has no link_dist%s", "");
    const int geom = (h->geom & 1) | (mode << 1);
    AuxArgs dummy;
    memset(&dummy, 0, sizeof(dummy));
    CK(cudaSetDevice(h->device));
    CK(k_prepare[geom][h->task](h->model, dummy, 0));       // the mode has its own kernel instantiations
    h->geom = geom;
    return URGYM_OK;
}
extern "C" int urgym_set_seed(urgym_env_t *h, uint64_t seed) {
    if (!h) return URGYM_EINVAL;
    h->seed = seed;
    return URGYM_OK;
}
extern "C" int urgym_get_event(const urgym_env_t *hc, uint32_t *event) {
    urgym_env *h = const_cast<urgym_env *>(hc);
    if (!h || !event) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    uint32_t all[URGYM_MAX_CHAINS];
    CK(cudaMemcpy(all, h->d_event, sizeof(all), cudaMemcpyDeviceToHost));
    uint32_t m = 0u;
    for (int c = 0; c < URGYM_MAX_CHAINS; c++) m = all[c] > m ? all[c] : m;     // the logical event: chains may lag
    *event = m;
    return URGYM_OK;
}
extern "C" int urgym_set_event(urgym_env_t *h, uint32_t event) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    uint32_t all[URGYM_MAX_CHAINS];
    for (int c = 0; c < URGYM_MAX_CHAINS; c++) all[c] = event;
    CK(cudaMemcpy(h->d_event, all, sizeof(all), cudaMemcpyHostToDevice));
    return URGYM_OK;
}

static StateView view_at(const StateView &v, int64_t off) {
    StateView o = v;
    o.q8 += 2 * off; o.ld4 += off; o.ld1 += off; o.v8 += 2 * off;
    for (int g = 0; g < 3; g++) { o.e8[g] += 2 * off; o.h8[g] += 2 * off; }
    return o;
}

// one env step of the envs [first, first + count) (array pointers are those of the WHOLE arrays)
static int step_range(urgym_env *h, int64_t first, int64_t count, int bump, int chain, int qslot, const float *actions, float *obs, float *achieved,
                      float *desired, float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                      float *terminal_obs, float *terminal_achieved, cudaStream_t stream) {
    const int D = obs_dim(h->task), G = goal_dim(h->task);
    StepArgs A;
    A.st = view_at(h->st, first); A.n = count;
    A.actions = actions + first * 6; A.obs = obs + first * D;
    A.ach = achieved ? achieved + first * G : nullptr; A.des = desired ? desired + first * G : nullptr;
    A.tobs = (h->autoreset && terminal_obs) ? terminal_obs + first * D : nullptr;
    A.tach = (h->autoreset && terminal_achieved) ? terminal_achieved + first * G : nullptr;
    A.rew = reward + first; A.term = terminated + first; A.trunc = truncated + first; A.succ = is_success + first;
    A.stats = h->stats; A.event = h->d_event; A.bump = bump; A.chain = chain; A.hull = h->hull;
    A.queue = h->autoreset ? h->queue + first : nullptr;
    A.qcount = h->qcount + 2 * qslot;     // ranges in flight at the same time use different slots
    const bool prof = h->profiling && h->prof_n < 256;
    if (prof) CK(cudaEventRecord(h->prof_ev[h->prof_n][0], stream));
    CK(k_step[h->geom][h->task](h->model, A, stream));
    h->launches++;
    if (prof) CK(cudaEventRecord(h->prof_ev[h->prof_n][1], stream));
    if (h->autoreset) {
        // the finished envs (terminated | truncated), queued by the step kernel, restart in a second, dense kernel
        AuxArgs R;
        memset(&R, 0, sizeof(R));
        R.st = A.st; R.n = count; R.offset = h->offset + first; R.key = key_of(h->seed);
        R.queue = A.queue; R.qcount = A.qcount; R.autoreset = 1;
        R.f_term = A.term; R.f_trunc = A.trunc; R.f_succ = A.succ;
        R.obs = A.obs; R.ach = A.ach; R.des = A.des;
        // (the step kernel has already written the terminal rows)
        R.stats = h->stats; R.event = h->d_event; R.chain = chain; R.hull = h->hull;
        CK(k_autoreset[h->geom][h->task](h->model, R, stream));
        h->launches++;
    }
    if (prof) { CK(cudaEventRecord(h->prof_ev[h->prof_n][2], stream)); h->prof_n++; }
    return URGYM_OK;
}

extern "C" int urgym_profile_enable(urgym_env_t *h, int enabled) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    if (enabled && !h->prof_ev[0][0])
        for (int k = 0; k < 256; k++)
            for (int j = 0; j < 3; j++) CK(cudaEventCreate(&h->prof_ev[k][j]));
    h->profiling = enabled ? 1 : 0;
    h->prof_n = 0;
    return URGYM_OK;
}
extern "C" int urgym_profile_read(urgym_env_t *h, double *step_kernel_ms, double *reset_kernel_ms, int *steps) {
    if (!h || !step_kernel_ms || !reset_kernel_ms || !steps) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    double a = 0.0, b = 0.0;
    for (int k = 0; k < h->prof_n; k++) {
        float x = 0.0f, y = 0.0f;
        CK(cudaEventElapsedTime(&x, h->prof_ev[k][0], h->prof_ev[k][1]));
        CK(cudaEventElapsedTime(&y, h->prof_ev[k][1], h->prof_ev[k][2]));
        a += x; b += y;
    }
    *steps = h->prof_n;
    *step_kernel_ms = h->prof_n ? a / h->prof_n : 0.0;
    *reset_kernel_ms = h->prof_n ? b / h->prof_n : 0.0;
    h->prof_n = 0;
    return URGYM_OK;
}

extern "C" int urgym_step_range(urgym_env_t *h, int64_t first, int64_t count, int chain, const float *actions, float *obs,
                                float *achieved, float *desired, float *reward, uint8_t *terminated, uint8_t *truncated,
                                uint8_t *is_success, float *terminal_obs, float *terminal_achieved, void *stream) {
    if (!h) return URGYM_EINVAL;
    if (!actions || !obs || !reward || !terminated || !truncated || !is_success)
        return fail(h, URGYM_EINVAL, "urgym_step_range: actions, obs, reward, terminated, truncated, is_success must not be NULL%s", "");
    if (first < 0 || count <= 0 || first + count > h->n || chain < 0 || chain >= URGYM_MAX_CHAINS)
        return fail(h, URGYM_EINVAL, "urgym_step_range: range or chain out of bounds%s", "");
    CK(cudaSetDevice(h->device));
    return step_range(h, first, count, 1, chain, chain, actions, obs, achieved, desired, reward, terminated, truncated, is_success,
                      terminal_obs, terminal_achieved, (cudaStream_t)stream);
}

extern "C" int urgym_step(urgym_env_t *h, const float *actions, float *obs, float *achieved, float *desired,
                          float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                          float *terminal_obs, float *terminal_achieved, void *stream) {
    if (!h) return URGYM_EINVAL;
    if (!actions || !obs || !reward || !terminated || !truncated || !is_success)
        return fail(h, URGYM_EINVAL, "urgym_step: actions, obs, reward, terminated, truncated, is_success must not be NULL%s", "");
    CK(cudaSetDevice(h->device));
    return step_range(h, 0, h->n, 2, 0, 0, actions, obs, achieved, desired, reward, terminated, truncated, is_success,
                      terminal_obs, terminal_achieved, (cudaStream_t)stream);
}

extern "C" int urgym_reset(urgym_env_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired, void *stream) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    AuxArgs A;
    memset(&A, 0, sizeof(A));
    A.st = h->st; A.n = h->n; A.offset = h->offset; A.key = key_of(h->seed);
    A.mask = mask; A.obs = obs; A.ach = achieved; A.des = desired; A.stats = h->stats; A.event = h->d_event; A.hull = h->hull;
    urgym_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->d_event, 1u);       // an explicit reset is a reset event of its own
    CK(cudaGetLastError());
    h->launches++;
    CK(k_reset[h->geom][h->task](h->model, A, (cudaStream_t)stream));
    h->launches++;
    return URGYM_OK;
}

extern "C" int urgym_observe(urgym_env_t *h, float *obs, float *achieved, float *desired, void *stream) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    AuxArgs A;
    memset(&A, 0, sizeof(A));
    A.st = h->st; A.n = h->n; A.offset = h->offset; A.obs = obs; A.ach = achieved; A.des = desired;
    CK(k_observe[h->geom][h->task](h->model, A, (cudaStream_t)stream));
    h->launches++;
    return URGYM_OK;
}

extern "C" int urgym_refresh(urgym_env_t *h, uint8_t *collision, void *stream) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    AuxArgs A;
    memset(&A, 0, sizeof(A));
    A.st = h->st; A.n = h->n; A.offset = h->offset; A.collision = collision; A.hull = h->hull;
    CK(k_refresh[h->geom][h->task](h->model, A, (cudaStream_t)stream));
    h->launches++;
    return URGYM_OK;
}

static int field_ok(urgym_env *h, int field) {
    if (field < 0 || field >= URGYM_F_COUNT) return 0;
    if (field == URGYM_F_OBSTACLE && h->task == 0) return 0;
    if (field == URGYM_F_LINK_DIST && h->task == 0) return 0;
    if ((field == URGYM_F_OBSTACLE_END || field == URGYM_F_OBSTACLE_START) && h->task != 2 && h->task != 3) return 0;
    if (field == URGYM_F_VELOCITY && h->task != 3) return 0;
    return 1;
}
static int field_io(urgym_env *h, int field, void *ext, int to_state, void *stream) {
    if (!h) return URGYM_EINVAL;
    if (!ext) return fail(h, URGYM_EINVAL, "urgym_get/set_state: NULL pointer%s", "");
    if (!field_ok(h, field)) return fail(h, URGYM_EUNSUPPORTED, "urgym_get/set_state: field not available for this task%s", "");
    CK(cudaSetDevice(h->device));
    FieldArgs A;
    A.st = h->st; A.n = h->n; A.field = field; A.task = h->task; A.ext = ext; A.to_state = to_state;
    urgym_field_kernel<<<grid_for(h->n), URGYM_BLOCK, 0, (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError());
    h->launches++;
    if (to_state && (field == URGYM_F_GOAL || field == URGYM_F_OBSTACLE || field == URGYM_F_OBSTACLE_END ||
                     field == URGYM_F_OBSTACLE_START)) {
        // the step reads the episode constants through the hot planes (with the episode cache): rebuild them
        AuxArgs D;
        memset(&D, 0, sizeof(D));
        D.st = h->st; D.n = h->n;
        CK(k_derive[h->geom][h->task](h->model, D, (cudaStream_t)stream));
        h->launches++;
    }
    return URGYM_OK;
}
extern "C" int urgym_get_state(urgym_env_t *h, int field, void *dst, void *stream) { return field_io(h, field, dst, 0, stream); }
extern "C" int urgym_set_state(urgym_env_t *h, int field, const void *src, void *stream) {
    return field_io(h, field, const_cast<void *>(src), 1, stream);
}

extern "C" int urgym_stats(urgym_env_t *h, double out[URGYM_STATS_COUNT], int reset_after, void *stream) {
    if (!h || !out) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    unsigned long long host[URGYM_STAT_SLOTS * URGYM_STATS_COUNT];
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(host, h->stats, sizeof(host), cudaMemcpyDeviceToHost, s));
    if (reset_after) CK(cudaMemsetAsync(h->stats, 0, sizeof(host), s));
    CK(cudaStreamSynchronize(s));
    unsigned long long sum[URGYM_STATS_COUNT] = {0};
    for (int b = 0; b < URGYM_STAT_SLOTS; b++)
        for (int k = 0; k < URGYM_STATS_COUNT; k++) sum[k] += host[b * URGYM_STATS_COUNT + k];
    for (int k = 0; k < URGYM_STATS_COUNT; k++) out[k] = (double)sum[k];
    out[1] = (double)(long long)sum[1] / (double)URGYM_RETURN_SCALE;
    return URGYM_OK;
}

// ------------------------------------------------------------------------------------------------ host-buffer entry points
static int ensure_stage(urgym_env *h, size_t bytes) {
    if (h->dstage_bytes >= bytes) return URGYM_OK;
    if (h->dstage) { cudaFree(h->dstage); h->dstage = nullptr; h->dstage_bytes = 0; }
    CK(cudaMalloc(&h->dstage, bytes));
    h->dstage_bytes = bytes;
    return URGYM_OK;
}

extern "C" int urgym_step_host(urgym_env_t *h, const float *actions, float *obs, float *achieved, float *desired,
                               float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                               float *terminal_obs, float *terminal_achieved) {
    if (!h) return URGYM_EINVAL;
    if (!actions || !obs || !reward || !terminated || !truncated || !is_success)
        return fail(h, URGYM_EINVAL, "urgym_step_host: actions, obs, reward, terminated, truncated, is_success must not be NULL%s", "");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n, D = obs_dim(h->task), G = goal_dim(h->task);
    const size_t b_act = align_up(n * 6 * 4, 256), b_obs = align_up(n * D * 4, 256), b_g = align_up(n * G * 4, 256),
                 b_r = align_up(n * 4, 256), b_f = align_up(n, 256);
    int rc = ensure_stage(h, b_act + 2 * b_obs + 3 * b_g + b_r + 3 * b_f);
    if (rc != URGYM_OK) return rc;
    char *p = (char *)h->dstage;
    float *d_act = (float *)p; p += b_act;
    float *d_obs = (float *)p; p += b_obs;
    float *d_tobs = (float *)p; p += b_obs;
    float *d_ach = (float *)p; p += b_g;
    float *d_des = (float *)p; p += b_g;
    float *d_tach = (float *)p; p += b_g;
    float *d_rew = (float *)p; p += b_r;
    uint8_t *d_term = (uint8_t *)p; p += b_f;
    uint8_t *d_trunc = (uint8_t *)p; p += b_f;
    uint8_t *d_succ = (uint8_t *)p; p += b_f;
    // Chunk pipeline over three streams: while chunk k steps, chunk k+1's actions travel host->device and chunk k-1's
    // results travel device->host (the copy engines work in both directions at once).  Chunks are multiples of the
    // reset kernel's 256-env groups; the reset event is bumped once, before the fork.
    const int64_t N = h->n;
    int64_t chunk = ((N / 8 + 255) / 256) * 256;
    if (chunk < 65536) chunk = N;                       // small batches: one chunk
    cudaStream_t s0 = h->hstream;
    urgym_bump_kernel<<<1, 1, 0, s0>>>(h->d_event, 1u);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaEventRecord(h->ev_fork, s0));
    int used = 0, k = 0;
    for (int64_t first = 0; first < N; first += chunk, k++) {
        const int64_t cnt = (N - first) < chunk ? (N - first) : chunk;
        cudaStream_t s = h->cstream[k % 3];
        if (k < 3) { CK(cudaStreamWaitEvent(s, h->ev_fork, 0)); used = k + 1; }
        CK(cudaMemcpyAsync(d_act + first * 6, actions + first * 6, (size_t)cnt * 6 * 4, cudaMemcpyHostToDevice, s));
        rc = step_range(h, first, cnt, 0, 0, k % 3, d_act, d_obs, achieved ? d_ach : nullptr, desired ? d_des : nullptr, d_rew, d_term,
                        d_trunc, d_succ, terminal_obs ? d_tobs : nullptr, terminal_achieved ? d_tach : nullptr, s);
        if (rc != URGYM_OK) return rc;
        CK(cudaMemcpyAsync(obs + first * D, d_obs + first * D, (size_t)cnt * D * 4, cudaMemcpyDeviceToHost, s));
        if (achieved) CK(cudaMemcpyAsync(achieved + first * G, d_ach + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, s));
        if (desired) CK(cudaMemcpyAsync(desired + first * G, d_des + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(reward + first, d_rew + first, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(terminated + first, d_term + first, (size_t)cnt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(truncated + first, d_trunc + first, (size_t)cnt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(is_success + first, d_succ + first, (size_t)cnt, cudaMemcpyDeviceToHost, s));
        // terminal rows are meaningful only where the env finished; the whole arrays are copied (rows of running envs
        // keep whatever the staging buffer held)
        if (terminal_obs) CK(cudaMemcpyAsync(terminal_obs + first * D, d_tobs + first * D, (size_t)cnt * D * 4, cudaMemcpyDeviceToHost, s));
        if (terminal_achieved) CK(cudaMemcpyAsync(terminal_achieved + first * G, d_tach + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, s));
    }
    for (int j = 0; j < used; j++) CK(cudaStreamSynchronize(h->cstream[j]));
    return URGYM_OK;
}


// ------------------------------------------------------------------------------------------------ asynchronous host step
// Two output slots.  Step k goes into slot k % 2: actions host -> device and the step kernels run on the chunk streams (as
// in urgym_step_host), the results leave on the slot's own copy-out stream, which waits for each chunk's kernels by event.
// The call returns without waiting; while slot s drains over PCIe the caller may already enqueue the next step into
// the other slot, whose kernels then run under this slot's device-to-host copies.
static int ensure_async(urgym_env *h, int slot, size_t bytes) {
    if (!h->ostream[slot]) {
        CK(cudaStreamCreateWithFlags(&h->ostream[slot], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_out[slot], cudaEventDisableTiming));
        for (int c = 0; c < 16; c++) CK(cudaEventCreateWithFlags(&h->ev_chunk[slot][c], cudaEventDisableTiming));
    }
    if (h->aslot_bytes[slot] >= bytes) return URGYM_OK;
    if (h->aslot[slot]) { cudaFree(h->aslot[slot]); h->aslot[slot] = nullptr; h->aslot_bytes[slot] = 0; }
    CK(cudaMalloc(&h->aslot[slot], bytes));
    h->aslot_bytes[slot] = bytes;
    return URGYM_OK;
}

extern "C" int urgym_step_host_async(urgym_env_t *h, int slot, const float *actions, float *obs, float *achieved, float *desired,
                                     float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                                     float *terminal_obs, float *terminal_achieved) {
    if (!h) return URGYM_EINVAL;
    if (slot != 0 && slot != 1) return fail(h, URGYM_EINVAL, "urgym_step_host_async: slot must be 0 or 1%s", "");
    if (!actions || !obs || !reward || !terminated || !truncated || !is_success)
        return fail(h, URGYM_EINVAL, "urgym_step_host_async: actions, obs, reward, terminated, truncated, is_success must not be NULL%s", "");
    if (h->aslot_busy[slot])
        return fail(h, URGYM_EINVAL, "urgym_step_host_async: slot still in flight, call urgym_host_wait(slot) first%s", "");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n, D = obs_dim(h->task), G = goal_dim(h->task);
    const size_t b_act = align_up(n * 6 * 4, 256), b_obs = align_up(n * D * 4, 256), b_g = align_up(n * G * 4, 256),
                 b_r = align_up(n * 4, 256), b_f = align_up(n, 256);
    int rc = ensure_async(h, slot, b_act + 2 * b_obs + 3 * b_g + b_r + 3 * b_f);
    if (rc != URGYM_OK) return rc;
    char *p = (char *)h->aslot[slot];
    float *d_act = (float *)p; p += b_act;
    float *d_obs = (float *)p; p += b_obs;
    float *d_tobs = (float *)p; p += b_obs;
    float *d_ach = (float *)p; p += b_g;
    float *d_des = (float *)p; p += b_g;
    float *d_tach = (float *)p; p += b_g;
    float *d_rew = (float *)p; p += b_r;
    uint8_t *d_term = (uint8_t *)p; p += b_f;
    uint8_t *d_trunc = (uint8_t *)p; p += b_f;
    uint8_t *d_succ = (uint8_t *)p; p += b_f;
    const int64_t N = h->n;
    int64_t chunk = ((N / 8 + 255) / 256) * 256;
    if (chunk < 65536) chunk = N;
    cudaStream_t s0 = h->hstream, so = h->ostream[slot];
    urgym_bump_kernel<<<1, 1, 0, s0>>>(h->d_event, 1u);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaEventRecord(h->ev_fork, s0));
    int k = 0;
    for (int64_t first = 0; first < N; first += chunk, k++) {
        const int64_t cnt = (N - first) < chunk ? (N - first) : chunk;
        cudaStream_t s = h->cstream[k % 3];
        if (k < 3) CK(cudaStreamWaitEvent(s, h->ev_fork, 0));
        CK(cudaMemcpyAsync(d_act + first * 6, actions + first * 6, (size_t)cnt * 6 * 4, cudaMemcpyHostToDevice, s));
        rc = step_range(h, first, cnt, 0, 0, k % 3, d_act, d_obs, achieved ? d_ach : nullptr, desired ? d_des : nullptr, d_rew, d_term,
                        d_trunc, d_succ, terminal_obs ? d_tobs : nullptr, terminal_achieved ? d_tach : nullptr, s);
        if (rc != URGYM_OK) return rc;
        CK(cudaEventRecord(h->ev_chunk[slot][k & 15], s));
        CK(cudaStreamWaitEvent(so, h->ev_chunk[slot][k & 15], 0));
        CK(cudaMemcpyAsync(obs + first * D, d_obs + first * D, (size_t)cnt * D * 4, cudaMemcpyDeviceToHost, so));
        if (achieved) CK(cudaMemcpyAsync(achieved + first * G, d_ach + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, so));
        if (desired) CK(cudaMemcpyAsync(desired + first * G, d_des + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, so));
        CK(cudaMemcpyAsync(reward + first, d_rew + first, (size_t)cnt * 4, cudaMemcpyDeviceToHost, so));
        CK(cudaMemcpyAsync(terminated + first, d_term + first, (size_t)cnt, cudaMemcpyDeviceToHost, so));
        CK(cudaMemcpyAsync(truncated + first, d_trunc + first, (size_t)cnt, cudaMemcpyDeviceToHost, so));
        CK(cudaMemcpyAsync(is_success + first, d_succ + first, (size_t)cnt, cudaMemcpyDeviceToHost, so));
        if (terminal_obs) CK(cudaMemcpyAsync(terminal_obs + first * D, d_tobs + first * D, (size_t)cnt * D * 4, cudaMemcpyDeviceToHost, so));
        if (terminal_achieved) CK(cudaMemcpyAsync(terminal_achieved + first * G, d_tach + first * G, (size_t)cnt * G * 4, cudaMemcpyDeviceToHost, so));
    }
    CK(cudaEventRecord(h->ev_out[slot], so));
    // the next step's bump (on hstream) must follow this step's kernels: join the chunk streams back into hstream
    for (int j = 0; j < 3 && j < k; j++) {
        CK(cudaEventRecord(h->ev_join[j], h->cstream[j]));
        CK(cudaStreamWaitEvent(s0, h->ev_join[j], 0));
    }
    h->aslot_busy[slot] = 1;
    return URGYM_OK;
}

extern "C" int urgym_host_wait(urgym_env_t *h, int slot) {
    if (!h) return URGYM_EINVAL;
    if (slot != 0 && slot != 1) return fail(h, URGYM_EINVAL, "urgym_host_wait: slot must be 0 or 1%s", "");
    if (!h->aslot_busy[slot]) return URGYM_OK;
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev_out[slot]));
    h->aslot_busy[slot] = 0;
    return URGYM_OK;
}

// ------------------------------------------------------------------------------------------------ replay ring
// One transition per env and step into a device-resident ring (what SB3's DictReplayBuffer.add does on the host for
// n_envs = 1: train.py:39-48 buffer_size=1e7; stable_baselines3 ReplayBuffer semantics): obs, action, reward, next_obs
// (the TERMINAL observation for envs that finished, not the first observation of the next episode), done, and the
// TimeLimit.truncated flag (handle_timeout_termination).  Row-major rings; the write cursor lives on the device so that
// a captured graph of (policy, step, replay write) advances it on every replay.  HBM-bound copy kernel: one warp moves
// whole rows with the lanes along the row.
struct ReplayArgs {
    int64_t n, capacity;
    int D;
    const float *obs, *act, *rew, *next_obs, *term_obs;
    const uint8_t *terminated, *truncated;
    float *r_obs, *r_next, *r_act, *r_rew;
    uint8_t *r_done, *r_timeout;
    unsigned long long *cursor;
};
__global__ void __launch_bounds__(256) urgym_replay_kernel(const ReplayArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned long long base = *A.cursor;
    for (int64_t i = warp; i < A.n; i += nwarps) {
        const int64_t slot = (int64_t)((base + (unsigned long long)i) % (unsigned long long)A.capacity);
        const bool done = A.terminated[i] | A.truncated[i];
        const float *nx = (done ? A.term_obs : A.next_obs) + i * A.D;
        for (int c = lane; c < A.D; c += 32) {
            A.r_obs[slot * A.D + c] = A.obs[i * A.D + c];
            A.r_next[slot * A.D + c] = nx[c];
        }
        if (lane < 6) A.r_act[slot * 6 + lane] = A.act[i * 6 + lane];
        if (lane == 6) A.r_rew[slot] = A.rew[i];
        if (lane == 7) A.r_done[slot] = done ? 1 : 0;
        if (lane == 8) A.r_timeout[slot] = (A.truncated[i] && !A.terminated[i]) ? 1 : 0;
    }
}
static __global__ void urgym_replay_advance_kernel(unsigned long long *cursor, unsigned long long n) { *cursor += n; }

extern "C" int urgym_replay_write(urgym_env_t *h, const float *obs, const float *actions, const float *reward,
                                  const uint8_t *terminated, const uint8_t *truncated, const float *next_obs,
                                  const float *terminal_obs, float *ring_obs, float *ring_next_obs, float *ring_actions,
                                  float *ring_reward, uint8_t *ring_done, uint8_t *ring_timeout, int64_t capacity,
                                  unsigned long long *cursor, void *stream) {
    if (!h) return URGYM_EINVAL;
    if (!obs || !actions || !reward || !terminated || !truncated || !next_obs || !terminal_obs || !ring_obs || !ring_next_obs ||
        !ring_actions || !ring_reward || !ring_done || !ring_timeout || !cursor)
        return fail(h, URGYM_EINVAL, "urgym_replay_write: NULL pointer%s", "");
    if (capacity < h->n) return fail(h, URGYM_EINVAL, "urgym_replay_write: capacity must be at least the number of envs%s", "");
    CK(cudaSetDevice(h->device));
    ReplayArgs A;
    A.n = h->n; A.capacity = capacity; A.D = obs_dim(h->task);
    A.obs = obs; A.act = actions; A.rew = reward; A.next_obs = next_obs; A.term_obs = terminal_obs;
    A.terminated = terminated; A.truncated = truncated;
    A.r_obs = ring_obs; A.r_next = ring_next_obs; A.r_act = ring_actions; A.r_rew = ring_reward;
    A.r_done = ring_done; A.r_timeout = ring_timeout; A.cursor = cursor;
    const int64_t warps = (h->n + 3) / 4;                   // four rows per warp
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    urgym_replay_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError());
    urgym_replay_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(cursor, (unsigned long long)h->n);
    CK(cudaGetLastError());
    h->launches += 2;
    return URGYM_OK;
}

extern "C" int urgym_reset_host(urgym_env_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired) {
    if (!h) return URGYM_EINVAL;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n, D = obs_dim(h->task), G = goal_dim(h->task);
    const size_t b_obs = align_up(n * D * 4, 256), b_g = align_up(n * G * 4, 256), b_f = align_up(n, 256);
    int rc = ensure_stage(h, b_obs + 2 * b_g + b_f);
    if (rc != URGYM_OK) return rc;
    char *p = (char *)h->dstage;
    float *d_obs = (float *)p; p += b_obs;
    float *d_ach = (float *)p; p += b_g;
    float *d_des = (float *)p; p += b_g;
    uint8_t *d_mask = (uint8_t *)p;
    cudaStream_t s = h->hstream;
    if (mask) CK(cudaMemcpyAsync(d_mask, mask, n, cudaMemcpyHostToDevice, s));
    if (mask && obs) {      // rows that are not reset must keep the caller's contents
        CK(cudaMemcpyAsync(d_obs, obs, n * D * 4, cudaMemcpyHostToDevice, s));
        if (achieved) CK(cudaMemcpyAsync(d_ach, achieved, n * G * 4, cudaMemcpyHostToDevice, s));
        if (desired) CK(cudaMemcpyAsync(d_des, desired, n * G * 4, cudaMemcpyHostToDevice, s));
    }
    rc = urgym_reset(h, mask ? d_mask : nullptr, obs ? d_obs : nullptr, achieved ? d_ach : nullptr, desired ? d_des : nullptr, s);
    if (rc != URGYM_OK) return rc;
    if (obs) CK(cudaMemcpyAsync(obs, d_obs, n * D * 4, cudaMemcpyDeviceToHost, s));
    if (achieved) CK(cudaMemcpyAsync(achieved, d_ach, n * G * 4, cudaMemcpyDeviceToHost, s));
    if (desired) CK(cudaMemcpyAsync(desired, d_des, n * G * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return URGYM_OK;
}
