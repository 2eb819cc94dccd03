// urgym_motor.cu -- the reference's MOTOR-DRIVEN robot path, batched: `UR5IAIReach-v1`
// (UR_gym/envs/ur_tasks.py:10-21: robot `UR5`, task `ReachIAI`), one env per thread, FP32, sm_100a.
//
// Reference semantics followed (file:line relative to the UR-gym repository):
//   RobotTaskEnv.step / reset / _get_obs   UR_gym/envs/core.py:252-273,303-317; TimeLimit(100) UR_gym/__init__.py:7-11
//   UR5.set_action                         UR_gym/envs/robots/UR5.py:44-50,76-90   target = joint angles + 0.1 pi clip(a)
//   PyBullet.control_joints                UR_gym/pyb_setup.py:365-380   setJointMotorControlArray(POSITION_CONTROL, forces)
//   PyBullet.step                          UR_gym/pyb_setup.py:52-55     20 x stepSimulation, dt = 1/500, gravity -9.81
//   UR5.get_obs                            UR5.py:92-97                  ee position + ee linear velocity (link 6)
//   ReachIAI                               UR_gym/envs/tasks/reach.py:9-66   goal box, success d < 0.005, reward -d
// What Bullet does inside those calls is restated from its published algorithm, unverified against a PyBullet run
// (oracle/ur_motor_oracle.c lists every [RECALLED] item): per substep forward dynamics without motor torques (gravity,
// velocity products, link damping), the six motor rows solved at velocity level by 50 sweeps of projected Gauss-Seidel with
// the impulse clamped to force * dt, semi-implicit Euler.  Not modelled: joint-limit rows, contact rows.
//
// Formulation here: recursive Newton-Euler in link frames (the bias forces, and the mass matrix column by column); the
// oracle assembles both from world-frame Jacobians.  State per env: q[6], qd[6], goal[3], elapsed, episode return (SoA planes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/urgym_b200.h"
#include "ur5_motor_data.h"
#include "urgym_device.cuh"

using namespace urgym;

namespace {

struct MotorConst {
    float xyz[7][3], rot[7][9], axis[6][3], mass[7], com[7][3], inertia[7][3], force[6];
    float neutral_q[6], neutral_ee[3];
};

#define MOTOR_SUBSTEPS 20          /* pyb_setup.py:24,50 */
#define MOTOR_DT (1.0f / 500.0f)   /* pyb_setup.py:25 */
#define MOTOR_KP 0.1f              /* pybullet POSITION_CONTROL defaults [RECALLED] */
#define MOTOR_KD 1.0f
#define MOTOR_ITERS 50             /* numSolverIterations [RECALLED] */
#define MOTOR_GRAVITY 9.81f        /* pyb_setup.py:44 */
#define MOTOR_DAMP 0.04f           /* default linear / angular damping of a multibody link [RECALLED] */

struct Rot3 { float m[9]; };
__device__ __forceinline__ float3 mulT(const Rot3 &R, float3 v) {    // R^T v
    return f3(fmaf(R.m[0], v.x, fmaf(R.m[3], v.y, R.m[6] * v.z)), fmaf(R.m[1], v.x, fmaf(R.m[4], v.y, R.m[7] * v.z)),
              fmaf(R.m[2], v.x, fmaf(R.m[5], v.y, R.m[8] * v.z)));
}
__device__ __forceinline__ float3 mul(const Rot3 &R, float3 v) { return rot(R.m, v); }

// rotation child -> parent of joint i at angle q:  Rorig_i * Rot(axis_i, q)   (Rodrigues; the axes are unit vectors)
__device__ __forceinline__ Rot3 joint_rot(const MotorConst &M, int i, float q) {
    Rot3 R;
    if (i >= 6) {
#pragma unroll
        for (int k = 0; k < 9; k++) R.m[k] = M.rot[i][k];
        return R;
    }
    float s, c;
    sincosf(q, &s, &c);
    const float ax = M.axis[i][0], ay = M.axis[i][1], az = M.axis[i][2], t = 1.0f - c;
    const float Q[9] = {t * ax * ax + c, t * ax * ay - s * az, t * ax * az + s * ay,
                        t * ax * ay + s * az, t * ay * ay + c, t * ay * az - s * ax,
                        t * ax * az - s * ay, t * ay * az + s * ax, t * az * az + c};
    const float *O = M.rot[i];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) R.m[3 * r + k] = fmaf(O[3 * r], Q[k], fmaf(O[3 * r + 1], Q[3 + k], O[3 * r + 2] * Q[6 + k]));
    return R;
}

// Recursive Newton-Euler over the 7 bodies (links 1..6, ee_link fixed to link 6), link frames.
//   FULL = true : tau = bias forces at (q, qd): gravity, velocity products, link damping (qdd = 0)
//   FULL = false: tau = M(q) e_col (unit acceleration of joint `col`, no velocities, no gravity)
template <bool FULL>
__device__ __forceinline__ void rnea(const MotorConst &M, const Rot3 *R, const float *qd, int col, float *tau) {
    float3 F[7], N[7];
    float3 wp = f3(0, 0, 0), alp = f3(0, 0, 0), aop = FULL ? f3(0, 0, MOTOR_GRAVITY) : f3(0, 0, 0), vop = f3(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const float3 p = f3(M.xyz[i][0], M.xyz[i][1], M.xyz[i][2]), c = f3(M.com[i][0], M.com[i][1], M.com[i][2]);
        const float3 I = f3(M.inertia[i][0], M.inertia[i][1], M.inertia[i][2]);
        const float3 a = i < 6 ? f3(M.axis[i][0], M.axis[i][1], M.axis[i][2]) : f3(0, 0, 0);
        float3 wi = mulT(R[i], wp), ali = mulT(R[i], alp);
        const float3 ao = mulT(R[i], aop + cross(alp, p) + (FULL ? cross(wp, cross(wp, p)) : f3(0, 0, 0)));
        float3 vo = f3(0, 0, 0);
        if (FULL) {
            vo = mulT(R[i], vop + cross(wp, p));
            if (i < 6) {
                const float3 rel = qd[i] * a;
                ali = ali + cross(wi, rel);
                wi = wi + rel;
            }
        } else if (i < 6 && i == col) {
            ali = ali + a;
        }
        const float3 ac = ao + cross(ali, c) + (FULL ? cross(wi, cross(wi, c)) : f3(0, 0, 0));
        const float m = M.mass[i];
        F[i] = m * ac;
        N[i] = f3(I.x * ali.x, I.y * ali.y, I.z * ali.z);
        if (FULL) {
            const float3 Iw = f3(I.x * wi.x, I.y * wi.y, I.z * wi.z);
            const float3 vc = vo + cross(wi, c);
            N[i] = N[i] + cross(wi, Iw);
            F[i] = F[i] + (m * (MOTOR_DAMP + MOTOR_DAMP * sqrtf(dot(vc, vc)))) * vc;      // link damping [RECALLED]
            N[i] = N[i] + (MOTOR_DAMP + MOTOR_DAMP * sqrtf(dot(wi, wi))) * Iw;
        }
        wp = wi; alp = ali; aop = ao; vop = vo;
    }
    float3 f = f3(0, 0, 0), n = f3(0, 0, 0);
#pragma unroll
    for (int i = 6; i >= 0; i--) {
        const float3 c = f3(M.com[i][0], M.com[i][1], M.com[i][2]);
        // wrench of the subtree at link i's origin, link-i frame
        f = F[i] + f;
        n = N[i] + cross(c, F[i]) + n;
        if (i < 6) tau[i] = M.axis[i][0] * n.x + M.axis[i][1] * n.y + M.axis[i][2] * n.z;
        // carry to the parent frame: rotate, shift by the joint origin
        const float3 p = f3(M.xyz[i][0], M.xyz[i][1], M.xyz[i][2]);
        const float3 fp = mul(R[i], f), np = mul(R[i], n);
        n = np + cross(p, fp);
        f = fp;
    }
}

// 6 x 6 SPD inverse (Gauss-Jordan without pivoting: the mass matrix is positive definite, condition ~100)
__device__ __forceinline__ void invert6(float (&A)[6][6], float (&B)[6][6]) {
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) B[r][c] = r == c ? 1.0f : 0.0f;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const float inv = 1.0f / A[c][c];
#pragma unroll
        for (int k = 0; k < 6; k++) { A[c][k] *= inv; B[c][k] *= inv; }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            if (r != c) {
                const float fct = A[r][c];
#pragma unroll
                for (int k = 0; k < 6; k++) { A[r][k] = fmaf(-fct, A[c][k], A[r][k]); B[r][k] = fmaf(-fct, B[c][k], B[r][k]); }
            }
        }
    }
}

// control_joints + sim.step(): 20 substeps with position motors towards `target`
__device__ __forceinline__ void motor_substeps(const MotorConst &M, float *q, float *qd, const float *target) {
#pragma unroll 1
    for (int s = 0; s < MOTOR_SUBSTEPS; s++) {
        Rot3 R[7];
#pragma unroll
        for (int i = 0; i < 7; i++) R[i] = joint_rot(M, i, i < 6 ? q[i] : 0.0f);
        float A[6][6], Minv[6][6], tau[6], v[6];
#pragma unroll 1
        for (int col = 0; col < 6; col++) {
            float t[6];
            rnea<false>(M, R, qd, col, t);
#pragma unroll
            for (int r = 0; r < 6; r++) {       // column `col` (static row index, dynamic column: a select per entry)
#pragma unroll
                for (int k = 0; k < 6; k++) if (k == col) A[r][k] = t[r];
            }
        }
        invert6(A, Minv);
        rnea<true>(M, R, qd, -1, tau);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            float a = 0.0f;
#pragma unroll
            for (int k = 0; k < 6; k++) a = fmaf(-Minv[i][k], tau[k], a);
            v[i] = fmaf(MOTOR_DT, a, qd[i]);
        }
        float want[6], lam[6], idiag[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            want[i] = MOTOR_KP * (target[i] - q[i]) / MOTOR_DT + v[i] + MOTOR_KD * (0.0f - v[i]);
            lam[i] = 0.0f;
            idiag[i] = 1.0f / Minv[i][i];       // effective mass of the row
        }
#pragma unroll 1
        for (int it = 0; it < MOTOR_ITERS; it++) {
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const float lim = M.force[i] * MOTOR_DT;
                float l1 = fmaf(want[i] - v[i], idiag[i], lam[i]);
                l1 = fminf(fmaxf(l1, -lim), lim);
                const float dl = l1 - lam[i];
                lam[i] = l1;
#pragma unroll
                for (int k = 0; k < 6; k++) v[k] = fmaf(Minv[k][i], dl, v[k]);
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { qd[i] = v[i]; q[i] = fmaf(MOTOR_DT, v[i], q[i]); }
    }
}

// ee_link world position and linear velocity (getLinkState(6)[0], [6])
__device__ __forceinline__ void ee_state(const MotorConst &M, const float *q, const float *qd, float *pos, float *vel) {
    float Rw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float3 pw = f3(0, 0, 0), w = f3(0, 0, 0), vp = f3(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const float3 d = rot(Rw, f3(M.xyz[i][0], M.xyz[i][1], M.xyz[i][2]));
        pw = pw + d;
        vp = vp + cross(w, d);
        const Rot3 Rj = joint_rot(M, i, i < 6 ? q[i] : 0.0f);
        float Rn[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) Rn[3 * r + k] = fmaf(Rw[3 * r], Rj.m[k], fmaf(Rw[3 * r + 1], Rj.m[3 + k], Rw[3 * r + 2] * Rj.m[6 + k]));
#pragma unroll
        for (int k = 0; k < 9; k++) Rw[k] = Rn[k];
        if (i < 6) w = w + qd[i] * rot(Rw, f3(M.axis[i][0], M.axis[i][1], M.axis[i][2]));
    }
    pos[0] = pw.x; pos[1] = pw.y; pos[2] = pw.z; vel[0] = vp.x; vel[1] = vp.y; vel[2] = vp.z;
}

struct MotorState {       // structure-of-arrays planes, one cudaMalloc
    float *q[6], *qd[6], *goal[3], *ep_ret;
    int *elapsed;
};
struct MotorArgs {
    MotorState st;
    int64_t n, offset;
    uint2 key;
    const float *actions;
    float *obs, *ach, *des, *rew, *tobs;
    uint8_t *term, *trunc, *succ;
    const uint8_t *mask;
    uint32_t *event;
    unsigned long long *stats;
};

// ReachIAI._sample_goal (reach.py:54-57): one Philox block per reset, slots 0..2
__device__ __forceinline__ void sample_goal(const MotorArgs &A, int64_t i, uint32_t event, float *goal) {
    const uint64_t genv = (uint64_t)(A.offset + i);
    const uint4 r = philox4x32_10(make_uint4(0u, event, (uint32_t)genv, (uint32_t)(genv >> 32)), A.key);
    goal[0] = 0.2f + (0.6f - 0.2f) * u01(r.x);         // goal_range_low / high, reach.py:20-21
    goal[1] = -0.4f + (0.4f - -0.4f) * u01(r.y);
    goal[2] = 0.0f + (0.8f - 0.0f) * u01(r.z);
}
__device__ __forceinline__ void write_rows(const MotorArgs &A, int64_t i, const float *pos, const float *vel, const float *goal) {
    if (A.obs) {
#pragma unroll
        for (int k = 0; k < 3; k++) { A.obs[i * 6 + k] = pos[k]; A.obs[i * 6 + 3 + k] = vel[k]; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (A.ach) A.ach[i * 3 + k] = pos[k];
        if (A.des) A.des[i * 3 + k] = goal[k];
    }
}
__device__ __forceinline__ void reset_env(const MotorConst &M, const MotorArgs &A, int64_t i, uint32_t event) {
    float goal[3];
    sample_goal(A, i, event, goal);
#pragma unroll
    for (int k = 0; k < 6; k++) { A.st.q[k][i] = M.neutral_q[k]; A.st.qd[k][i] = 0.0f; }     // set_joint_neutral: resetJointState
#pragma unroll
    for (int k = 0; k < 3; k++) A.st.goal[k][i] = goal[k];
    A.st.elapsed[i] = 0;
    A.st.ep_ret[i] = 0.0f;
    const float zero[3] = {0.0f, 0.0f, 0.0f};
    write_rows(A, i, M.neutral_ee, zero, goal);
}

__global__ void __launch_bounds__(128) urgym_motor_reset_kernel(const __grid_constant__ MotorConst M, const MotorArgs A) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= A.n) return;
    if (A.mask && !A.mask[i]) return;
    reset_env(M, A, i, A.event[0]);
}
__global__ void urgym_motor_bump_kernel(uint32_t *event) { event[0] += 1u; }

// RobotTaskEnv.step (core.py:303-317) + TimeLimit + auto-reset of the finished envs (DummyVecEnv semantics)
__global__ void __launch_bounds__(128) urgym_motor_step_kernel(const __grid_constant__ MotorConst M, const MotorArgs A) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= A.n) return;
    if (threadIdx.x == 0) atomicAdd(A.stats + 6, (unsigned long long)min((int64_t)128, A.n - i));      // env steps
    float q[6], qd[6], target[6], goal[3];
#pragma unroll
    for (int k = 0; k < 6; k++) { q[k] = A.st.q[k][i]; qd[k] = A.st.qd[k][i]; }
#pragma unroll
    for (int k = 0; k < 3; k++) goal[k] = A.st.goal[k][i];
    // UR5.set_action: clip, * pi, * 0.1 in float32, added to the current joint angles          UR5.py:44-50,76-90
#pragma unroll
    for (int k = 0; k < 6; k++) target[k] = q[k] + (clampf(A.actions[i * 6 + k], -1.0f, 1.0f) * URGYM_PI_F) * 0.1f;
    motor_substeps(M, q, qd, target);
    float pos[3], vel[3];
    ee_state(M, q, qd, pos, vel);
    // ReachIAI: success d < 0.005; check_collision() returns None, so terminated == success      reach.py:59-68, core.py:310-315
    const float dx = pos[0] - goal[0], dy = pos[1] - goal[1], dz = pos[2] - goal[2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    const bool succ = d < 0.005f;
    const float r = -d;
    const int elapsed = A.st.elapsed[i] + 1;
    const bool trunc = elapsed >= URGYM_MAX_EPISODE_STEPS;
    const float ep_ret = A.st.ep_ret[i] + r;
    A.rew[i] = r;
    A.term[i] = succ ? 1 : 0;
    A.trunc[i] = trunc ? 1 : 0;
    A.succ[i] = succ ? 1 : 0;
    if (succ || trunc) {
        if (A.tobs) {
#pragma unroll
            for (int k = 0; k < 3; k++) { A.tobs[i * 6 + k] = pos[k]; A.tobs[i * 6 + 3 + k] = vel[k]; }
        }
        atomicAdd(A.stats + 0, 1ull);
        atomicAdd(A.stats + 1, (unsigned long long)__float2ll_rn(ep_ret * 65536.0f));
        atomicAdd(A.stats + 2, (unsigned long long)elapsed);
        if (succ) atomicAdd(A.stats + 3, 1ull);
        if (trunc && !succ) atomicAdd(A.stats + 5, 1ull);
        reset_env(M, A, i, A.event[0]);
        return;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { A.st.q[k][i] = q[k]; A.st.qd[k][i] = qd[k]; }
    A.st.elapsed[i] = elapsed;
    A.st.ep_ret[i] = ep_ret;
    write_rows(A, i, pos, vel, goal);
}

__global__ void __launch_bounds__(128) urgym_motor_field_kernel(MotorState st, int64_t n, int field, float *buf, int set) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    if (field == URGYM_MOTOR_F_ELAPSED) {
        int *b = reinterpret_cast<int *>(buf);
        if (set) st.elapsed[i] = b[i]; else b[i] = st.elapsed[i];
        return;
    }
    float **p = field == URGYM_MOTOR_F_Q ? st.q : (field == URGYM_MOTOR_F_QD ? st.qd : st.goal);
    const int w = field == URGYM_MOTOR_F_GOAL ? 3 : 6;
    for (int k = 0; k < w; k++) {
        if (set) p[k][i] = buf[i * w + k]; else buf[i * w + k] = p[k][i];
    }
}

void build_motor_const(MotorConst &M) {
    memset(&M, 0, sizeof(M));
    for (int i = 0; i < 7; i++) {
        for (int k = 0; k < 3; k++) {
            M.xyz[i][k] = (float)UR5M_JOINT_XYZ[3 * i + k]; M.com[i][k] = (float)UR5M_LINK_COM[3 * i + k];
            M.inertia[i][k] = (float)UR5M_LINK_INERTIA[3 * i + k];
        }
        for (int k = 0; k < 9; k++) M.rot[i][k] = (float)UR5M_JOINT_ROT[9 * i + k];
        M.mass[i] = (float)UR5M_LINK_MASS[i];
    }
    for (int i = 0; i < 6; i++) {
        for (int k = 0; k < 3; k++) M.axis[i][k] = (float)UR5M_JOINT_AXIS[3 * i + k];
        M.force[i] = (float)UR5M_JOINT_EFFORT[i];                  // = UR5.joint_forces, UR5.py:36
    }
    const double qn[6] = {0.0, -1.5708, 0.0, 0.0, 0.0, 0.0};       // UR5.py:39
    // ee position at the reset pose, in double on the host
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
    for (int i = 0; i < 7; i++) {
        const double *X = &UR5M_JOINT_XYZ[3 * i], *O = &UR5M_JOINT_ROT[9 * i];
        for (int r = 0; r < 3; r++) p[r] += R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2];
        double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (i < 6) {
            const double *a = &UR5M_JOINT_AXIS[3 * i], c = cos(qn[i]), s = sin(qn[i]), t = 1.0 - c;
            const double Qa[9] = {t * a[0] * a[0] + c, t * a[0] * a[1] - s * a[2], t * a[0] * a[2] + s * a[1],
                                  t * a[0] * a[1] + s * a[2], t * a[1] * a[1] + c, t * a[1] * a[2] - s * a[0],
                                  t * a[0] * a[2] - s * a[1], t * a[1] * a[2] + s * a[0], t * a[2] * a[2] + c};
            memcpy(Q, Qa, sizeof(Q));
        }
        double A[9], B[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) A[3 * r + c] = R[3 * r] * O[c] + R[3 * r + 1] * O[3 + c] + R[3 * r + 2] * O[6 + c];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) B[3 * r + c] = A[3 * r] * Q[c] + A[3 * r + 1] * Q[3 + c] + A[3 * r + 2] * Q[6 + c];
        memcpy(R, B, sizeof(R));
    }
    for (int k = 0; k < 6; k++) M.neutral_q[k] = (float)qn[k];
    for (int k = 0; k < 3; k++) M.neutral_ee[k] = (float)p[k];
}

}  // namespace

struct urgym_motor {
    int device;
    int64_t n, offset;
    uint64_t seed;
    MotorConst model;
    MotorState st;
    void *pool;
    uint32_t *d_event;
    unsigned long long *d_stats;
    long long launches;
    char err[256];
};
static char g_motor_err[256] = "";
#define MCK(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            snprintf(h->err, sizeof(h->err), "%s failed: %s", #call, cudaGetErrorString(e_));           \
            return URGYM_ECUDA;                                                                         \
        }                                                                                               \
    } while (0)

extern "C" int urgym_motor_create(urgym_motor_t **out, int64_t n_envs, int64_t env_index_offset, uint64_t seed, int device) {
    if (!out || n_envs <= 0) { snprintf(g_motor_err, sizeof(g_motor_err), "urgym_motor_create: bad argument"); return URGYM_EINVAL; }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        snprintf(g_motor_err, sizeof(g_motor_err), "urgym_motor_create: no usable CUDA device %d (this library has no CPU path)", device);
        return URGYM_ENODEVICE;
    }
    urgym_motor *h = new urgym_motor();
    memset(h, 0, sizeof(*h));
    h->device = device; h->n = n_envs; h->offset = env_index_offset; h->seed = seed;
    build_motor_const(h->model);
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return URGYM_ECUDA; }
    const size_t plane = ((size_t)n_envs * 4 + 255) / 256 * 256;
    const size_t total = 17 * plane + 256 + 256;
    if (cudaMalloc(&h->pool, total) != cudaSuccess) { snprintf(g_motor_err, sizeof(g_motor_err), "urgym_motor_create: cudaMalloc of %zu bytes failed", total); delete h; return URGYM_ENOMEM; }
    cudaMemset(h->pool, 0, total);
    char *p = (char *)h->pool;
    for (int k = 0; k < 6; k++) { h->st.q[k] = (float *)p; p += plane; }
    for (int k = 0; k < 6; k++) { h->st.qd[k] = (float *)p; p += plane; }
    for (int k = 0; k < 3; k++) { h->st.goal[k] = (float *)p; p += plane; }
    h->st.ep_ret = (float *)p; p += plane;
    h->st.elapsed = (int *)p; p += plane;
    h->d_event = (uint32_t *)p; p += 256;
    h->d_stats = (unsigned long long *)p;
    *out = h;
    return URGYM_OK;
}
extern "C" int urgym_motor_destroy(urgym_motor_t *h) {
    if (!h) return URGYM_EINVAL;
    cudaSetDevice(h->device);
    cudaFree(h->pool);
    delete h;
    return URGYM_OK;
}
extern "C" const char *urgym_motor_last_error(const urgym_motor_t *h) { return h ? h->err : g_motor_err; }

static MotorArgs motor_args(urgym_motor *h) {
    MotorArgs A;
    memset(&A, 0, sizeof(A));
    A.st = h->st; A.n = h->n; A.offset = h->offset;
    A.key = make_uint2((uint32_t)h->seed, (uint32_t)(h->seed >> 32));
    A.event = h->d_event; A.stats = h->d_stats;
    return A;
}
extern "C" int urgym_motor_reset(urgym_motor_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired, void *stream) {
    if (!h) return URGYM_EINVAL;
    MCK(cudaSetDevice(h->device));
    MotorArgs A = motor_args(h);
    A.mask = mask; A.obs = obs; A.ach = achieved; A.des = desired;
    cudaStream_t s = (cudaStream_t)stream;
    urgym_motor_bump_kernel<<<1, 1, 0, s>>>(h->d_event);
    urgym_motor_reset_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, s>>>(h->model, A);
    MCK(cudaGetLastError());
    h->launches += 2;
    return URGYM_OK;
}
extern "C" int urgym_motor_step(urgym_motor_t *h, const float *actions, float *obs, float *achieved, float *desired, float *reward,
                                uint8_t *terminated, uint8_t *truncated, uint8_t *is_success, float *terminal_obs, void *stream) {
    if (!h || !actions || !reward || !terminated || !truncated || !is_success) {
        if (h) snprintf(h->err, sizeof(h->err), "urgym_motor_step: actions, reward and the three flag arrays are required");
        return URGYM_EINVAL;
    }
    MCK(cudaSetDevice(h->device));
    MotorArgs A = motor_args(h);
    A.actions = actions; A.obs = obs; A.ach = achieved; A.des = desired; A.rew = reward; A.tobs = terminal_obs;
    A.term = terminated; A.trunc = truncated; A.succ = is_success;
    cudaStream_t s = (cudaStream_t)stream;
    urgym_motor_bump_kernel<<<1, 1, 0, s>>>(h->d_event);
    urgym_motor_step_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, s>>>(h->model, A);
    MCK(cudaGetLastError());
    h->launches += 2;
    return URGYM_OK;
}
static int motor_field(urgym_motor *h, int field, float *buf, int set, void *stream) {
    if (!h || !buf || field < 0 || field >= URGYM_MOTOR_F_COUNT) {
        if (h) snprintf(h->err, sizeof(h->err), "urgym_motor_get/set_state: bad field %d", field);
        return URGYM_EINVAL;
    }
    MCK(cudaSetDevice(h->device));
    urgym_motor_field_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->st, h->n, field, buf, set);
    MCK(cudaGetLastError());
    h->launches++;
    return URGYM_OK;
}
extern "C" int urgym_motor_get_state(urgym_motor_t *h, int field, void *dst, void *stream) { return motor_field(h, field, (float *)dst, 0, stream); }
extern "C" int urgym_motor_set_state(urgym_motor_t *h, int field, const void *src, void *stream) { return motor_field(h, field, (float *)const_cast<void *>(src), 1, stream); }
extern "C" int urgym_motor_stats(urgym_motor_t *h, double *out8, int reset) {
    if (!h || !out8) return URGYM_EINVAL;
    MCK(cudaSetDevice(h->device));
    unsigned long long v[8];
    MCK(cudaMemcpy(v, h->d_stats, sizeof(v), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; k++) out8[k] = (double)v[k];
    out8[1] = (double)(long long)v[1] / 65536.0;
    if (reset) MCK(cudaMemset(h->d_stats, 0, sizeof(v)));
    return URGYM_OK;
}
extern "C" int64_t urgym_motor_launch_count(const urgym_motor_t *h) { return h ? h->launches : 0; }
