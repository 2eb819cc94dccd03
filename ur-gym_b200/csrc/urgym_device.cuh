// urgym_device.cuh -- device-side math of the batched UR5e reach simulator (sm_100a, FP32).
//
// Everything the step needs per environment, one environment per thread:
//   action -> joint target          UR_gym/envs/robots/UR5.py:273-279,304-318
//   UR5e forward kinematics         getLinkState      UR_gym/pyb_setup.py:221-253 (ur5e.urdf:232-279)
//   Euler / quaternion conventions  getEulerFromQuaternion / getQuaternionFromEuler / getDifferenceQuaternion /
//                                   getAxisAngleFromQuaternion   pyb_setup.py:152,190,248,359,363
//   closest-point distances         getClosestPoints  pyb_setup.py:382-456   (GJK on convex cores + margins)
//   distance / angular_distance     UR_gym/utils.py:5-69
//   Euler samplers                  UR_gym/utils.py:81-101
// No tensor cores: the work is chains of 3x3 transforms and small convex-distance iterations, not a contraction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// Every function here is __host__ __device__: the product only ever runs them inside kernels; the host
// instantiation exists so tests/hostcheck can unit-test the very same arithmetic against the oracle on a box
// without a GPU (it is not linked into liburgym_b200.so's call paths -- there is no CPU fallback).
#define URGYM_HD __host__ __device__ __forceinline__
// Out-of-line helpers: the libdevice bodies of sincosf / atan2f (range reduction, slow paths) and the Philox rounds
// are big; one shared copy each keeps the step kernel's instruction footprint inside the instruction cache (the
// first profile of the fully inlined kernel -- 13 k SASS instructions -- stalled on instruction fetch).
// warp-level barrier inside the per-env functions (the capsule scratch overlays the warp's I/O tile); nothing to do
// in the single-threaded host instantiation
#ifdef __CUDA_ARCH__
#define URGYM_WARP_SYNC() __syncwarp()
#else
#define URGYM_WARP_SYNC() ((void)0)
#endif
#if defined(__CUDACC__)
#define URGYM_OOL __host__ __device__ __noinline__
#else
#define URGYM_OOL __attribute__((noinline))
#endif

namespace urgym {

// sin / cos of an angle of a few tens of radians at most (joint angles grow by <= 0.1 pi per step for <= 100 steps):
// two-constant Cody-Waite reduction to [-pi, pi], then the SFU (MUFU.SIN / MUFU.COS, abs error 2^-21.4 on that
// range = 3.6e-7, i.e. < 1e-6 m at the end of the 1 m chain: inside the 1e-5 m / 1e-5 rad FK tolerance).  About
// 10 instructions instead of libdevice's ~45 + slow path, and it is called 19 times per env step.
// The host instantiation (tests/hostcheck) uses libm.
URGYM_HD void sincos_fast(float x, float *s, float *c) {
#ifdef __CUDA_ARCH__
    float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);      // 2 pi rounded to float
    r = fmaf(-k, -1.7484555e-7f, r);                 // 2 pi - float(2 pi)
    *s = __sinf(r);
    *c = __cosf(r);
#else
    sincosf(x, s, c);
#endif
}
// the same for an argument known to lie in [-pi, pi] (half Euler angles, half rotation angles): no reduction
URGYM_HD void sincos_small(float x, float *s, float *c) {
#ifdef __CUDA_ARCH__
    *s = __sinf(x);
    *c = __cosf(x);
#else
    sincosf(x, s, c);
#endif
}
// atan2 with 4e-7 absolute accuracy in ~23 instructions (libdevice's is ~55 and is called 7 times per env step):
// reduce to t = min(|x|,|y|)/max(|x|,|y|) in [0,1], minimax odd polynomial for atan(t), undo the reductions.
// The host instantiation uses libm.
URGYM_HD float atan2_inl(float y, float x) {
#ifdef __CUDA_ARCH__
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
    float t2 = t * t;
    // atan(t)/t on [0,1] as a degree-6 polynomial in t^2 (weighted minimax fit: max abs error of atan 2.5e-7 in exact
    // arithmetic, 3.4e-7 evaluated in FP32; two more terms buy nothing in FP32 -- 1.2e-7 -- and cost two dependent FMAs
    // in each of the seven atan2 of a step)
    float p = 0.0068117789924144745f;
    p = fmaf(p, t2, -0.03360418975353241f);
    p = fmaf(p, t2, 0.07962365448474884f);
    p = fmaf(p, t2, -0.1323334276676178f);
    p = fmaf(p, t2, 0.19807817041873932f);
    p = fmaf(p, t2, -0.3331736922264099f);
    p = fmaf(p, t2, 0.9999961256980896f);
    float r = p * t;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.0f) r = 3.14159274f - r;
    return copysignf(r, y);
#else
    return atan2f(y, x);
#endif
}
static URGYM_OOL float atan2_fast(float y, float x) { return atan2_inl(y, x); }
// three atan2 in one out-of-line call (roll, pitch, yaw): the Euler extractions need three at a time, their polynomials
// (serial chains of nine FMAs each) interleave, and two call / return pairs go away (-1.4 % step-kernel time)
static URGYM_OOL float3 atan2_fast_x3(float y0, float x0, float y1, float x1, float y2, float x2) {
    return make_float3(atan2_inl(y0, x0), atan2_inl(y1, x1), atan2_inl(y2, x2));
}
URGYM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
URGYM_HD float rsqrt_f(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
URGYM_HD int __ffs_hd(unsigned x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
URGYM_HD float fdiv(float a, float b) {               // division where 2 ulp are enough (clamped line parameters)
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// ------------------------------------------------------------------------------------------------ constants
struct alignas(16) ModelConst {
    // table / track constants of the branch-free test of box_tests (urgym_env.cuh); both boxes are centred on y = 0:
    //   [0] top z  [1] cx  [2] hx  [3] hy | [4] cx - hx  [5] cx + hx  [6] -hy  [7] hy
    float box_k[2][8];
    float box_reach[8];         // per link: collision margin + fit_box[l] + box margin (the same for both boxes)
    // upper arm (link 2) vs table / track: constants of the separating-axis filter of robot_pass_capsule:
    //   [0] cx  [1] cz  [2] hx  [3] hz  [4] reach = margin + fit_box[2] + box margin
    float sat2[2][8];
    // packed-math copies (first, so that the pairs sit 8-byte aligned in the constant bank)
    float joint_rot_p[6][3][4]; // rows of joint_rot padded to 4: (F[k][0], F[k][1]) is a constant pair
    float cap_pp[7][3][2];      // (cap_p0[l][k], cap_p1[l][k])
    float joint_xyz[6][3];
    float joint_rot[6][9];      // Rz(y)Ry(p)Rx(r) of each joint origin, row-major
    float cap_p0[7][3];         // bounding capsule of each link hull, link frame
    float cap_p1[7][3];
    float cap_m[7];             // capsule margin = bounding radius + hull_margin (so capsule distance <= hull distance)
    float cap_hl[7];            // half length of the capsule segment (+ round-off allowance), for the sphere broad phase
    float cap_ia[7];            // 1 / |segment|^2
    float obst_cap_ie;          // 1 / |obstacle capsule segment|^2 (bounding capsule, hull-mode broad phase)
    // capsule geometry proper (urgym_capsule_fit.h): calibrated against the hull geometry
    float fit_obst[7], fit_box[7], fit_self[9];
    float self_far2[9];         // squared broad-phase thresholds of the self pairs (urgym_model.h)
    float self_reach2[9];       // (margin + fit_self)^2
    float fit_obst_h, fit_obst_ie;
    float box_top;              // highest top face of the table / track cores (z), for the height broad phase
    int hull_off[8];            // vertex ranges of links 0..6 inside the packed float4 vertex array
    // margins of Bullet's GJK pair detector (distance = |core gap| - marginA - marginB)
    float hull_margin;          // URDF mesh links: 0.001
    float box_c[2][3], box_he[2][3], box_margin[2];    // 0 = table, 1 = track; core half extents (shrunk by the margin)
    float obst_r, obst_h, obst_margin;             // hull mode: cylinder core radius / half height, margin
    float tgt_box_he, tgt_box_margin;              // hull mode: Sta/Dyn target box core
    float tgt_sphere_margin;                       // Obs target sphere: point core, margin = radius
    // capsule geometry mode stand-ins (bounding shapes of the reference's obstacle / target)
    float obst_cap_h, obst_cap_m;                  // obstacle: segment half length, margin (= radius)
    float tgt_cap_m[4];                            // target as a sphere, per task (Ori unused)
    // the reset pose (UR5.py:262) is the same for every env: its FK is computed once on the host in double
    float neutral_q[6];
    float neutral_ee[6];                           // EE position + PyBullet Euler triple
    float neutral_ca[7][3], neutral_cb[7][3];      // world capsule segments of links 0..6
    float neutral_R[7][9], neutral_p[7][3];        // world link poses (hull mode)
};

#define URGYM_PI_F 3.14159265358979323846f

// ------------------------------------------------------------------------------------------------ float3 helpers
URGYM_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
URGYM_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
URGYM_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
URGYM_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
URGYM_HD float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
URGYM_HD float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
URGYM_HD float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
URGYM_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ------------------------------------------------------------------------------------------------ packed FP32 pairs
// Blackwell's FFMA2 / FMUL2 / FADD2 do two FP32 operations in one issue slot (operands are 64-bit register pairs,
// a scalar register or constant can be broadcast to both halves).  The step kernel is bound by the issue rate, so
// the vector math of the kinematic chain is written on pairs.  The host instantiation does the two halves in turn
// with the same roundings.
URGYM_HD float2 f2(float x, float y) { return make_float2(x, y); }
URGYM_HD float2 bc2(float s) { return make_float2(s, s); }
URGYM_HD float2 fma2(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
URGYM_HD float2 mul2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
URGYM_HD float2 add2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}

struct Pose {           // world pose of a link frame
    float R[9];         // row-major
    float3 p;
};
URGYM_HD float3 rot(const float *R, float3 v) {
    return f3(fmaf(R[0], v.x, fmaf(R[1], v.y, R[2] * v.z)), fmaf(R[3], v.x, fmaf(R[4], v.y, R[5] * v.z)),
              fmaf(R[6], v.x, fmaf(R[7], v.y, R[8] * v.z)));
}
URGYM_HD float3 rotT(const float *R, float3 v) {
    return f3(fmaf(R[0], v.x, fmaf(R[3], v.y, R[6] * v.z)), fmaf(R[1], v.x, fmaf(R[4], v.y, R[7] * v.z)),
              fmaf(R[2], v.x, fmaf(R[5], v.y, R[8] * v.z)));
}

// ------------------------------------------------------------------------------------------------ quaternions (x,y,z,w)
struct Quat { float x, y, z, w; };

// pybullet getQuaternionFromEuler(roll,pitch,yaw) = Rz(yaw)Ry(pitch)Rx(roll)          pyb_setup.py:152,313-314
URGYM_HD Quat quat_from_euler(float roll, float pitch, float yaw) {
    float sr, cr, sp, cp, sy, cy;
    sincos_fast(0.5f * roll, &sr, &cr);
    sincos_fast(0.5f * pitch, &sp, &cp);
    sincos_fast(0.5f * yaw, &sy, &cy);
    Quat q;
    q.x = sr * cp * cy - cr * sp * sy;
    q.y = cr * sp * cy + sr * cp * sy;
    q.z = cr * cp * sy - sr * sp * cy;
    q.w = cr * cp * cy + sr * sp * sy;
    return q;
}
URGYM_HD Quat quat_mul(Quat a, Quat b) {
    Quat o;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    o.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return o;
}
// pybullet getEulerFromQuaternion -> (roll, pitch, yaw), gimbal branch at |sarg| >= 0.99999   pyb_setup.py:190,248
URGYM_HD float3 euler_from_quat(Quat q) {
    float sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, squ = q.w * q.w;
    float sarg = -2.0f * (q.x * q.z - q.w * q.y);
    float3 e;
    if (sarg <= -0.99999f) {
        e.y = -0.5f * URGYM_PI_F; e.x = 0.0f; e.z = 2.0f * atan2_fast(q.x, -q.y);
    } else if (sarg >= 0.99999f) {
        e.y = 0.5f * URGYM_PI_F; e.x = 0.0f; e.z = 2.0f * atan2_fast(-q.x, q.y);
    } else {
        // pitch = asin(sarg), evaluated as atan2(sin, cos) with cos(pitch) = |(R21, R22)|: same angle, but not
        // ill-conditioned in FP32 when |pitch| approaches 90 degrees
        float r21 = 2.0f * (q.y * q.z + q.w * q.x), r22 = squ - sqx - sqy + sqz;
        e = atan2_fast_x3(r21, r22, sarg, sqrtf(r21 * r21 + r22 * r22), 2.0f * (q.x * q.y + q.w * q.z), squ + sqx - sqy - sqz);
    }
    return e;
}
// btMatrix3x3::getRotation
URGYM_HD Quat quat_from_mat(const float *R) {
    Quat q;
    float tr = R[0] + R[4] + R[8];
    if (tr > 0.0f) {
        float s = sqrtf(tr + 1.0f);
        q.w = 0.5f * s; s = 0.5f / s;
        q.x = (R[7] - R[5]) * s; q.y = (R[2] - R[6]) * s; q.z = (R[3] - R[1]) * s;
    } else if (R[0] >= R[4] && R[0] >= R[8]) {
        float s = sqrtf(R[0] - R[4] - R[8] + 1.0f);
        q.x = 0.5f * s; s = 0.5f / s;
        q.w = (R[7] - R[5]) * s; q.y = (R[3] + R[1]) * s; q.z = (R[6] + R[2]) * s;
    } else if (R[4] >= R[8]) {
        float s = sqrtf(R[4] - R[8] - R[0] + 1.0f);
        q.y = 0.5f * s; s = 0.5f / s;
        q.w = (R[2] - R[6]) * s; q.z = (R[7] + R[5]) * s; q.x = (R[3] + R[1]) * s;
    } else {
        float s = sqrtf(R[8] - R[0] - R[4] + 1.0f);
        q.z = 0.5f * s; s = 0.5f / s;
        q.w = (R[3] - R[1]) * s; q.x = (R[6] + R[2]) * s; q.y = (R[7] + R[5]) * s;
    }
    return q;
}
// Euler triple of a rotation matrix the way PyBullet produces it: matrix -> quaternion (btMatrix3x3::getRotation)
// -> getEulerFromQuaternion.  The quaternion is normalised on the way: an FP32 chain product is orthonormal only to
// a few 1e-7, and reading roll and yaw from different entries of such a matrix amplifies that defect by
// 1/cos(pitch) near gimbal lock; a unit quaternion is an exact rotation, so the triple always encodes a rotation
// within round-off of the true one.
static URGYM_OOL float3 euler_via_quat_ool(const float *R) {     // near gimbal lock: rare, kept out of the hot instruction stream
    Quat q = quat_from_mat(R);
    float n = rsqrt_f(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x *= n; q.y *= n; q.z *= n; q.w *= n;
    return euler_from_quat(q);
}
URGYM_HD float3 euler_from_mat(const float *R) {
    // For a unit quaternion the regular branch of getEulerFromQuaternion reads
    //   -2(xz-wy) = -R20,  2(yz+wx) = R21,  w2-x2-y2+z2 = R22,  2(xy+wz) = R10,  w2+x2-y2-z2 = R00,
    // so away from gimbal lock (amplification 1/cos(pitch) < 7) the triple comes straight from the matrix.
    const float sarg = -R[6];
    if (fabsf(sarg) < 0.99f)
        return f3(atan2_fast(R[7], R[8]), atan2_fast(sarg, sqrtf(R[7] * R[7] + R[8] * R[8])), atan2_fast(R[3], R[0]));
    return euler_via_quat_ool(R);
}
URGYM_HD void mat_from_quat(Quat q, float *R) {
    float n = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w, s = 2.0f / n;
    R[0] = 1.0f - s * (q.y * q.y + q.z * q.z); R[1] = s * (q.x * q.y - q.w * q.z); R[2] = s * (q.x * q.z + q.w * q.y);
    R[3] = s * (q.x * q.y + q.w * q.z); R[4] = 1.0f - s * (q.x * q.x + q.z * q.z); R[5] = s * (q.y * q.z - q.w * q.x);
    R[6] = s * (q.x * q.z - q.w * q.y); R[7] = s * (q.y * q.z + q.w * q.x); R[8] = 1.0f - s * (q.x * q.x + q.y * q.y);
}

// scipy Rotation.from_euler('ZYX', e).as_quat(): R = Rz(e0)Ry(e1)Rx(e2)   utils.py:48-54 (quirk Q2: e is PyBullet's
// (roll,pitch,yaw), so roll is used as the z angle -- replicated literally)
// SMALL: the angles are Euler angles that came out of atan2 (|e| <= pi): no range reduction
template <bool SMALL = false>
URGYM_HD Quat quat_ZYX(float e0, float e1, float e2) {
    float sz, cz, sy, cy, sx, cx;
    if (SMALL) {
        sincos_small(0.5f * e0, &sz, &cz);
        sincos_small(0.5f * e1, &sy, &cy);
        sincos_small(0.5f * e2, &sx, &cx);
    } else {
        sincos_fast(0.5f * e0, &sz, &cz);
        sincos_fast(0.5f * e1, &sy, &cy);
        sincos_fast(0.5f * e2, &sx, &cx);
    }
    Quat q;
    q.x = cz * cy * sx - sz * sy * cx;
    q.y = cz * sy * cx + sz * cy * sx;
    q.z = sz * cy * cx - cz * sy * sx;
    q.w = cz * cy * cx + sz * sy * sx;
    return q;
}
// utils.py:34-69: 2*arccos(|<qa,qb>|).  Evaluated as 2*atan2(|vec(qa^-1 qb)|, |<qa,qb>|), the same angle but
// well conditioned in FP32 near 0 (arccos near 1 would lose half the digits).
URGYM_HD float angular_distance(Quat a, Quat b) {
    float d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    float vx = a.w * b.x - a.x * b.w - a.y * b.z + a.z * b.y;
    float vy = a.w * b.y - a.y * b.w - a.z * b.x + a.x * b.z;
    float vz = a.w * b.z - a.z * b.w - a.x * b.y + a.y * b.x;
    return 2.0f * atan2_fast(sqrtf(vx * vx + vy * vy + vz * vz), fabsf(d));
}

// ------------------------------------------------------------------------------------------------ forward kinematics
// advance the chain by joint j (0..5): pose of PyBullet link j+1 from the pose of link j
//   T_{j+1} = T_j * Trans(xyz_j) * Rz(y)Ry(p)Rx(r) * Rz(q_j)                      ur5e.urdf:232-279
URGYM_HD void fk_advance(const ModelConst &M, Pose &T, int j, float qj) {
    const float *X = M.joint_xyz[j];
    const float *F = M.joint_rot[j];
    T.p = T.p + rot(T.R, f3(X[0], X[1], X[2]));
    float A[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            A[3 * r + c] = fmaf(T.R[3 * r], F[c], fmaf(T.R[3 * r + 1], F[3 + c], T.R[3 * r + 2] * F[6 + c]));
    float s, c;
    sincos_fast(qj, &s, &c);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        T.R[3 * r] = fmaf(A[3 * r], c, A[3 * r + 1] * s);
        T.R[3 * r + 1] = fmaf(A[3 * r + 1], c, -A[3 * r] * s);
        T.R[3 * r + 2] = A[3 * r + 2];
    }
}
URGYM_HD void pose_identity(Pose &T) {
    T.R[0] = 1; T.R[1] = 0; T.R[2] = 0; T.R[3] = 0; T.R[4] = 1; T.R[5] = 0; T.R[6] = 0; T.R[7] = 0; T.R[8] = 1;
    T.p = f3(0, 0, 0);
}
// The same chain on packed pairs.  R is kept by columns, each column as its (x, y) pair plus the z parts:
//   ck = (R[0][k], R[1][k]) for k = 0..2,  z01 = (R[2][0], R[2][1]),  z2 = R[2][2];  position (pxy, pz).
struct PoseP {
    float2 c0, c1, c2, z01, pxy;
    float z2, pz;
};
URGYM_HD void posep_identity(PoseP &T) {
    T.c0 = f2(1, 0); T.c1 = f2(0, 1); T.c2 = f2(0, 0); T.z01 = f2(0, 0); T.z2 = 1; T.pxy = f2(0, 0); T.pz = 0;
}
URGYM_HD void fkp_advance(const ModelConst &M, PoseP &T, int j, float qj) {
    const float *X = M.joint_xyz[j];
    const float (*F)[4] = M.joint_rot_p[j];
    T.pxy = fma2(T.c0, bc2(X[0]), fma2(T.c1, bc2(X[1]), fma2(T.c2, bc2(X[2]), T.pxy)));
    T.pz = fmaf(T.z01.x, X[0], fmaf(T.z01.y, X[1], fmaf(T.z2, X[2], T.pz)));
    // A = R F: (x, y) parts of the three columns, then the z parts ((A20, A21) as a pair)
    const float2 a0 = fma2(T.c0, bc2(F[0][0]), fma2(T.c1, bc2(F[1][0]), mul2(T.c2, bc2(F[2][0]))));
    const float2 a1 = fma2(T.c0, bc2(F[0][1]), fma2(T.c1, bc2(F[1][1]), mul2(T.c2, bc2(F[2][1]))));
    const float2 a2 = fma2(T.c0, bc2(F[0][2]), fma2(T.c1, bc2(F[1][2]), mul2(T.c2, bc2(F[2][2]))));
    const float2 az = fma2(bc2(T.z01.x), f2(F[0][0], F[0][1]),
                           fma2(bc2(T.z01.y), f2(F[1][0], F[1][1]), mul2(bc2(T.z2), f2(F[2][0], F[2][1]))));
    const float az2 = fmaf(T.z01.x, F[0][2], fmaf(T.z01.y, F[1][2], T.z2 * F[2][2]));
    float s, c;
    sincos_fast(qj, &s, &c);
    T.c0 = fma2(a0, bc2(c), mul2(a1, bc2(s)));
    T.c1 = fma2(a1, bc2(c), mul2(a0, bc2(-s)));
    T.c2 = a2;
    T.z01 = f2(fmaf(az.x, c, az.y * s), fmaf(az.y, c, -az.x * s));
    T.z2 = az2;
}
// joint 0 from the identity pose: R = F Rz(q), p = x  (a third of the general step's instructions)
URGYM_HD void fkp_first(const ModelConst &M, PoseP &T, float q0) {
    const float (*F)[4] = M.joint_rot_p[0];
    float s, c;
    sincos_fast(q0, &s, &c);
    T.c0 = fma2(f2(F[0][0], F[1][0]), bc2(c), mul2(f2(F[0][1], F[1][1]), bc2(s)));
    T.c1 = fma2(f2(F[0][1], F[1][1]), bc2(c), mul2(f2(F[0][0], F[1][0]), bc2(-s)));
    T.c2 = f2(F[0][2], F[1][2]);
    T.z01 = f2(fmaf(F[2][0], c, F[2][1] * s), fmaf(F[2][1], c, -F[2][0] * s));
    T.z2 = F[2][2];
    T.pxy = f2(M.joint_xyz[0][0], M.joint_xyz[0][1]); T.pz = M.joint_xyz[0][2];
}
// world capsule segment of link l: a = p + R cap_p0, b = p + R cap_p1; returns (a.x, a.y), (b.x, b.y), (a.z, b.z)
URGYM_HD void capsule_world(const ModelConst &M, const PoseP &T, int l, float2 &axy, float2 &bxy, float2 &abz) {
    const float (*P)[2] = M.cap_pp[l];
    axy = fma2(T.c0, bc2(P[0][0]), fma2(T.c1, bc2(P[1][0]), fma2(T.c2, bc2(P[2][0]), T.pxy)));
    bxy = fma2(T.c0, bc2(P[0][1]), fma2(T.c1, bc2(P[1][1]), fma2(T.c2, bc2(P[2][1]), T.pxy)));
    abz = fma2(bc2(T.z01.x), f2(P[0][0], P[0][1]),
               fma2(bc2(T.z01.y), f2(P[1][0], P[1][1]), fma2(bc2(T.z2), f2(P[2][0], P[2][1]), bc2(T.pz))));
}
// PyBullet Euler triple of the pose (see euler_from_mat): needs R20, R21, R22, R10, R00
URGYM_HD float3 euler_from_posep(const PoseP &T) {
    const float sarg = -T.z01.x;
    if (fabsf(sarg) < 0.99f)
        return atan2_fast_x3(T.z01.y, T.z2, sarg, sqrtf(T.z01.y * T.z01.y + T.z2 * T.z2), T.c0.y, T.c0.x);
    const float R[9] = {T.c0.x, T.c1.x, T.c2.x, T.c0.y, T.c1.y, T.c2.y, T.z01.x, T.z01.y, T.z2};
    return euler_via_quat_ool(R);
}
// pose of link `link` (1..6)
URGYM_HD void fk_link(const ModelConst &M, const float *q, int link, Pose &T) {
    pose_identity(T);
    for (int j = 0; j < link; j++) fk_advance(M, T, j, q[j]);
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Philox4x32-10 (Salmon et al., SC'11).  One 32x32->64 multiply gives both halves of each product (IMAD.WIDE).
static URGYM_OOL uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r) { k.x += 0x9E3779B9u; k.y += 0xBB67AE85u; }
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
    }
    return c;
}
// N independent blocks with their rounds interleaved: a single Philox block is a serial chain of ten dependent
// multiply rounds, and the auto-reset kernel is bound by exactly that latency (few warps, long dependent chains).
template <int N> URGYM_HD void philox4x32_10_n(uint4 (&c)[N], uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r) { k.x += 0x9E3779B9u; k.y += 0xBB67AE85u; }
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c[i].x, p1 = (uint64_t)0xCD9E8D57u * c[i].z;
            c[i] = make_uint4((uint32_t)(p1 >> 32) ^ c[i].y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c[i].w ^ k.y, (uint32_t)p0);
        }
    }
}
URGYM_HD float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }   // 2^-24

struct ResetStream {        // counter = (iteration*blocks_per_iter + block, episode, env_lo, env_hi); key = seed
    uint2 key; uint32_t episode, env_lo, env_hi, bpi, iter;
    URGYM_HD uint4 block(uint32_t b) const {
        return philox4x32_10(make_uint4(iter * bpi + b, episode, env_lo, env_hi), key);
    }
    template <int N> URGYM_HD void blocks(uint32_t b0, uint4 *out) const {      // blocks b0 .. b0 + N - 1, interleaved
        uint4 c[N];
#pragma unroll
        for (int i = 0; i < N; i++) c[i] = make_uint4(iter * bpi + b0 + (uint32_t)i, episode, env_lo, env_hi);
        philox4x32_10_n<N>(c, key);
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = c[i];
    }
};

// utils.py:81-86  sample_euler_constrained: deg2rad([U(-90,-180), 0, U(0,-180)]) with numpy's low+(high-low)*u
URGYM_HD void euler_constrained(float u_roll, float u_yaw, float &roll, float &yaw) {
    const float D2R = URGYM_PI_F / 180.0f;
    roll = (-90.0f + -90.0f * u_roll) * D2R;
    yaw = (-180.0f * u_yaw) * D2R;
}
// utils.py:88-101  sample_euler_obstacle
URGYM_HD void euler_obstacle(float u_sign, float u_roll, float u_pitch, float &roll, float &pitch) {
    const float D2R = URGYM_PI_F / 180.0f;
    float r = (u_sign < 0.5f) ? (-30.0f + -120.0f * u_roll) : (30.0f + 120.0f * u_roll);
    float p = (r < -90.0f || r > 90.0f) ? (-30.0f + -120.0f * u_pitch) : (30.0f + 120.0f * u_pitch);
    roll = r * D2R; pitch = p * D2R;
}

// ------------------------------------------------------------------------------------------------ convex cores
// support points of the cores Bullet's GJK sees (btConvexShape::localGetSupportVertexWithoutMarginNonVirtual)
// The hull "blob" every hull-geometry kernel stages in shared memory: float4 vertices of all links (link frame), then
// the adjacency of the triangulated hulls (tools/make_hull_adjacency.py): uint16 offsets per vertex (+1), uint16 LOCAL
// neighbour indices.  Sizes in ur5e_hull_adjacency.h / ur5e_model_data.h; layout fixed here.
#define URGYM_HULL_NV 3793
#define URGYM_HULL_NADJ 22674
#define URGYM_HULL_OFF_U16 3800                     /* NV + 1 offsets, padded to a multiple of 8 */
#define URGYM_HULL_ADJ_U16 22680                    /* padded to a multiple of 8 */
// ... then the start table of the hill-climbing support function: per link, for each cell of an 8 x 8 grid on each face of
// the direction cube (link frame), the support vertex of the cell's centre direction (LOCAL index), built at create time.
#define URGYM_HULL_DIR_CELLS 384                    /* 6 faces x 8 x 8 */
#define URGYM_HULL_DIR_U16 (7 * URGYM_HULL_DIR_CELLS)
#define URGYM_HULL_BLOB_F4 (URGYM_HULL_NV + (URGYM_HULL_OFF_U16 + URGYM_HULL_ADJ_U16 + URGYM_HULL_DIR_U16) / 8)
URGYM_HD const unsigned short *hull_adj_off(const float4 *blob) { return reinterpret_cast<const unsigned short *>(blob + URGYM_HULL_NV); }
URGYM_HD const unsigned short *hull_adj(const float4 *blob) { return hull_adj_off(blob) + URGYM_HULL_OFF_U16; }
URGYM_HD const unsigned short *hull_dirmap(const float4 *blob) { return hull_adj(blob) + URGYM_HULL_ADJ_U16; }
// cell of direction l on the direction cube: face = dominant axis and sign, (u, v) = the other two components / |dominant|
URGYM_HD int hull_dir_cell(float3 l) {
    const float ax = fabsf(l.x), ay = fabsf(l.y), az = fabsf(l.z);
    int face; float m, u, v;
    if (ax >= ay && ax >= az) { face = l.x > 0.0f ? 0 : 1; m = ax; u = l.y; v = l.z; }
    else if (ay >= az) { face = l.y > 0.0f ? 2 : 3; m = ay; u = l.x; v = l.z; }
    else { face = l.z > 0.0f ? 4 : 5; m = az; u = l.x; v = l.y; }
    if (!(m > 0.0f)) return 0;
    const float k = 4.0f / m;
    int iu = (int)fmaf(u, k, 4.0f), iv = (int)fmaf(v, k, 4.0f);
    iu = iu < 0 ? 0 : (iu > 7 ? 7 : iu); iv = iv < 0 ? 0 : (iv > 7 ? 7 : iv);
    return face * 64 + iv * 8 + iu;
}
// centre direction of a cell (host side, table construction)
inline void hull_cell_dir(int cell, double d[3]) {
    const int face = cell / 64, iv = (cell / 8) % 8, iu = cell % 8;
    const double u = (iu + 0.5) / 4.0 - 1.0, v = (iv + 0.5) / 4.0 - 1.0, sg = (face & 1) ? -1.0 : 1.0;
    if (face < 2) { d[0] = sg; d[1] = u; d[2] = v; }
    else if (face < 4) { d[0] = u; d[1] = sg; d[2] = v; }
    else { d[0] = u; d[1] = v; d[2] = sg; }
}

// Support vertex of a link hull by steepest ascent over its edges: on a convex polytope a vertex with no better
// neighbour is the maximum.  l: direction in the link frame.  Starts from the table's vertex for l (the exact support
// vertex of a direction a few degrees away) or from the previous query's vertex `cur` when that one is better.
// Out of line with every argument and the result in registers: handing it the hull by reference had put the link pose and
// the hull descriptor of every caller into local memory, and with 215 KB of the SM's 256 KB configured as shared memory
// the 512 threads' stack frames do not fit the remaining L1 (long-scoreboard stall 6.8 warps per issue slot).
static URGYM_OOL int hull_climb(const float4 *v, const unsigned short *aoff, const unsigned short *adj,
                                const unsigned short *start, int n, int cur, float lx, float ly, float lz) {
    int c = start[hull_dir_cell(f3(lx, ly, lz))];
    float4 p = v[c];
    float best = fmaf(p.x, lx, fmaf(p.y, ly, p.z * lz));
    if (cur >= 0) {
        const float4 pc = v[cur];
        const float sc = fmaf(pc.x, lx, fmaf(pc.y, ly, pc.z * lz));
        if (sc > best) { best = sc; c = cur; }
    }
    for (int it = 0; it < n; it++) {
        int nxt = -1;
        const int e1 = aoff[c + 1];
        for (int e = aoff[c]; e < e1; e++) {
            const int k = adj[e];
            const float4 q = v[k];
            const float s = fmaf(q.x, lx, fmaf(q.y, ly, q.z * lz));
            if (s > best) { best = s; nxt = k; }
        }
        if (nxt < 0) break;
        c = nxt;
    }
    return c;
}
struct HullW {                      // link hull: vertices in the link frame (shared memory), posed by T (held by value)
    const float4 *v; int n; Pose T;
    const unsigned short *aoff, *adj;       // adjacency of THIS link: aoff[k] .. aoff[k + 1] index adj (local neighbour ids)
    const unsigned short *start;            // start table of this link (hull_dirmap)
    mutable int cur;                        // last support vertex: the next query starts its walk there (-1: none yet)
    URGYM_HD float3 center() const { return T.p; }
    URGYM_HD float3 support(float3 d) const {
        const float3 l = rotT(T.R, d);
        cur = hull_climb(v, aoff, adj, start, n, cur, l.x, l.y, l.z);
        const float4 p = v[cur];
        return rot(T.R, f3(p.x, p.y, p.z)) + T.p;
    }
};
struct SegW {                       // capsule core: a segment in world space
    float3 a, b;
    URGYM_HD float3 center() const { return 0.5f * (a + b); }
    URGYM_HD float3 support(float3 d) const { return dot(d, b - a) > 0.0f ? b : a; }
};
struct CylW {                       // obstacle core: cylinder about unit axis u through c  (btCylinderShapeZ)
    float3 c, u; float r, h;
    URGYM_HD float3 center() const { return c; }
    URGYM_HD float3 support(float3 d) const {
        float dz = dot(d, u);
        float3 dp = d - dz * u;
        dp = dp - dot(dp, u) * u;       // second Gram-Schmidt pass: when d is nearly parallel to the axis the first
                                        // difference is mostly round-off and would tilt the rim point off the cap
        float s2 = dot(dp, dp);
        float3 o = c + (dz < 0.0f ? -h : h) * u;
        if (s2 > 0.0f) o = o + (r * rsqrt_f(s2)) * dp;
        return o;
    }
};
struct BoxA {                       // axis-aligned box core (table, track)
    float3 c, he;
    URGYM_HD float3 center() const { return c; }
    URGYM_HD float3 support(float3 d) const {
        return f3(c.x + (d.x >= 0.0f ? he.x : -he.x), c.y + (d.y >= 0.0f ? he.y : -he.y), c.z + (d.z >= 0.0f ? he.z : -he.z));
    }
};
struct BoxO {                       // oriented cube core (Sta/Dyn target)
    float3 c; float R[9]; float he;
    URGYM_HD float3 center() const { return c; }
    URGYM_HD float3 support(float3 d) const {
        float3 l = rotT(R, d);
        return c + rot(R, f3(l.x >= 0.0f ? he : -he, l.y >= 0.0f ? he : -he, l.z >= 0.0f ? he : -he));
    }
};

// closest point to the origin on triangle abc, barycentric weights.  Interior solution from the 2x2 normal
// equations; when it falls outside (or the triangle is degenerate) the best of the three edge projections.
// Written for FP32 robustness rather than minimum flops: region classification by products of large dot products
// (the classic formulation) stalls GJK when the simplex is small next to its distance from the origin.
static URGYM_OOL void closest_tri(float3 a, float3 b, float3 c, float &la, float &lb, float &lc) {
    float3 ab = b - a, ac = c - a, bc = c - b;
    float s = -dot(a, ab), t = -dot(a, ac);
    float E = dot(ab, ab), F = dot(ab, ac), G = dot(ac, ac);
    float det = E * G - F * F;
    if (det > 1e-10f * E * G) {
        float u = s * G - t * F, v = t * E - s * F;
        if (u >= 0.0f && v >= 0.0f && u + v <= det) {
            float inv = 1.0f / det;
            lb = u * inv; lc = v * inv; la = 1.0f - lb - lc;
            return;
        }
    }
    float t1 = E > 0.0f ? clampf(s / E, 0.0f, 1.0f) : 0.0f;
    float t2 = G > 0.0f ? clampf(t / G, 0.0f, 1.0f) : 0.0f;
    float H = dot(bc, bc);
    float t3 = H > 0.0f ? clampf(-dot(b, bc) / H, 0.0f, 1.0f) : 0.0f;
    float3 x1 = a + t1 * ab, x2 = a + t2 * ac, x3 = b + t3 * bc;
    float d1 = dot(x1, x1), d2 = dot(x2, x2), d3 = dot(x3, x3);
    if (d1 <= d2 && d1 <= d3) { la = 1.0f - t1; lb = t1; lc = 0.0f; }
    else if (d2 <= d3) { la = 1.0f - t2; lb = 0.0f; lc = t2; }
    else { la = 0.0f; lb = 1.0f - t3; lc = t3; }
}

// Reduce the simplex to the sub-simplex supporting the point closest to the origin (v).  true = origin enclosed.
static URGYM_OOL bool closest_simplex(float3 (&W)[4], int &n, float3 &v) {
    float l[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (n == 1) {
        l[0] = 1.0f;
    } else if (n == 2) {
        float3 ab = W[1] - W[0];
        float den = dot(ab, ab), t = den > 0.0f ? -dot(W[0], ab) / den : 0.0f;
        t = clampf(t, 0.0f, 1.0f);
        l[0] = 1.0f - t; l[1] = t;
    } else if (n == 3) {
        closest_tri(W[0], W[1], W[2], l[0], l[1], l[2]);
    } else {
        // tetrahedron: barycentric coordinates of the origin from signed volumes.  The origin is enclosed only if
        // all four are positive and the tetrahedron is not (numerically) flat; otherwise the closest point lies on
        // a face whose opposite vertex has a non-positive coordinate.
        float3 ad = W[0] - W[3], bd = W[1] - W[3], cd = W[2] - W[3], od = -W[3];
        float det = dot(ad, cross(bd, cd));
        float L2 = fmaxf(fmaxf(dot(ad, ad), dot(bd, bd)), dot(cd, cd));
        bool flat = det * det <= 1e-10f * L2 * L2 * L2;
        float la = dot(od, cross(bd, cd)), lb = dot(ad, cross(od, cd)), lc = dot(ad, cross(bd, od));
        if (det < 0.0f) { la = -la; lb = -lb; lc = -lc; det = -det; }
        float ld = det - la - lb - lc;
        if (!flat && la > 0.0f && lb > 0.0f && lc > 0.0f && ld > 0.0f) return true;
        float best = 3.0e38f;
        float a, b, c;
        if (flat || ld <= 0.0f) {       // face 0 1 2 (opposite vertex 3)
            closest_tri(W[0], W[1], W[2], a, b, c);
            float3 p = a * W[0] + b * W[1] + c * W[2]; float dd = dot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = b; l[2] = c; l[3] = 0; }
        }
        if (flat || lb <= 0.0f) {       // face 0 2 3 (opposite vertex 1)
            closest_tri(W[0], W[2], W[3], a, b, c);
            float3 p = a * W[0] + b * W[2] + c * W[3]; float dd = dot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = 0; l[2] = b; l[3] = c; }
        }
        if (flat || lc <= 0.0f) {       // face 0 3 1 (opposite vertex 2)
            closest_tri(W[0], W[3], W[1], a, b, c);
            float3 p = a * W[0] + b * W[3] + c * W[1]; float dd = dot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = c; l[2] = 0; l[3] = b; }
        }
        if (flat || la <= 0.0f) {       // face 1 3 2 (opposite vertex 0)
            closest_tri(W[1], W[3], W[2], a, b, c);
            float3 p = a * W[1] + b * W[3] + c * W[2]; float dd = dot(p, p);
            if (dd < best) { best = dd; l[0] = 0; l[1] = a; l[2] = c; l[3] = b; }
        }
    }
    float3 nv = f3(0, 0, 0);
    float3 T[4];
    int m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < n && l[i] > 0.0f) {
            nv = nv + l[i] * W[i];
#pragma unroll
            for (int j = 0; j <= i; j++)
                if (m == j) T[j] = W[i];
            m++;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < m) W[i] = T[i];
    n = m; v = nv;
    return false;
}

#ifndef URGYM_GJK_MAX_ITER
#define URGYM_GJK_MAX_ITER 48
#endif
#define URGYM_GJK_REL_TOL 1.0e-6f     /* Bullet's REL_ERROR2 on the squared distance (btGjkPairDetector) */

// distance between the cores of A and B (0 and deep=true when they intersect)
template <class SA, class SB>
URGYM_HD float gjk_distance(const SA &A, const SB &B, bool &deep, float3 *v_out = nullptr, float *lb_out = nullptr) {
    float3 W[4];
    int n = 1;
    float3 d = A.center() - B.center();
    if (dot(d, d) < 1e-12f) d = f3(1, 0, 0);
    float3 v = A.support(-d) - B.support(d);
    W[0] = v;
    float vv = dot(v, v);
    float lb = 0.0f;                    // best proven lower bound of the distance (separating-axis bound v.w / |v|)
    deep = false;
    for (int it = 0; it < URGYM_GJK_MAX_ITER; it++) {
        if (vv < 1e-14f) { deep = true; vv = 0.0f; break; }
        float3 w = A.support(-v) - B.support(v);
        float delta = dot(v, w);
        if (delta > 0.0f) lb = fmaxf(lb, delta * rsqrt_f(vv));
        if (vv - delta <= URGYM_GJK_REL_TOL * vv) break;
        bool dup = false;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < n) { float3 e = w - W[i]; dup = dup || (dot(e, e) <= 1e-14f); }
        if (dup) break;
        if (n == 1) W[1] = w; else if (n == 2) W[2] = w; else W[3] = w;
        n++;
        float3 vn;
        if (closest_simplex(W, n, vn)) {
            if (lb > 1e-5f) break;      // a separating axis was already found: the enclosure is round-off, keep v
            deep = true; vv = 0.0f; break;
        }
        float vvn = dot(vn, vn);
        if (vvn >= vv) break;
        v = vn; vv = vvn;
    }
    if (v_out) *v_out = v;          // closest vector (from B to A) of the last iterate
    if (lb_out) *lb_out = lb;       // proven lower bound of the distance
    return sqrtf(vv);
}

// ------------------------------------------------------------------------------------------------ FP64 refinement
// The FP32 iteration above ends within ~1e-6 m of the FP64 answer in all but a few pairs per ten thousand: contacts with
// a cylinder RIM (GJK converges only sublinearly on a curved edge: 48 iterations are not enough) and faces that are
// nearly parallel (the FP32 simplex update stalls).  Those are recognisable -- the iteration ends with a gap between its
// proven lower bound and its distance -- and are redone here: the same algorithm with the simplex arithmetic in double
// and up to 256 iterations.  Rare, out of line.
struct D3 { double x, y, z; };
URGYM_HD D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
URGYM_HD D3 d3(float3 a) { return d3((double)a.x, (double)a.y, (double)a.z); }
URGYM_HD D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
URGYM_HD D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
URGYM_HD D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
URGYM_HD double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
URGYM_HD D3 dcross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
URGYM_HD double dclamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static URGYM_OOL void closest_tri_d(D3 a, D3 b, D3 c, double &la, double &lb, double &lc) {
    D3 ab = b - a, ac = c - a, bc = c - b;
    double s = -ddot(a, ab), t = -ddot(a, ac);
    double E = ddot(ab, ab), F = ddot(ab, ac), G = ddot(ac, ac);
    double det = E * G - F * F;
    if (det > 1e-20 * E * G) {
        double u = s * G - t * F, v = t * E - s * F;
        if (u >= 0.0 && v >= 0.0 && u + v <= det) {
            double inv = 1.0 / det;
            lb = u * inv; lc = v * inv; la = 1.0 - lb - lc;
            return;
        }
    }
    double t1 = E > 0.0 ? dclamp(s / E, 0.0, 1.0) : 0.0;
    double t2 = G > 0.0 ? dclamp(t / G, 0.0, 1.0) : 0.0;
    double H = ddot(bc, bc);
    double t3 = H > 0.0 ? dclamp(-ddot(b, bc) / H, 0.0, 1.0) : 0.0;
    D3 x1 = a + t1 * ab, x2 = a + t2 * ac, x3 = b + t3 * bc;
    double d1 = ddot(x1, x1), d2 = ddot(x2, x2), d3_ = ddot(x3, x3);
    if (d1 <= d2 && d1 <= d3_) { la = 1.0 - t1; lb = t1; lc = 0.0; }
    else if (d2 <= d3_) { la = 1.0 - t2; lb = 0.0; lc = t2; }
    else { la = 0.0; lb = 1.0 - t3; lc = t3; }
}
static URGYM_OOL bool closest_simplex_d(D3 (&W)[4], int &n, D3 &v) {
    double l[4] = {0.0, 0.0, 0.0, 0.0};
    if (n == 1) {
        l[0] = 1.0;
    } else if (n == 2) {
        D3 ab = W[1] - W[0];
        double den = ddot(ab, ab), t = den > 0.0 ? -ddot(W[0], ab) / den : 0.0;
        t = dclamp(t, 0.0, 1.0);
        l[0] = 1.0 - t; l[1] = t;
    } else if (n == 3) {
        closest_tri_d(W[0], W[1], W[2], l[0], l[1], l[2]);
    } else {
        D3 ad = W[0] - W[3], bd = W[1] - W[3], cd = W[2] - W[3], od = d3(-W[3].x, -W[3].y, -W[3].z);
        double det = ddot(ad, dcross(bd, cd));
        double m1 = ddot(ad, ad), m2 = ddot(bd, bd), m3 = ddot(cd, cd);
        double L2 = m1 > m2 ? (m1 > m3 ? m1 : m3) : (m2 > m3 ? m2 : m3);
        bool flat = det * det <= 1e-24 * L2 * L2 * L2;
        double la = ddot(od, dcross(bd, cd)), lb = ddot(ad, dcross(od, cd)), lc = ddot(ad, dcross(bd, od));
        if (det < 0.0) { la = -la; lb = -lb; lc = -lc; det = -det; }
        double ld = det - la - lb - lc;
        if (!flat && la > 0.0 && lb > 0.0 && lc > 0.0 && ld > 0.0) return true;
        double best = 1e300, a, b, c;
        if (flat || ld <= 0.0) {
            closest_tri_d(W[0], W[1], W[2], a, b, c);
            D3 p = a * W[0] + b * W[1] + c * W[2]; double dd = ddot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = b; l[2] = c; l[3] = 0; }
        }
        if (flat || lb <= 0.0) {
            closest_tri_d(W[0], W[2], W[3], a, b, c);
            D3 p = a * W[0] + b * W[2] + c * W[3]; double dd = ddot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = 0; l[2] = b; l[3] = c; }
        }
        if (flat || lc <= 0.0) {
            closest_tri_d(W[0], W[3], W[1], a, b, c);
            D3 p = a * W[0] + b * W[3] + c * W[1]; double dd = ddot(p, p);
            if (dd < best) { best = dd; l[0] = a; l[1] = c; l[2] = 0; l[3] = b; }
        }
        if (flat || la <= 0.0) {
            closest_tri_d(W[1], W[3], W[2], a, b, c);
            D3 p = a * W[1] + b * W[3] + c * W[2]; double dd = ddot(p, p);
            if (dd < best) { best = dd; l[0] = 0; l[1] = a; l[2] = c; l[3] = b; }
        }
    }
    D3 nv = d3(0, 0, 0), T[4];
    int m = 0;
    for (int i = 0; i < 4; i++) {
        if (i < n && l[i] > 0.0) { nv = nv + l[i] * W[i]; T[m] = W[i]; m++; }
    }
    for (int i = 0; i < m; i++) W[i] = T[i];
    n = m; v = nv;
    return false;
}
#define URGYM_GJK_REFINE_ITER 256
#define URGYM_GJK_REFINE_GAP 2.0e-6f      /* FP32 result accepted when its distance is proven to this (lower bound known) */
template <class SA, class SB>
static URGYM_OOL float gjk_distance_refine(const SA A, const SB B, float3 v0) {     // shapes by value: the callers' stay in registers
    D3 W[4];
    int n = 1;
    float3 vf = v0;
    if (dot(vf, vf) < 1e-12f) vf = f3(1, 0, 0);
    D3 v = d3(A.support(-vf)) - d3(B.support(vf));
    W[0] = v;
    double vv = ddot(v, v);
    for (int it = 0; it < URGYM_GJK_REFINE_ITER; it++) {
        if (vv < 1e-24) return 0.0f;
        const float3 dir = f3((float)v.x, (float)v.y, (float)v.z);
        D3 w = d3(A.support(-dir)) - d3(B.support(dir));
        double delta = ddot(v, w);
        if (vv - delta <= 1e-9 * vv) break;
        bool dup = false;
        for (int i = 0; i < n; i++) { D3 e = w - W[i]; dup = dup || (ddot(e, e) <= 1e-24); }
        if (dup) break;
        W[n] = w; n++;
        D3 vn;
        if (closest_simplex_d(W, n, vn)) return 0.0f;
        double vvn = ddot(vn, vn);
        if (vvn >= vv) break;
        v = vn; vv = vvn;
    }
    return (float)sqrt(vv);
}

// ------------------------------------------------------------------------------------------------ closed forms
// squared distance between segments p1q1 and p2q2 (Ericson, RTCD 5.1.9)
URGYM_HD float segseg_dist2(float3 p1, float3 q1, float3 p2, float3 q2) {
    float3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    float s, t;
    const float EPS = 1e-12f;
    if (a <= EPS && e <= EPS) { return dot(r, r); }
    if (a <= EPS) { s = 0.0f; t = clampf(fdiv(f, e), 0.0f, 1.0f); }
    else {
        float c = dot(d1, r);
        if (e <= EPS) { t = 0.0f; s = clampf(fdiv(-c, a), 0.0f, 1.0f); }
        else {
            float b = dot(d1, d2), denom = a * e - b * b;
            s = denom > EPS * a * e ? clampf(fdiv(b * f - c * e, denom), 0.0f, 1.0f) : 0.0f;
            t = fdiv(b * s + f, e);
            if (t < 0.0f) { t = 0.0f; s = clampf(fdiv(-c, a), 0.0f, 1.0f); }
            else if (t > 1.0f) { t = 1.0f; s = clampf(fdiv(b - c, a), 0.0f, 1.0f); }
        }
    }
    float3 c1 = p1 + s * d1, c2 = p2 + t * d2, dd = c1 - c2;
    return dot(dd, dd);
}
// squared distance from point p to the axis-aligned box (c, he)
URGYM_HD float point_box_dist2(float3 p, float3 c, float3 he) {
    float ex = fmaxf(fabsf(p.x - c.x) - he.x, 0.0f), ey = fmaxf(fabsf(p.y - c.y) - he.y, 0.0f),
          ez = fmaxf(fabsf(p.z - c.z) - he.z, 0.0f);
    return ex * ex + ey * ey + ez * ez;
}
// lower bound of the distance between segment ab and box (c, he): distance between their AABBs
// squared distance between two NON-DEGENERATE segments, branch-free: clamp the unconstrained s, take the best t for
// it, clamp t, take the best s for that t (one round of exact coordinate descent past Ericson's case analysis, which
// it reproduces).  inv_a = 1/|q1-p1|^2 and inv_e = 1/|q2-p2|^2 are constants of the capsules.
URGYM_HD float segseg_dist2_fast(float3 p1, float3 q1, float3 p2, float3 q2, float inv_a, float inv_e) {
    float3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
    float denom = a * e - b * b;
    float s = denom > 1e-12f * a * e ? clampf(fdiv(b * f - c * e, denom), 0.0f, 1.0f) : 0.0f;
    float t = clampf((b * s + f) * inv_e, 0.0f, 1.0f);
    s = clampf((b * t - c) * inv_a, 0.0f, 1.0f);
    float3 dd = (r + s * d1) - t * d2;
    return dot(dd, dd);
}
// squared distance between the link segment a-b and the obstacle's axis segment (centre c, UNIT axis u, half length h):
// segseg_dist2_fast with the second segment parametrised by arc length about its centre (e = 1, no division by it)
URGYM_HD float seg_axis_dist2(float3 a, float3 b, float3 c, float3 u, float h, float inv_a) {
    const float3 d1 = b - a, r = a - c;
    const float A = dot(d1, d1), B = dot(d1, u), C = dot(d1, r), F = dot(u, r);
    const float denom = fmaf(-B, B, A);
    float s = denom > 1e-12f * A ? clampf(fdiv(fmaf(B, F, -C), denom), 0.0f, 1.0f) : 0.0f;
    const float t = clampf(fmaf(B, s, F), -h, h);
    s = clampf(fmaf(B, t, -C) * inv_a, 0.0f, 1.0f);
    const float3 dd = (r + s * d1) - t * u;
    return dot(dd, dd);
}
URGYM_HD float seg_box_lower2(float3 a, float3 b, float3 c, float3 he) {
    float gx = fmaxf(fmaxf(fminf(a.x, b.x) - (c.x + he.x), (c.x - he.x) - fmaxf(a.x, b.x)), 0.0f);
    float gy = fmaxf(fmaxf(fminf(a.y, b.y) - (c.y + he.y), (c.y - he.y) - fmaxf(a.y, b.y)), 0.0f);
    float gz = fmaxf(fmaxf(fminf(a.z, b.z) - (c.z + he.z), (c.z - he.z) - fmaxf(a.z, b.z)), 0.0f);
    return gx * gx + gy * gy + gz * gz;
}
// exact distance between segment ab and the box: f(t) = dist^2(a + t(b-a), box) is convex and C1; bisect f'.
URGYM_HD float seg_box_dist(float3 a, float3 b, float3 c, float3 he) {
    float3 d = b - a, a0 = a - c;
    float lo = 0.0f, hi = 1.0f;
    auto fprime = [&](float t) {
        float3 p = a0 + t * d;
        float ex = fmaxf(fabsf(p.x) - he.x, 0.0f), ey = fmaxf(fabsf(p.y) - he.y, 0.0f), ez = fmaxf(fabsf(p.z) - he.z, 0.0f);
        return copysignf(ex, p.x) * d.x + copysignf(ey, p.y) * d.y + copysignf(ez, p.z) * d.z;
    };
    if (fprime(0.0f) >= 0.0f) return sqrtf(point_box_dist2(a, c, he));
    if (fprime(1.0f) <= 0.0f) return sqrtf(point_box_dist2(b, c, he));
    for (int i = 0; i < 26; i++) {
        float m = 0.5f * (lo + hi);
        if (fprime(m) < 0.0f) lo = m; else hi = m;
    }
    float t = 0.5f * (lo + hi);
    return sqrtf(point_box_dist2(a + t * d, c, he));
}
// distance from point p to the cylinder core (c, unit axis u, r, h)
URGYM_HD float point_cyl_dist(float3 p, const CylW &C) {
    float3 w = p - C.c;
    float z = dot(w, C.u);
    float3 pr = w - z * C.u;
    float dr = fmaxf(sqrtf(dot(pr, pr)) - C.r, 0.0f), dz = fmaxf(fabsf(z) - C.h, 0.0f);
    return sqrtf(dr * dr + dz * dz);
}

}  // namespace urgym
