// urgym_env.cuh -- one environment's step / reset / observe, as the kernels run it (one env per thread).
//
// Reference semantics followed here (file:line relative to the UR-gym repository):
//   RobotTaskEnv.step / reset / _get_obs        UR_gym/envs/core.py:252-273,303-317
//   TimeLimit(max_episode_steps=100)            UR_gym/__init__.py:19-42
//   UR5Ori.set_action / get_obs / reset         UR_gym/envs/robots/UR5.py:273-279,304-332
//   ReachOri / ReachObs / ReachSta / ReachDyn   UR_gym/envs/tasks/reach.py:141-236,239-374,377-573,576-785
//   PyBullet.check_collision / get_link_distances / get_target_to_obstacle_distance
//                                               UR_gym/pyb_setup.py:382-456
// State words per env ("episode constants" E are rewritten only at reset):
//   q[6], elapsed (TimeLimit counter == ReachDyn.step_num), ep_return, link_dist[5] (== last_dist), E[EW]
//   E: Ori goal[6] | Obs goal[3] obstacle[6] | Sta goal[6] obstacle[6] obstacle_end[6] obstacle_start[6]
//      | Dyn goal[6] obstacle_start[6] obstacle_end[6]
//   (ReachSta.obstacle_end / obstacle_start are zero until an 18-value set_goal_and_obstacle sets them, reach.py:491-499,
//    and are NOT cleared by reset, reach.py:465-481: the moving-obstacle path of core.py:307-308)
#pragma once
#include "urgym_device.cuh"

namespace urgym {

enum { TASK_ORI = 0, TASK_OBS = 1, TASK_STA = 2, TASK_DYN = 3 };
// GEOM template values: bit 0 = link geometry, bit 1 = link-distance mode "workbench" (urgym_b200.h URGYM_LD_WORKBENCH).
// The mode is a compile-time property of its own kernel instantiations: the default kernels carry none of its code.
enum { GEOM_HULL = 0, GEOM_CAPSULE = 1, GEOM_WB = 2 };
#define URGYM_BASE(G) ((G) & 1)
#define URGYM_WB(G) (((G) & GEOM_WB) != 0)

#define URGYM_MAX_STEPS 100          /* UR_gym/__init__.py:22,28,34,41 */
#define URGYM_MAX_RESET_ITERS 256    /* rejection loop bound; P(reached) ~ 0.83^256 */
#define URGYM_COLLISION_MARGIN 0.01f /* getClosestPoints(distance=0.01)  pyb_setup.py:401,410,421 */
#define URGYM_DT_ENV 0.04f           /* 20 substeps x 1/500 s            pyb_setup.py:25,40,50 */

// EH: leading words of E the step reads; CW: words of the episode cache C (derive_cache), quantities that depend on
// the episode constants only and are therefore computed once per episode, at reset, instead of once per step:
//   Ori  C = goal quaternion[4]                                  (utils.py:48-54 applied to the goal)
//   Obs  C = obstacle axis[3]
//   Sta  C = goal quaternion[4], m, then  m == 0 (obstacle at rest): axis[3], Euler read-back[3]   (reach.py:455-457)
//                                         m  > 0 (moving for m steps):  start quaternion[4], twist[6]  (reach.py:518-541)
//   Dyn  C = goal quaternion[4], start quaternion[4], ReachDyn.velocity[6]       (reach.py:728-753)
// The step kernel reads the "hot" words H = E[0..EH) ++ C[0..CW) from their own planes.
template <int TASK> struct Traits;
template <> struct Traits<TASK_ORI> {
    static constexpr int OBS = 18, GOAL = 6, EW = 6, BPI = 2, OBST = -1, EH = 6, CW = 4;
    static constexpr bool HAS_OBST = false, ORI = true, DYN = false;
};
template <> struct Traits<TASK_OBS> {
    static constexpr int OBS = 26, GOAL = 3, EW = 9, BPI = 3, OBST = 3, EH = 9, CW = 3;
    static constexpr bool HAS_OBST = true, ORI = false, DYN = false;
};
template <> struct Traits<TASK_STA> {
    static constexpr int OBS = 29, GOAL = 6, EW = 24, BPI = 3, OBST = 6, EH = 9, CW = 15;
    static constexpr bool HAS_OBST = true, ORI = true, DYN = false;
};
template <> struct Traits<TASK_DYN> {
    static constexpr int OBS = 35, GOAL = 6, EW = 18, BPI = 5, OBST = 6, EH = 9, CW = 14;   // OBST = obstacle_start; end = OBST + 6
    static constexpr bool HAS_OBST = true, ORI = true, DYN = true;
};

// goal / obstacle sampling boxes   reach.py:151-152, 248-251, 385-388, 584-587
template <int TASK> URGYM_HD void goal_range(float3 &lo, float3 &hi) {
    if (TASK == TASK_OBS) { lo = f3(0.3f, -0.5f, -0.1f); hi = f3(0.75f, 0.5f, 0.2f); }
    else if (TASK == TASK_DYN) { lo = f3(0.4f, -0.5f, 0.0f); hi = f3(0.75f, 0.5f, 0.2f); }
    else { lo = f3(0.3f, -0.5f, 0.0f); hi = f3(0.75f, 0.5f, 0.2f); }
}
template <int TASK> URGYM_HD void obstacle_range(float3 &lo, float3 &hi) {
    if (TASK == TASK_DYN) { lo = f3(0.5f, -0.8f, 0.25f); hi = f3(1.2f, 0.8f, 0.75f); }
    else { lo = f3(0.5f, -0.5f, 0.25f); hi = f3(1.0f, 0.5f, 0.55f); }
}

struct EnvState {
    float q[6];
    int elapsed;
    float ep_ret;
    float ld[5];
    float E[24];
    float C[15];        // episode cache (derive_cache)
};

struct StepOut {
    float reward;
    bool terminated, truncated, success, collision;
};

// ------------------------------------------------------------------------------------------------ obstacle pose
struct ObstW {          // obstacle in the world: centre, orientation, unit axis (local z)
    float3 c;
    Quat q;
    float3 u;
};
URGYM_HD ObstW obstacle_none() {
    ObstW O;
    O.c = f3(0, 0, 0); O.q.x = O.q.y = O.q.z = 0.0f; O.q.w = 1.0f; O.u = f3(0, 0, 1);
    return O;
}
URGYM_HD float3 quat_axis_z(Quat q) {     // third column of the rotation matrix of a unit quaternion
    return f3(2.0f * (q.x * q.z + q.w * q.y), 2.0f * (q.y * q.z - q.w * q.x), 1.0f - 2.0f * (q.x * q.x + q.y * q.y));
}
URGYM_HD ObstW obstacle_static(const float *o) {      // set_base_pose(position, euler)   pyb_setup.py:305-317
    ObstW O;
    O.c = f3(o[0], o[1], o[2]);
    O.q = quat_from_euler(o[3], o[4], o[5]);
    O.u = quat_axis_z(O.q);
    return O;
}
// ReachDyn.set_velocity (reach.py:728-753) + the kinematic motion of a mass-0 base under resetBaseVelocity over
// the substeps (pyb_setup.py:52-55,340-349).  The twist is constant while step_num < 25, so the pose after `moved`
// env steps is closed form: translation moved*0.04*v, rotation by moved*0.04*|w| about w (Bullet composes one fixed
// world-frame increment per substep).  vel = [v, w] as ReachDyn.velocity holds it.
URGYM_HD void dyn_twist(const float *start, const float *end, float *vel, Quat &qs, float3 &axis, float &angle,
                        float inv_duration = 0.5f) {
    qs = quat_from_euler(start[3], start[4], start[5]);
    Quat qe = quat_from_euler(end[3], end[4], end[5]);
    // getDifferenceQuaternion(start, end) = nearest(end) * start^-1                  pyb_setup.py:351-359
    float dm = (qs.x - qe.x) * (qs.x - qe.x) + (qs.y - qe.y) * (qs.y - qe.y) + (qs.z - qe.z) * (qs.z - qe.z) +
               (qs.w - qe.w) * (qs.w - qe.w);
    float dp = (qs.x + qe.x) * (qs.x + qe.x) + (qs.y + qe.y) * (qs.y + qe.y) + (qs.z + qe.z) * (qs.z + qe.z) +
               (qs.w + qe.w) * (qs.w + qe.w);
    if (!(dm < dp)) { qe.x = -qe.x; qe.y = -qe.y; qe.z = -qe.z; qe.w = -qe.w; }
    Quat si; si.x = -qs.x; si.y = -qs.y; si.z = -qs.z; si.w = qs.w;
    Quat d = quat_mul(qe, si);
    // getAxisAngleFromQuaternion: angle = 2 acos(w), axis = xyz / sqrt(1 - w^2), (1,0,0) if degenerate  pyb_setup.py:361-363
    float w = clampf(d.w, -1.0f, 1.0f);
    float s2 = 1.0f - d.w * d.w;
    float vn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);      // = sqrt(1 - w^2) for a unit quaternion, better conditioned
    angle = 2.0f * atan2_fast(vn, w);                             // = 2 acos(w)
    if (s2 < 10.0f * 1.1920929e-7f || vn == 0.0f) axis = f3(1.0f, 0.0f, 0.0f);
    else axis = (1.0f / vn) * f3(d.x, d.y, d.z);
    vel[0] = (end[0] - start[0]) * inv_duration; vel[1] = (end[1] - start[1]) * inv_duration; vel[2] = (end[2] - start[2]) * inv_duration;
    vel[3] = axis.x * angle * inv_duration; vel[4] = axis.y * angle * inv_duration; vel[5] = axis.z * angle * inv_duration;
}
// episode cache: everything the step needs that depends on the episode constants E only (see Traits)
template <int TASK> URGYM_HD void derive_cache(const float *E, float *C) {
    typedef Traits<TASK> TT;
    if (TT::ORI) {                      // angular_distance's quaternion of the goal          utils.py:48-54
        Quat g = quat_ZYX(E[3], E[4], E[5]);
        C[0] = g.x; C[1] = g.y; C[2] = g.z; C[3] = g.w;
    }
    if (TASK == TASK_OBS) {
        ObstW O = obstacle_static(&E[3]);
        C[0] = O.u.x; C[1] = O.u.y; C[2] = O.u.z;
    } else if (TASK == TASK_STA) {
        // ReachSta.set_velocity runs when obstacle_end is not all zero (core.py:307-308).  While the obstacle is farther
        // than 0.05 m from obstacle_end[:3] it moves with the twist (end - start) / 1 s (reach.py:518-541); the test is
        // made before every step, so the number of moving steps m follows from the episode constants alone.
        bool armed = false;
#pragma unroll
        for (int k = 0; k < 6; k++) armed = armed || (E[12 + k] != 0.0f);
        int m = 0;
        float tw[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        Quat qs; qs.x = qs.y = qs.z = 0.0f; qs.w = 1.0f;
        if (armed) {
            float3 axis; float angle;
            dyn_twist(&E[18], &E[12], tw, qs, axis, angle, 1.0f);       // twist from obstacle_START to obstacle_end
            m = URGYM_MAX_STEPS;
            for (int k = 0; k < URGYM_MAX_STEPS; k++) {
                const float t = (float)k * URGYM_DT_ENV;
                const float dx = E[12] - fmaf(t, tw[0], E[6]), dy = E[13] - fmaf(t, tw[1], E[7]), dz = E[14] - fmaf(t, tw[2], E[8]);
                if (!(sqrtf(dx * dx + dy * dy + dz * dz) > 0.05f)) { m = k; break; }
            }
        }
        C[4] = (float)m;
        ObstW O = obstacle_static(&E[6]);       // orientation of the obstacle as placed (its own Euler triple)
        if (m == 0) {                           // at rest: axis and the pose read-back of get_obs   reach.py:455-457
            float3 e = euler_from_quat(O.q);
            C[5] = O.u.x; C[6] = O.u.y; C[7] = O.u.z; C[8] = e.x; C[9] = e.y; C[10] = e.z;
            C[11] = C[12] = C[13] = C[14] = 0.0f;
        } else {
            C[5] = O.q.x; C[6] = O.q.y; C[7] = O.q.z; C[8] = O.q.w;
#pragma unroll
            for (int k = 0; k < 6; k++) C[9 + k] = tw[k];
        }
    } else if (TASK == TASK_DYN) {      // ReachDyn.set_velocity's twist is the same at every step of an episode
        Quat qs; float3 axis; float angle; float tw[6];
        dyn_twist(&E[6], &E[12], tw, qs, axis, angle);
        C[4] = qs.x; C[5] = qs.y; C[6] = qs.z; C[7] = qs.w;
#pragma unroll
        for (int k = 0; k < 6; k++) C[8 + k] = tw[k];
    }
}
// obstacle pose after `moved` env steps of motion and its Euler read-back (get_base_rotation), from E[0..EH) and C.
// Dyn: translation moved*0.04*v, rotation by moved*0.04*|w| about w = axis*angle/2 (Bullet composes one fixed
// world-frame increment per substep), so the half angle is 0.5*t*|w| and sin(half)*axis = w * sin(half)/|w|.
// pose after `moved` env steps of motion with twist tw = [v, w] from the pose (p0, qs)
URGYM_HD void twist_pose(const float *p0, Quat qs, const float *tw, int moved, ObstW &O, float3 &euler) {
    const float t = (float)moved * URGYM_DT_ENV;
    O.c = f3(fmaf(t, tw[0], p0[0]), fmaf(t, tw[1], p0[1]), fmaf(t, tw[2], p0[2]));
    const float3 w = f3(tw[3], tw[4], tw[5]);
    const float wn2 = dot(w, w);
    const float inv = wn2 > 0.0f ? rsqrt_f(wn2) : 0.0f;
    float sn, cs;
    sincos_fast(0.5f * t * (wn2 * inv), &sn, &cs);
    const float k = sn * inv;
    Quat r; r.x = w.x * k; r.y = w.y * k; r.z = w.z * k; r.w = cs;
    O.q = quat_mul(r, qs);                              // world-frame increment on the left
    O.u = quat_axis_z(O.q);
    euler = euler_from_quat(O.q);
}
// ReachSta with a moving obstacle (injection-only path): rare.  Out of line, and called with VALUES: handing it pointers
// into the env's E / C words would force those arrays into local memory for the whole step kernel.
struct StaMotion { float p0[3], tw[6]; Quat qs; int m; };
static URGYM_OOL void sta_moving_pose(StaMotion mo, int steps, ObstW &O, float3 &euler) {
    twist_pose(mo.p0, mo.qs, mo.tw, steps < mo.m ? steps : mo.m, O, euler);
}
// `steps`: env steps of the episode the obstacle has been through (capped inside: Dyn moves for 25, reach.py:735)
template <int TASK> URGYM_HD ObstW obstacle_cached(const float *E, const float *C, int steps, float3 &euler) {
    ObstW O = obstacle_none();
    euler = f3(0, 0, 0);
    if (TASK == TASK_OBS) {
        O.c = f3(E[3], E[4], E[5]); O.u = f3(C[0], C[1], C[2]);
        euler = f3(E[6], E[7], E[8]);                   // Obs shows the Euler triple as sampled (quirk Q3)
    } else if (TASK == TASK_STA) {
        if (C[4] == 0.0f) {
            O.c = f3(E[6], E[7], E[8]); O.u = f3(C[5], C[6], C[7]);
            euler = f3(C[8], C[9], C[10]);
        } else {
            StaMotion mo;
            mo.p0[0] = E[6]; mo.p0[1] = E[7]; mo.p0[2] = E[8];
            mo.qs.x = C[5]; mo.qs.y = C[6]; mo.qs.z = C[7]; mo.qs.w = C[8];
#pragma unroll
            for (int k = 0; k < 6; k++) mo.tw[k] = C[9 + k];
            mo.m = (int)C[4];
            ObstW Om; float3 em;
            sta_moving_pose(mo, steps, Om, em);
            O = Om; euler = em;
        }
    } else if (TASK == TASK_DYN) {
        const int moved = steps < 25 ? steps : 25;
        const float t = (float)moved * URGYM_DT_ENV;
        O.c = f3(fmaf(t, C[8], E[6]), fmaf(t, C[9], E[7]), fmaf(t, C[10], E[8]));
        const float3 w = f3(C[11], C[12], C[13]);
        const float wn2 = dot(w, w);
        const float inv = wn2 > 0.0f ? rsqrt_f(wn2) : 0.0f;
        float sn, cs;
        sincos_small(0.5f * t * (wn2 * inv), &sn, &cs);      // t |w| <= the angle between two orientations <= pi
        const float k = sn * inv;
        Quat r; r.x = w.x * k; r.y = w.y * k; r.z = w.z * k; r.w = cs;
        Quat qs; qs.x = C[4]; qs.y = C[5]; qs.z = C[6]; qs.w = C[7];
        O.q = quat_mul(r, qs);                          // world-frame increment on the left
        O.u = quat_axis_z(O.q);
        euler = euler_from_quat(O.q);
    }
    return O;
}

// ------------------------------------------------------------------------------------------------ robot geometry
// exact squared distance between segment ab and the axis-aligned box (c, he).  g(t) = 1/2 d/dt dist^2 is monotone
// and piecewise linear with breakpoints where the point crosses a slab face; bracket the root between breakpoints.
// (out of line: it runs only for links that come close to the table or the track)
template <bool ROLLED>
static URGYM_OOL float seg_box_dist2_t(float3 a, float3 b, float3 c, float3 he) {
    float3 p0 = a - c, d = b - a;
    auto ex = [](float p, float h) { return p - clampf(p, -h, h); };
    auto g = [&](float t) {
        return ex(fmaf(t, d.x, p0.x), he.x) * d.x + ex(fmaf(t, d.y, p0.y), he.y) * d.y + ex(fmaf(t, d.z, p0.z), he.z) * d.z;
    };
    auto d2 = [&](float t) {
        float x = ex(fmaf(t, d.x, p0.x), he.x), y = ex(fmaf(t, d.y, p0.y), he.y), z = ex(fmaf(t, d.z, p0.z), he.z);
        return x * x + y * y + z * z;
    };
    float lo = 0.0f, hi = 1.0f, glo = g(0.0f), ghi = g(1.0f);
    if (glo >= 0.0f) return d2(0.0f);
    if (ghi <= 0.0f) return d2(1.0f);
    // The six slab-face crossings.  ROLLED: a rolled loop -- the routine runs for one or two lanes of a warp, and unrolled its
    // 250 instructions pushed the UR5StaReach / UR5DynReach step kernels past the 32 KB instruction cache (DESIGN.md section
    // 5: misses 582 k -> 133 k per launch, Sta -3.3 %, Dyn -1.3 %); the smaller UR5OriReach / UR5ObsReach kernels keep the
    // unrolled form (its dependent chain is shorter: +1 % rolled).
#pragma unroll (ROLLED ? 1 : 6)
    for (int k = 0; k < 6; k++) {
        const int i = k >> 1;
        const float pi = i == 0 ? p0.x : (i == 1 ? p0.y : p0.z), di = i == 0 ? d.x : (i == 1 ? d.y : d.z);
        const float hi_ = i == 0 ? he.x : (i == 1 ? he.y : he.z);
        if (di != 0.0f) {
            const float t = (((k & 1) ? hi_ : -hi_) - pi) * fdiv(1.0f, di);
            if (t > lo && t < hi) {
                const float gt = g(t);
                if (gt < 0.0f) { lo = t; glo = gt; } else { hi = t; ghi = gt; }
            }
        }
    }
    float t = (ghi > glo) ? lo + (hi - lo) * fdiv(-glo, ghi - glo) : lo;
    return d2(t);
}
static URGYM_OOL float seg_box_dist2(float3 a, float3 b, float3 c, float3 he) { return seg_box_dist2_t<true>(a, b, c, he); }
URGYM_HD float point_seg_dist2(float3 p, float3 a, float3 b) {
    float3 ab = b - a, ap = p - a;
    float den = dot(ab, ab);
    float t = den > 0.0f ? clampf(fdiv(dot(ap, ab), den), 0.0f, 1.0f) : 0.0f;
    float3 e = ap - t * ab;
    return dot(e, e);
}

// Collision shape of ONE link in the world.  The step walks the chain once (a rolled loop over the links: the code
// of one iteration stays resident in the instruction cache) and tests each link as soon as its pose is known, so only
// the shapes of links 1..3 (the first members of the self-collision pairs) are kept.
template <int GEOM> struct LinkShape;

template <> struct LinkShape<GEOM_CAPSULE> {
    float3 a, b;            // capsule segment
    URGYM_HD void set(const ModelConst &M, int l, const Pose &T, const float4 *) {
        a = T.p + rot(T.R, f3(M.cap_p0[l][0], M.cap_p0[l][1], M.cap_p0[l][2]));
        b = T.p + rot(T.R, f3(M.cap_p1[l][0], M.cap_p1[l][1], M.cap_p1[l][2]));
    }
    URGYM_HD void set_neutral(const ModelConst &M, int l, const float4 *) {
        a = f3(M.neutral_ca[l][0], M.neutral_ca[l][1], M.neutral_ca[l][2]);
        b = f3(M.neutral_cb[l][0], M.neutral_cb[l][1], M.neutral_cb[l][2]);
    }
    // getClosestPoints(UR5, obstacle, linkIndexA=l)[0][8], capsule geometry                 pyb_setup.py:439-456
    URGYM_HD float obstacle_dist(const ModelConst &M, int l, const ObstW &O) const {
        return sqrtf(seg_axis_dist2(a, b, O.c, O.u, M.fit_obst_h, M.cap_ia[l])) - M.fit_obst[l];     // as in robot_pass_capsule
    }
    // getClosestPoints(UR5, table | track, linkIndexA=l)[0][8], capsule geometry (link-distance mode "workbench")
    URGYM_HD float box_dist(const ModelConst &M, int l, int box) const {
        float3 bc = f3(M.box_c[box][0], M.box_c[box][1], M.box_c[box][2]);
        float3 bh = f3(M.box_he[box][0], M.box_he[box][1], M.box_he[box][2]);
        return sqrtf(seg_box_dist2(a, b, bc, bh)) - M.fit_box[l] - M.box_margin[box];
    }
    // the methods below use the BOUNDING capsules: they are the exact-safe broad phase of the hull geometry
    // box 0 = table, 1 = track
    URGYM_HD bool box_hit(const ModelConst &M, int l, int box) const {
        float margin = M.box_margin[box];
        float reach = URGYM_COLLISION_MARGIN + M.cap_m[l] + margin;
        float3 bc = f3(M.box_c[box][0], M.box_c[box][1], M.box_c[box][2]);
        float3 bh = f3(M.box_he[box][0], M.box_he[box][1], M.box_he[box][2]);
        if (seg_box_lower2(a, b, bc, bh) > reach * reach) return false;      // broad phase (exact bound)
        return sqrtf(seg_box_dist2(a, b, bc, bh)) - M.cap_m[l] - margin <= URGYM_COLLISION_MARGIN;
    }
    URGYM_HD bool link_hit(const ModelConst &M, int l, int l2, const LinkShape &o) const {
        float reach = URGYM_COLLISION_MARGIN + M.cap_m[l] + M.cap_m[l2];
        // broad phase (exact bound): segments live inside the spheres about their midpoints
        float3 dm = 0.5f * ((a + b) - (o.a + o.b));
        float far = reach + M.cap_hl[l] + M.cap_hl[l2];
        if (dot(dm, dm) > far * far) return false;
        return segseg_dist2(a, b, o.a, o.b) <= reach * reach;
    }
};

// the reference's geometry: convex hulls of the collision meshes, cylinder obstacle, GJK with Bullet's margins.
// The bounding capsules serve as an exact-safe broad phase for the collision booleans (capsule distance is a lower
// bound of the hull distance); the hull GJK runs only where the capsule bound cannot decide.
template <> struct LinkShape<GEOM_HULL> {
    Pose T;
    LinkShape<GEOM_CAPSULE> cap;
    const float4 *hv;       // the hull blob: packed vertices + adjacency (shared memory on the device)
    URGYM_HD void set(const ModelConst &M, int l, const Pose &P, const float4 *verts) { T = P; hv = verts; cap.set(M, l, P, verts); }
    URGYM_HD void set_neutral(const ModelConst &M, int l, const float4 *verts) {
        hv = verts;
        cap.set_neutral(M, l, verts);
#pragma unroll
        for (int k = 0; k < 9; k++) T.R[k] = M.neutral_R[l][k];
        T.p = f3(M.neutral_p[l][0], M.neutral_p[l][1], M.neutral_p[l][2]);
    }
    URGYM_HD HullW hull(const ModelConst &M, int l) const {
        HullW H; H.v = hv + M.hull_off[l]; H.n = M.hull_off[l + 1] - M.hull_off[l]; H.T = T;
        H.aoff = hull_adj_off(hv) + M.hull_off[l]; H.adj = hull_adj(hv); H.start = hull_dirmap(hv) + l * URGYM_HULL_DIR_CELLS; H.cur = -1;
        return H;
    }
    // link <-> obstacle distance in two stages (the hull-geometry step kernel runs the second one as a compacted task):
    // GJK converges only sublinearly against the curved wall of a cylinder (hundreds of iterations for 1e-6).
    // Against the cylinder's AXIS SEGMENT the problem is polytope-polytope and terminates in a few iterations;
    // when the closest vector comes out perpendicular to the axis, the nearest cylinder point lies on the side
    // wall and distance(hull, cylinder) = distance(hull, axis segment) - radius exactly.
    URGYM_HD bool obstacle_dist_side(const ModelConst &M, int l, const ObstW &O, float &d) const {
        bool deep;
        float3 v;
        float lb;
        SegW S; S.a = O.c - M.obst_h * O.u; S.b = O.c + M.obst_h * O.u;
        const HullW H = hull(M, l);
        float ds = gjk_distance(H, S, deep, &v, &lb);
        if (!deep && ds > M.obst_r && fabsf(dot(v, O.u)) <= 1e-5f * ds) {
            if (ds - lb > URGYM_GJK_REFINE_GAP) ds = gjk_distance_refine(H, S, v);
            d = ds - M.obst_r - M.hull_margin - M.obst_margin;
            return true;
        }
        return false;
    }
    URGYM_HD float obstacle_dist_caps(const ModelConst &M, int l, const ObstW &O) const {      // end caps / rims / penetration
        bool deep;
        float3 v;
        float lb;
        const HullW H = hull(M, l);
        CylW C; C.c = O.c; C.u = O.u; C.r = M.obst_r; C.h = M.obst_h;
        float d = gjk_distance(H, C, deep, &v, &lb);
        if (!deep && d - lb > URGYM_GJK_REFINE_GAP) d = gjk_distance_refine(H, C, v);
        return d - M.hull_margin - M.obst_margin;
    }
    URGYM_HD float obstacle_dist(const ModelConst &M, int l, const ObstW &O) const {
        float d;
        if (obstacle_dist_side(M, l, O, d)) return d;
        return obstacle_dist_caps(M, l, O);
    }
    URGYM_HD float box_dist(const ModelConst &M, int l, int box) const {      // link-distance mode "workbench"
        BoxA B; B.c = f3(M.box_c[box][0], M.box_c[box][1], M.box_c[box][2]);
        B.he = f3(M.box_he[box][0], M.box_he[box][1], M.box_he[box][2]);
        bool deep; float3 v; float lb;
        const HullW H = hull(M, l);
        float d = gjk_distance(H, B, deep, &v, &lb);
        if (!deep && d - lb > URGYM_GJK_REFINE_GAP) d = gjk_distance_refine(H, B, v);
        return d - M.hull_margin - M.box_margin[box];
    }
    // the exact halves of the collision booleans: run where the capsule broad phase (cap.box_hit / cap.link_hit) cannot decide
    URGYM_HD bool box_hit_exact(const ModelConst &M, int l, int box) const {
        BoxA B; B.c = f3(M.box_c[box][0], M.box_c[box][1], M.box_c[box][2]);
        B.he = f3(M.box_he[box][0], M.box_he[box][1], M.box_he[box][2]);
        bool deep; float3 v; float lb;
        const HullW H = hull(M, l);
        float d = gjk_distance(H, B, deep, &v, &lb);
        const float thr = URGYM_COLLISION_MARGIN + M.hull_margin + M.box_margin[box];
        if (!deep && d - lb > URGYM_GJK_REFINE_GAP && fabsf(d - thr) < 1e-3f) d = gjk_distance_refine(H, B, v);   // undecided near the threshold
        return d <= thr;
    }
    URGYM_HD bool link_hit_exact(const ModelConst &M, int l, int l2, const LinkShape &o) const {
        bool deep; float3 v; float lb;
        const HullW H1 = hull(M, l), H2 = o.hull(M, l2);
        float d = gjk_distance(H1, H2, deep, &v, &lb);
        const float thr = URGYM_COLLISION_MARGIN + 2.0f * M.hull_margin;
        if (!deep && d - lb > URGYM_GJK_REFINE_GAP && fabsf(d - thr) < 1e-3f) d = gjk_distance_refine(H1, H2, v);
        return d <= thr;
    }
    URGYM_HD bool box_hit(const ModelConst &M, int l, int box) const { return cap.box_hit(M, l, box) && box_hit_exact(M, l, box); }
    URGYM_HD bool link_hit(const ModelConst &M, int l, int l2, const LinkShape &o) const {
        return cap.link_hit(M, l, l2, o.cap) && link_hit_exact(M, l, l2, o);
    }
};

// getClosestPoints(target, obstacle, distance=5)[0][8]                                  pyb_setup.py:431-437
template <int TASK, int GEOM>
URGYM_HD float target_obstacle_dist(const ModelConst &M, const float *goal, const ObstW &O) {
    float3 g = f3(goal[0], goal[1], goal[2]);
    if (URGYM_BASE(GEOM) == GEOM_CAPSULE) {
        float3 oa = O.c - M.fit_obst_h * O.u, ob = O.c + M.fit_obst_h * O.u;
        return sqrtf(point_seg_dist2(g, oa, ob)) - M.tgt_cap_m[TASK] - M.obst_cap_m;
    }
    CylW C; C.c = O.c; C.u = O.u; C.r = M.obst_r; C.h = M.obst_h;
    if (TASK == TASK_OBS) return point_cyl_dist(g, C) - M.tgt_sphere_margin - M.obst_margin;   // sphere r 0.02  reach.py:270-277
    BoxO B; B.c = g; B.he = M.tgt_box_he;                                                      // box he 0.025  reach.py:418-426
    mat_from_quat(quat_from_euler(goal[3], goal[4], goal[5]), B.R);
    bool deep;
    float d = gjk_distance(B, C, deep);
    return d - M.tgt_box_margin - M.obst_margin;
}

// Forward kinematics + PyBullet.check_collision (pyb_setup.py:382-429) + get_link_distances (pyb_setup.py:439-456) in
// one pass over the chain.  Same pairs as the reference (the early-out order does not change the boolean):
//   links 2..6 vs obstacle (when keys[5] == 'obstacle'), links 2..6 vs table and track,
//   self pairs (1:3,4,5,6) (2:4,5,6) (3:5,6).
// ee[6] = EE position + PyBullet Euler triple (UR5.py:320-325,334-340); d0..d4 = link 2..6 <-> obstacle distances.
// q is read through `qrow` (memory, e.g. the joint columns of the observation row) so the loop can stay rolled.
template <int TASK, int GEOM>
URGYM_HD bool robot_pass_rolled(const ModelConst &M, const float *qrow, const ObstW &O, const float4 *hv, bool collide,
                         float *ee, float &d0, float &d1, float &d2, float &d3, float &d4) {
    Pose T;
    pose_identity(T);
    LinkShape<URGYM_BASE(GEOM)> s1, s2, s3, cur;
    bool hit = false;
#pragma unroll 1
    for (int l = 1; l < 7; l++) {
        fk_advance(M, T, l - 1, qrow[l - 1]);
        if (collide) {
            cur.set(M, l, T, hv);
            if (l >= 2) {
                if (Traits<TASK>::HAS_OBST) {       // keys[5] == 'obstacle'   pyb_setup.py:398-399
                    float d = cur.obstacle_dist(M, l, O);
                    hit = hit || (d <= URGYM_COLLISION_MARGIN);
                    if (URGYM_WB(GEOM)) d = fminf(d, fminf(cur.box_dist(M, l, 0), cur.box_dist(M, l, 1)));
                    if (l == 2) d0 = d; else if (l == 3) d1 = d; else if (l == 4) d2 = d; else if (l == 5) d3 = d; else d4 = d;
                }
#pragma unroll 1
                for (int box = 0; box < 2; box++) hit = hit || cur.box_hit(M, l, box);
            }
            if (l >= 3) hit = hit || cur.link_hit(M, l, 1, s1);
            if (l >= 4) hit = hit || cur.link_hit(M, l, 2, s2);
            if (l >= 5) hit = hit || cur.link_hit(M, l, 3, s3);
            if (l == 1) s1 = cur; else if (l == 2) s2 = cur; else if (l == 3) s3 = cur;
        }
    }
    float3 e = euler_from_mat(T.R);
    ee[0] = T.p.x; ee[1] = T.p.y; ee[2] = T.p.z; ee[3] = e.x; ee[4] = e.y; ee[5] = e.z;
    return hit;
}

// The capsule-geometry version of the same pass.  Phase 1, fully unrolled (the per-joint constants become immediate
// constant-bank operands): FK and the world capsule segment of every link, parked in a per-thread scratch column
// (`cap[k * cs]`: shared memory on the device, conflict-free with cs = 32).  Phase 2, rolled loops over links and
// link pairs, so each distance routine exists once in the instruction stream.
// scratch layout: 36 floats capsule endpoints (link 1..6: a.xyz b.xyz), 5 floats link-obstacle distances.
#define URGYM_SCRATCH_FLOATS 41
// link-distance mode "workbench": fold the link's distances to the table and the track into the link-obstacle
// distances of the scratch column
static URGYM_OOL void workbench_link_dist(const ModelConst &M, float *cap, int cs) {
#pragma unroll 1
    for (int l = 2; l < 7; l++) {
        const float *c = cap + (l - 1) * 6 * cs;
        LinkShape<GEOM_CAPSULE> L;
        L.a = f3(c[0], c[cs], c[2 * cs]); L.b = f3(c[3 * cs], c[4 * cs], c[5 * cs]);
        float &d = cap[(36 + l - 2) * cs];
        d = fminf(d, fminf(L.box_dist(M, l, 0), L.box_dist(M, l, 1)));
    }
}
// Table and track (pyb_setup.py:408-417) for ONE link with its end points in registers, branch-free for the common
// cases.  With L the lower end point of the segment and dz = L.z - (top face of the box core): dz > reach means the whole
// segment is farther than `reach` above the box, no hit; otherwise, if L lies over the box's footprint, dist(segment,
// box) <= dist(L, box) = max(dz, 0) <= reach, a hit.  Only a low segment whose lower end is beside the footprint (table
// edges, most track cases) needs the exact segment-box distance; those (link, box) pairs are collected in `slow`.
// l is a compile-time value at every call site (the unrolled chain walk), so the per-(link, box) limits are immediate
// constant-bank operands: done in a rolled loop over the links they cost one LDC each, a third of that loop's 340
// instructions per warp.
URGYM_HD void box_tests(const ModelConst &M, int l, float3 a, float3 b, bool &hit, unsigned &slow) {
    const bool a_low = a.z <= b.z;
    const float Lx = a_low ? a.x : b.x, Ly = a_low ? a.y : b.y, zmin = fminf(a.z, b.z);
    const float minx = fminf(a.x, b.x), maxx = fmaxf(a.x, b.x), miny = fminf(a.y, b.y), maxy = fmaxf(a.y, b.y);
    // Upper arm only: in steady state three quarters of the `slow` cases are the upper arm hanging down beside the track
    // or behind the table, and 87 % of those are no hits.  The axis n = d x e_y (perpendicular to the segment and to the
    // boxes' long edges) separates every one of them: the segment projects to a point, the box to |n_x| hx + |n_z| hz,
    // and a gap beyond `reach` is a proven miss.
    bool apart[2] = {false, false};
    if (l == 2) {
        const float dx = b.x - a.x, dz = b.z - a.z, mx = 0.5f * (a.x + b.x), mz = 0.5f * (a.z + b.z);
        const float nn = sqrtf(fmaf(dx, dx, dz * dz));
#pragma unroll
        for (int box = 0; box < 2; box++) {
            const float *S = M.sat2[box];           // cx, cz, hx, hz, reach
            const float gap = fabsf(fmaf(dx, mz - S[1], -dz * (mx - S[0]))) - fmaf(S[2], fabsf(dz), S[3] * fabsf(dx));
            apart[box] = gap > S[4] * nn;
        }
    }
    // The reach (collision margin + capsule radius + box margin) depends on the link only, the boxes on nothing: the link's
    // extents are grown by the reach and compared with per-box constants that stay in uniform registers for all links.
    const float r = M.box_reach[l];
    const float zr = zmin - r, xlo = minx - r, xhi = maxx + r, ylo = miny - r, yhi = maxy + r;
#pragma unroll
    for (int box = 0; box < 2; box++) {
        const float *K = M.box_k[box];              // ztop, cx, hx, hy, cx - hx, cx + hx, -hy, hy   (cy = 0)
        const bool near_z = zr <= K[0];
        const bool over = fabsf(Lx - K[1]) <= K[2] && fabsf(Ly) <= K[3];
        const bool beside = xhi >= K[4] && xlo <= K[5] && yhi >= K[6] && ylo <= K[7];
        hit = hit || (near_z && over);
        if (near_z && !over && beside && !apart[box]) slow |= 1u << (2 * (l - 2) + box);
    }
}
template <int TASK, bool WORKBENCH>
URGYM_HD bool robot_pass_capsule(const ModelConst &M, const float *q, const ObstW &O, bool collide, float *ee,
                                 float *dist, float *cap, int cs) {
    // Where the table / track tests run.  In the chain walk (end points still in registers, no link loop at all for
    // UR5OriReach) they cost the fewest instructions: Ori -11 %, Obs -2.7 %, Sta -1 % kernel time.  UR5DynReach keeps them
    // in the link loop: its walk already carries the most live state, and the 300 extra static instructions pushed its
    // instruction-fetch stalls from 0.66 to 1.17 warps per issue slot (+1.5 %, measured with 3.5 % fewer instructions).
#ifdef URGYM_BOX_IN_LOOP
    constexpr bool BOX_IN_LOOP = URGYM_BOX_IN_LOOP != 0;
#else
    constexpr bool BOX_IN_LOOP = TASK == TASK_DYN;
#endif
    PoseP T;
    posep_identity(T);
    bool hit = false;
    unsigned slow = 0u;
#pragma unroll
    for (int l = 1; l < 7; l++) {
        if (l == 1) fkp_first(M, T, q[0]);          // from the identity pose: a third of the general step
        else fkp_advance(M, T, l - 1, q[l - 1]);
        if (collide) {
            float2 axy, bxy, abz;
            capsule_world(M, T, l, axy, bxy, abz);
            float *c = cap + (l - 1) * 6 * cs;
            c[0] = axy.x; c[cs] = axy.y; c[2 * cs] = abz.x; c[3 * cs] = bxy.x; c[4 * cs] = bxy.y; c[5 * cs] = abz.y;
            if (!BOX_IN_LOOP && l >= 2) box_tests(M, l, f3(axy.x, axy.y, abz.x), f3(bxy.x, bxy.y, abz.y), hit, slow);
        }
    }
    float3 e = euler_from_posep(T);
    ee[0] = T.pxy.x; ee[1] = T.pxy.y; ee[2] = T.pz; ee[3] = e.x; ee[4] = e.y; ee[5] = e.z;
    if (!collide) return false;
    // links 2..6 vs obstacle (distances kept: they are get_link_distances' values)
    if (BOX_IN_LOOP || Traits<TASK>::HAS_OBST) {
#pragma unroll 2     // two links per trip: their dependent chains interleave (rolled: +10 % kernel time; fully unrolled it spills)
        for (int l = 2; l < 7; l++) {
            const float *c = cap + (l - 1) * 6 * cs;
            const float3 a = f3(c[0], c[cs], c[2 * cs]), b = f3(c[3 * cs], c[4 * cs], c[5 * cs]);
            if (Traits<TASK>::HAS_OBST) {       // keys[5] == 'obstacle'   pyb_setup.py:398-399
                float d = sqrtf(seg_axis_dist2(a, b, O.c, O.u, M.fit_obst_h, M.cap_ia[l])) - M.fit_obst[l];
                hit = hit || (d <= URGYM_COLLISION_MARGIN);
                cap[(36 + l - 2) * cs] = d;
            }
            if (BOX_IN_LOOP) box_tests(M, l, a, b, hit, slow);
        }
    }
    while (slow) {          // exact segment-box distance for the few (link, box) cases left
        const int k = __ffs_hd(slow) - 1;
        slow &= slow - 1u;
        const int l = 2 + (k >> 1), box = k & 1;
        const float *c = cap + (l - 1) * 6 * cs;
        const float3 a = f3(c[0], c[cs], c[2 * cs]), b = f3(c[3 * cs], c[4 * cs], c[5 * cs]);
        const float3 bc = f3(M.box_c[box][0], M.box_c[box][1], M.box_c[box][2]);
        const float3 bh = f3(M.box_he[box][0], M.box_he[box][1], M.box_he[box][2]);
#ifdef URGYM_SLOW_PROBE
        // most of what is left are hits, and for most hits an end point or the midpoint is already within reach
        const float reach = URGYM_COLLISION_MARGIN + M.fit_box[l] + M.box_margin[box];
        const float pr = fminf(fminf(point_box_dist2(a, bc, bh), point_box_dist2(b, bc, bh)), point_box_dist2(0.5f * (a + b), bc, bh));
        if (pr <= reach * reach) { hit = true; continue; }
#endif
        constexpr bool ROLLED = TASK == TASK_STA || TASK == TASK_DYN;
        hit = hit || (sqrtf(seg_box_dist2_t<ROLLED>(a, b, bc, bh)) - M.fit_box[l] - M.box_margin[box] <= URGYM_COLLISION_MARGIN);
    }
    // self pairs (1:3,4,5,6) (2:4,5,6) (3:5,6), pair index p = 0..8 in that order: a broad phase for all nine with static
    // indices, then the exact segment-segment test only for the pairs it leaves, in a rolled loop.
    {
        // Broad phase.  Every pair has a short member (links 1, 4, 5, 6: capsule half lengths 2..34 mm), contained
        // in the ball of that radius about its midpoint, so  dist(pair) >= dist(midpoint, other segment) - half length.
        // For the pairs with a long member (upper arm 2, forearm 3) that point-segment bound is far tighter than the
        // sphere-sphere bound: in steady state the spheres let 100 % of the (3,5) and 28 % of the (3,6) pairs through
        // (1.4 exact tests per env, 3.6 for the slowest lane of a warp), the point-segment bound ~1 % in total.
        // (1,4) (1,5) (1,6) keep the sphere test (both members short).
        float3 mid[6], sa[2], sd[2];        // midpoints of links 1, 4, 5, 6 (index l - 1); a and b - a of links 2, 3
#pragma unroll
        for (int l = 0; l < 6; l++) {
            const float *c = cap + l * 6 * cs;
            const float3 a = f3(c[0], c[cs], c[2 * cs]), b = f3(c[3 * cs], c[4 * cs], c[5 * cs]);
            if (l == 1 || l == 2) { sa[l - 1] = a; sd[l - 1] = b - a; }
            else mid[l] = 0.5f * (a + b);
        }
        unsigned need = 0u;
#pragma unroll
        for (int p = 0; p < 9; p++) {
            const int l1 = p < 4 ? 1 : (p < 7 ? 2 : 3);
            const int l2 = p < 4 ? p + 3 : (p < 7 ? p : p - 2);
            float d2;
            if (l1 == 1 && l2 != 3) {                       // short - short: sphere test
                const float3 dm = mid[0] - mid[l2 - 1];
                d2 = dot(dm, dm);
            } else {                                        // midpoint of the short member vs the long member's segment
                const int ll = (l1 == 1) ? 3 : l1, ls = (l1 == 1) ? 1 : l2;
                const float3 ap = mid[ls - 1] - sa[ll - 2];
                const float t = clampf(dot(ap, sd[ll - 2]) * M.cap_ia[ll], 0.0f, 1.0f);
                const float3 e = ap - t * sd[ll - 2];
                d2 = dot(e, e);
            }
            if (d2 <= M.self_far2[p]) need |= 1u << p;
        }
        while (need) {
            const int p = __ffs_hd(need) - 1;
            need &= need - 1u;
            const int l1 = p < 4 ? 1 : (p < 7 ? 2 : 3);
            const int l2 = p < 4 ? p + 3 : (p < 7 ? p : p - 2);
            const float *c1 = cap + (l1 - 1) * 6 * cs, *c2 = cap + (l2 - 1) * 6 * cs;
            const float3 a1 = f3(c1[0], c1[cs], c1[2 * cs]), b1 = f3(c1[3 * cs], c1[4 * cs], c1[5 * cs]);
            const float3 a2 = f3(c2[0], c2[cs], c2[2 * cs]), b2 = f3(c2[3 * cs], c2[4 * cs], c2[5 * cs]);
            hit = hit || (segseg_dist2_fast(a1, b1, a2, b2, M.cap_ia[l1], M.cap_ia[l2]) <= M.self_reach2[p]);
        }
    }
    if (Traits<TASK>::HAS_OBST) {
        if (WORKBENCH) workbench_link_dist(M, cap, cs);
#pragma unroll
        for (int k = 0; k < 5; k++) dist[k] = cap[(36 + k) * cs];
    }
    return hit;
}

// geometry dispatch.  qrow: joint angles in memory (rolled hull pass); q: the same in registers (capsule pass)
template <int TASK, int GEOM>
URGYM_HD bool robot_pass(const ModelConst &M, const float *q, const float *qrow, const ObstW &O, const float4 *hv,
                         bool collide, float *ee, float *dist, float *scratch, int cs) {
    if (URGYM_BASE(GEOM) == GEOM_CAPSULE) return robot_pass_capsule<TASK, URGYM_WB(GEOM)>(M, q, O, collide, ee, dist, scratch, cs);
    return robot_pass_rolled<TASK, GEOM>(M, qrow, O, hv, collide, ee, dist[0], dist[1], dist[2], dist[3], dist[4]);
}

// task part of the observation   reach.py:189-190 (Ori), 307-308 (Obs), 454-458 (Sta), 653-657 (Dyn)
// row[0..11] = robot obs (ee pos, ee euler, q).  vel == nullptr keeps the velocity columns as they are (stale, quirk Q4).
template <int TASK>
URGYM_HD void write_obs_row(float *row, const float *ee, const float *q, const float *E, const ObstW &O, float3 e,
                            const float *vel, const float *ld) {
    typedef Traits<TASK> TT;
#pragma unroll
    for (int k = 0; k < 6; k++) { row[k] = ee[k]; row[6 + k] = q[k]; }
#pragma unroll
    for (int k = 0; k < TT::GOAL; k++) row[12 + k] = E[k];
    if (TASK == TASK_OBS) {             // obstacle as sampled (quirk Q3)
#pragma unroll
        for (int k = 0; k < 6; k++) row[15 + k] = E[3 + k];
#pragma unroll
        for (int k = 0; k < 5; k++) row[21 + k] = ld[k];
    } else if (TASK == TASK_STA || TASK == TASK_DYN) {   // obstacle pose read back: position + getEulerFromQuaternion (e)
        row[18] = O.c.x; row[19] = O.c.y; row[20] = O.c.z; row[21] = e.x; row[22] = e.y; row[23] = e.z;
        if (TASK == TASK_DYN) {
            if (vel) {
#pragma unroll
                for (int k = 0; k < 6; k++) row[24 + k] = vel[k];
            }
#pragma unroll
            for (int k = 0; k < 5; k++) row[30 + k] = ld[k];
        } else {
#pragma unroll
            for (int k = 0; k < 5; k++) row[24 + k] = ld[k];
        }
    }
}

// success test + the two goal distances   reach.py:212-215,348-350,543-546,755-758; utils.py:5-69
template <int TASK>
URGYM_HD bool goal_metrics(const float *ee, const float *E, const float *C, float &d, float &ang) {
    float dx = ee[0] - E[0], dy = ee[1] - E[1], dz = ee[2] - E[2];
    d = sqrtf(dx * dx + dy * dy + dz * dz);
    bool ok = d < 0.05f;
    ang = 0.0f;
    if (Traits<TASK>::ORI) {
        Quat g; g.x = C[0]; g.y = C[1]; g.z = C[2]; g.w = C[3];
        ang = angular_distance(quat_ZYX<true>(ee[3], ee[4], ee[5]), g);        // ee[3..5] come out of atan2
        ok = ok && (ang < 0.0873f);
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------ step
// RobotTaskEnv.step (core.py:303-317) + TimeLimit, without the auto-reset, in three parts: env_step_begin (action,
// obstacle motion), the robot pass (FK, collision, link distances), env_step_finish (observation, termination, reward).
// The hull-geometry step kernel runs its own robot pass between the two (urgym_kernels.cuh: the exact GJK tests of a
// whole block are compacted into dense task lists); everything else calls env_step.
// `row` receives the observation (OBS floats).  vel_out (Dyn, 6 floats): ReachDyn.velocity after this step.
template <int TASK, int GEOM>
URGYM_HD ObstW env_step_begin(EnvState &s, const float *act, float *row, float *vel_out, float *vel, float3 &oe) {
    typedef Traits<TASK> TT;
    // 1. UR5Ori.set_action: clip, * pi, * 0.1 (float32 like the numpy expression), teleport      UR5.py:273-279,314-317
#pragma unroll
    for (int j = 0; j < 6; j++) s.q[j] += (clampf(act[j], -1.0f, 1.0f) * URGYM_PI_F) * 0.1f;
    // 2. task.set_velocity + sim.step: obstacle pose after this step (twist from the episode cache)   core.py:305-309
#pragma unroll
    for (int k = 0; k < 6; k++) vel[k] = 0.0f;
    const ObstW O = obstacle_cached<TASK>(s.E, s.C, s.elapsed + 1, oe);
    if (TT::DYN) {
        if (s.elapsed < 25) {           // ReachDyn.velocity: the twist while step_num < 25, zeros afterwards
#pragma unroll
            for (int k = 0; k < 6; k++) vel[k] = s.C[8 + k];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) vel_out[k] = vel[k];
    }
    if (URGYM_BASE(GEOM) != GEOM_CAPSULE) {     // the rolled hull pass reads the joint angles from memory
#pragma unroll
        for (int k = 0; k < 6; k++) row[6 + k] = s.q[k];
    }
    return O;
}
// `done_hook(finished)` is called (by every lane, converged) as soon as it is known whether the episode ends with this step:
// the step kernel reserves its auto-reset queue entries there, ~150 instructions before it needs the reply.
struct NoDoneHook { URGYM_HD void operator()(bool) const {} };
template <int TASK, class DoneHook = NoDoneHook>
URGYM_HD void env_step_finish(EnvState &s, float *row, const float *ee, const float *dist, bool coll, const ObstW &O, float3 oe,
                              const float *vel, StepOut &o, DoneHook done_hook = DoneHook()) {
    // 5. termination (computed ahead of the observation row, which does not depend on it)      core.py:313-315
    float d, ang;
    bool succ = goal_metrics<TASK>(ee, s.E, s.C, d, ang);
    o.collision = coll;
    o.terminated = succ || coll;
    o.success = o.terminated && !coll;
    done_hook(o.terminated || s.elapsed + 1 >= URGYM_MAX_STEPS);
    // 4. observation: carries link_dist from BEFORE this step's reward (quirk Q1)               core.py:311
    write_obs_row<TASK>(row, ee, s.q, s.E, O, oe, vel, s.ld);
    // 6. reward                                                                                core.py:316
    float r;
    if (TASK == TASK_ORI) {             // reach.py:221-236
        r = (succ ? 200.0f : 0.0f) - 70.0f * d - 30.0f * ang - (coll ? 500.0f : 0.0f);
    } else if (TASK == TASK_OBS) {      // reach.py:356-374
        r = (succ ? 200.0f : 0.0f) - (coll ? 500.0f : 0.0f) - 100.0f * d;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            if (dist[k] < 0.2f) acc += 100.0f * (dist[k] - s.ld[k]);
            s.ld[k] = dist[k];
        }
        r += acc;
    } else {                            // reach.py:552-573, 764-785
        if (coll) r = -500.0f;
        else if (succ) r = 200.0f;
        else {
            const float w[5] = {8.0f / 13.0f * 50.0f, 2.4f / 13.0f * 50.0f, 1.2f / 13.0f * 50.0f, 1.2f / 13.0f * 50.0f,
                                0.2f / 13.0f * 50.0f};
            r = -70.0f * d - 30.0f * ang;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                if (dist[k] < 0.2f) acc += w[k] * (dist[k] - s.ld[k]);
                s.ld[k] = dist[k];
            }
            r += acc;
        }
    }
    o.reward = r;
    // 7. TimeLimit
    s.elapsed += 1;
    o.truncated = s.elapsed >= URGYM_MAX_STEPS;
    s.ep_ret += r;
}
template <int TASK, int GEOM, class DoneHook = NoDoneHook>
URGYM_HD void env_step(const ModelConst &M, EnvState &s, const float *act, const float4 *hv, float *row, StepOut &o,
                       float *vel_out, float *scratch, int cs, DoneHook done_hook = DoneHook()) {
    float3 oe;
    float vel[6];
    const ObstW O = env_step_begin<TASK, GEOM>(s, act, row, vel_out, vel, oe);
    // 3. FK and collision                                                                      core.py:310
    float ee[6], dist[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    bool coll = robot_pass<TASK, GEOM>(M, s.q, row + 6, O, hv, true, ee, dist, scratch, cs);
    URGYM_WARP_SYNC();      // every lane is done with the scratch before observation rows are written over it
    env_step_finish<TASK>(s, row, ee, dist, coll, O, oe, vel, o, done_hook);
}

// RobotTaskEnv._get_obs (core.py:252-261) from the current state, no stepping.  stale_vel: ReachDyn.velocity as
// stored at the last reset, shown while elapsed == 0 (quirk Q4).
template <int TASK, int GEOM>
URGYM_HD void env_observe(const ModelConst &M, const EnvState &s, const float *stale_vel, float *row) {
    typedef Traits<TASK> TT;
    float3 oe;
    float vel[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const ObstW O = obstacle_cached<TASK>(s.E, s.C, s.elapsed, oe);
    if (TT::DYN) {
#pragma unroll
        for (int k = 0; k < 6; k++) vel[k] = s.elapsed == 0 ? stale_vel[k] : (s.elapsed <= 25 ? s.C[8 + k] : 0.0f);
    }
    float ee[6], du[5];
#pragma unroll
    for (int k = 0; k < 6; k++) row[6 + k] = s.q[k];
    robot_pass<TASK, GEOM_CAPSULE>(M, s.q, row + 6, O, nullptr, false, ee, du, nullptr, 1);   // EE pose only
    write_obs_row<TASK>(row, ee, s.q, s.E, O, oe, vel, s.ld);
}

// tail of set_goal_and_obstacle / reset: collision flag and link_dist = last_dist at the current state
// reach.py:322-324,333-335,477-479,501-503,679-681,711-713
template <int TASK, int GEOM>
URGYM_HD bool env_refresh(const ModelConst &M, EnvState &s, const float4 *hv, float *scratch, int cs) {
    typedef Traits<TASK> TT;
    float3 oe;
    const ObstW O = obstacle_cached<TASK>(s.E, s.C, s.elapsed, oe);
    float ee[6], dist[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    bool coll = robot_pass<TASK, GEOM>(M, s.q, s.q, O, hv, true, ee, dist, scratch, cs);
    if (TT::HAS_OBST) {
#pragma unroll
        for (int k = 0; k < 5; k++) s.ld[k] = dist[k];
    }
    return coll;
}

// ------------------------------------------------------------------------------------------------ reset
// uniform draw `slot` of rejection iteration rs.iter: Philox block slot/4, lane slot%4
struct Draws {
    uint4 b[5];
    URGYM_HD float u(int slot) const {
        uint4 x = b[slot >> 2];
        int k = slot & 3;
        return u01(k == 0 ? x.x : (k == 1 ? x.y : (k == 2 ? x.z : x.w)));
    }
};
URGYM_HD float lerp_u(float lo, float hi, float u) { return lo + (hi - lo) * u; }     // np_random.uniform(low, high)

// ReachDyn's cheap half of the rejection rule for iteration rs.iter: d(obstacle_end, obstacle_start) >= 1 m
// (reach.py:674-675).  Only ~17.5 % of the draws pass, so the auto-reset kernel searches iterations in parallel
// with this test before the full sample is evaluated.
URGYM_HD bool dyn_pair_far_enough(const ResetStream &rs) {
    float3 olo, ohi;
    obstacle_range<TASK_DYN>(olo, ohi);
    Draws D;
    rs.blocks<2>(0, &D.b[0]);
    float dx = lerp_u(olo.x, ohi.x, D.u(3)) - lerp_u(olo.x, ohi.x, D.u(0));
    float dy = lerp_u(olo.y, ohi.y, D.u(4)) - lerp_u(olo.y, ohi.y, D.u(1));
    float dz = lerp_u(olo.z, ohi.z, D.u(5)) - lerp_u(olo.z, ohi.z, D.u(2));
    return !(sqrtf(dx * dx + dy * dy + dz * dz) < 1.0f);
}

// Reach*.reset (reach.py:197-200,313-326,465-481,664-683) with _sample_goal / _sample_obstacle
// (reach.py:206-210,337-346,505-516,715-726): fills E, returns the number of rejection iterations used.
// k_start: the first rejection iteration to look at (the caller may already know that earlier ones fail).
template <int TASK, int GEOM>
URGYM_HD int sample_episode(const ModelConst &M, ResetStream rs, float *E, int k_start = 0) {
    typedef Traits<TASK> TT;
    float3 glo, ghi, olo, ohi;
    goal_range<TASK>(glo, ghi);
    obstacle_range<TASK>(olo, ohi);
    int k = k_start;
    for (;;) {
        rs.iter = (uint32_t)k;
        Draws D;
        bool fail = false;
        if (TASK == TASK_ORI) {                     // slots: goal 0-2, goal_roll 3, goal_yaw 4
            rs.blocks<2>(0, &D.b[0]);
            E[0] = lerp_u(glo.x, ghi.x, D.u(0)); E[1] = lerp_u(glo.y, ghi.y, D.u(1)); E[2] = lerp_u(glo.z, ghi.z, D.u(2));
            E[4] = 0.0f;
            euler_constrained(D.u(3), D.u(4), E[3], E[5]);
        } else if (TASK == TASK_OBS) {              // goal 0-2, obstacle 3-5, sign 6, roll 7, pitch 8
            rs.blocks<3>(0, &D.b[0]);
            E[0] = lerp_u(glo.x, ghi.x, D.u(0)); E[1] = lerp_u(glo.y, ghi.y, D.u(1)); E[2] = lerp_u(glo.z, ghi.z, D.u(2));
            E[3] = lerp_u(olo.x, ohi.x, D.u(3)); E[4] = lerp_u(olo.y, ohi.y, D.u(4)); E[5] = lerp_u(olo.z, ohi.z, D.u(5));
            euler_obstacle(D.u(6), D.u(7), D.u(8), E[6], E[7]);
            E[8] = 0.0f;
            fail = target_obstacle_dist<TASK, GEOM>(M, E, obstacle_static(&E[3])) < 0.1f;        // reach.py:321
        } else if (TASK == TASK_STA) {              // goal 0-2, roll 3, yaw 4, obstacle 5-7, sign 8, roll 9, pitch 10
            rs.blocks<3>(0, &D.b[0]);
            E[0] = lerp_u(glo.x, ghi.x, D.u(0)); E[1] = lerp_u(glo.y, ghi.y, D.u(1)); E[2] = lerp_u(glo.z, ghi.z, D.u(2));
            E[4] = 0.0f;
            euler_constrained(D.u(3), D.u(4), E[3], E[5]);
            E[6] = lerp_u(olo.x, ohi.x, D.u(5)); E[7] = lerp_u(olo.y, ohi.y, D.u(6)); E[8] = lerp_u(olo.z, ohi.z, D.u(7));
            euler_obstacle(D.u(8), D.u(9), D.u(10), E[9], E[10]);
            E[11] = 0.0f;
            fail = target_obstacle_dist<TASK, GEOM>(M, E, obstacle_static(&E[6])) < 0.1f;        // reach.py:473
        } else {                                    // Dyn: start 0-2, end 3-5, goal 6-8, roll 9, yaw 10, start s/r/p 11-13, end 14-16
            rs.blocks<2>(0, &D.b[0]);
            E[6] = lerp_u(olo.x, ohi.x, D.u(0)); E[7] = lerp_u(olo.y, ohi.y, D.u(1)); E[8] = lerp_u(olo.z, ohi.z, D.u(2));
            E[12] = lerp_u(olo.x, ohi.x, D.u(3)); E[13] = lerp_u(olo.y, ohi.y, D.u(4)); E[14] = lerp_u(olo.z, ohi.z, D.u(5));
            float dx = E[12] - E[6], dy = E[13] - E[7], dz = E[14] - E[8];
            fail = sqrtf(dx * dx + dy * dy + dz * dz) < 1.0f;                                     // reach.py:674-675
            if (!fail || k + 1 >= URGYM_MAX_RESET_ITERS) {      // the other draws matter only for a surviving iteration
                rs.blocks<3>(2, &D.b[2]);
                E[0] = lerp_u(glo.x, ghi.x, D.u(6)); E[1] = lerp_u(glo.y, ghi.y, D.u(7)); E[2] = lerp_u(glo.z, ghi.z, D.u(8));
                E[4] = 0.0f;
                euler_constrained(D.u(9), D.u(10), E[3], E[5]);
                euler_obstacle(D.u(11), D.u(12), D.u(13), E[9], E[10]);
                E[11] = 0.0f;
                euler_obstacle(D.u(14), D.u(15), D.u(16), E[15], E[16]);
                E[17] = 0.0f;
                fail = fail || (target_obstacle_dist<TASK, GEOM>(M, E, obstacle_static(&E[12])) < 0.1f);
            }
        }
        k++;
        if (!fail || k >= URGYM_MAX_RESET_ITERS) break;
    }
    return k;
}

// RobotTaskEnv.reset (core.py:263-273): neutral pose, new episode constants, link_dist = last_dist at the reset
// pose, first observation.  The velocity columns of `row` are left untouched (quirk Q4: the caller puts the
// previous ReachDyn.velocity there).  ReachSta: s.E[12..24) (obstacle_end / obstacle_start) must hold the env's current
// values on entry, the reset keeps them (reach.py:465-481).  Returns the rejection iterations used.
template <int TASK, int GEOM>
URGYM_HD int env_reset(const ModelConst &M, EnvState &s, ResetStream rs, const float4 *hv, float *row, int k_start = 0) {
    typedef Traits<TASK> TT;
    int iters = sample_episode<TASK, GEOM>(M, rs, s.E, k_start);
    derive_cache<TASK>(s.E, s.C);
#pragma unroll
    for (int j = 0; j < 6; j++) s.q[j] = M.neutral_q[j];
    s.elapsed = 0;
    s.ep_ret = 0.0f;
    float3 oe;
    const ObstW O = obstacle_cached<TASK>(s.E, s.C, 0, oe);     // Dyn: obstacle placed at START after sampling   reach.py:678
    if (TT::HAS_OBST) {
#pragma unroll 1
        for (int l = 2; l < 7; l++) {                   // reach.py:323-324,478-479,680-681
            LinkShape<URGYM_BASE(GEOM)> L;
            L.set_neutral(M, l, hv);
            float d = L.obstacle_dist(M, l, O);
            if (URGYM_WB(GEOM)) d = fminf(d, fminf(L.box_dist(M, l, 0), L.box_dist(M, l, 1)));
            if (l == 2) s.ld[0] = d; else if (l == 3) s.ld[1] = d; else if (l == 4) s.ld[2] = d; else if (l == 5) s.ld[3] = d; else s.ld[4] = d;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 5; k++) s.ld[k] = 0.0f;
    }
    write_obs_row<TASK>(row, M.neutral_ee, s.q, s.E, O, oe, nullptr, s.ld);
    return iters;
}

}  // namespace urgym
