/* oracle/ur_oracle_sim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (double precision) restatement of the *physics-engine tier* ("T2") of UR-gym's reach-task step path:
 * the subset of PyBullet / Bullet behaviour that UR_gym/pyb_setup.py calls on that path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY STATUS: **parity unpinned** for this tier.  PyBullet (third-party, declared *unpinned* at
 * /root/reference/setup.py:22, source not vendored) cannot be installed or run in this environment, and the
 * reference ships no tests or golden vectors.  Every function below restates Bullet's
 * published algorithm from memory and names the reference call site it stands in for; every assumption that a
 * real PyBullet run could overturn is collected in `orc_flags_t` so it can be flipped in one place.
 *
 * What stands in for what (reference call site -> function here):
 *   getLinkState(body, link)[0:2]        pyb_setup.py:231,244   -> orc_fk()
 *   getQuaternionFromEuler               pyb_setup.py:152,314   -> orc_quat_from_euler()
 *   getEulerFromQuaternion               pyb_setup.py:190,248   -> orc_euler_from_quat()
 *   getDifferenceQuaternion              pyb_setup.py:359       -> orc_quat_difference()
 *   getAxisAngleFromQuaternion           pyb_setup.py:363       -> orc_axis_angle()
 *   resetBaseVelocity + 20x stepSimulation on a mass-0 body  pyb_setup.py:348,52-55 -> orc_integrate_base()
 *   getClosestPoints(...)[0][8]          pyb_setup.py:401,410,421,436,452 -> orc_pair_distance() and the
 *                                        scene-level helpers orc_check_collision / orc_link_distances /
 *                                        orc_target_obstacle_distance
 *
 * Geometry model of getClosestPoints (Bullet btGjkPairDetector semantics):
 *   signed distance = || closest points of the two CORE shapes || - marginA - marginB
 *   - URDF mesh links load as convex hulls of the mesh vertices with margin 0.001 (core = the hull itself)
 *   - createCollisionShape primitives (box, cylinder) get margin 0.001 and a core shrunk by that margin
 *     (flag prim_margin_mode=0), or Bullet's constructor "safe margin" (flag =1; SURVEY App. B-8's reading)
 *   - a sphere is a point core with margin = radius
 *   GJK here runs to a much tighter tolerance than Bullet's (REL_ERROR2 = 1e-6 on squared distance), so for
 *   separated cores it returns the true distance, which Bullet's own answer exceeds by at most 1e-6 * d.
 *   When the two cores interpenetrate Bullet switches to EPA (accuracy 1e-4 m); here that case is reported
 *   with *deep = 1 and distance = -(marginA+marginB) (penetration depth of the cores taken as 0).
 *
 * Geometry modes:  ORC_GEOM_HULL    links are the reference's convex hulls (margin 0.001)
 *                  ORC_GEOM_CAPSULE the product's throughput geometry, not the reference's: links are the
 *                                   extractor's bounding capsules (segment core, margin = r + hull margin, so a
 *                                   capsule distance never exceeds the hull distance), the obstacle cylinder is its
 *                                   bounding capsule (segment half length 0.2, margin 0.05), the target is a sphere
 *                                   (Obs: r 0.02 as in the reference; Sta/Dyn: the cube's bounding sphere);
 *                                   table and track stay the reference's boxes.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include "../ur-gym_b200/csrc/ur5e_model_data.h"
#include "../ur-gym_b200/csrc/urgym_capsule_fit.h"

#define ORC_GEOM_HULL 0
#define ORC_GEOM_CAPSULE 1

typedef struct {
    int prim_margin_mode;   /* 0: primitives margin 0.001 (processCreateCollisionShapeCommand default margin)
                               1: Bullet constructor safe margin (10% of the smallest half extent, <= 0.04) */
    double hull_margin;     /* URDF default collision margin for mesh links: 0.001 */
    double gjk_rel_tol;     /* termination: vv - v.w <= tol * vv */
    int gjk_max_iter;
} orc_flags_t;

static orc_flags_t g_flags = {0, 0.001, 1e-13, 2000};

/* capsule geometry: what is subtracted from the distance between the segment cores, per pair class (defaults from the
 * generated urgym_capsule_fit.h; oracle/calibrate_capsules.py overrides them while fitting) */
static double g_fit_obst[7], g_fit_box[7], g_fit_self[9], g_fit_obst_h;
static int g_fit_init = 0;
static void fit_init(void) {
    if (g_fit_init) return;
    for (int i = 0; i < 7; i++) { g_fit_obst[i] = URGYM_FIT_OBST[i]; g_fit_box[i] = URGYM_FIT_BOX[i]; }
    for (int i = 0; i < 9; i++) g_fit_self[i] = URGYM_FIT_SELF[i];
    g_fit_obst_h = URGYM_FIT_OBST_H;
    g_fit_init = 1;
}
void orc_set_capsule_fit(const double obst[7], const double box[7], const double self[9], double obst_h) {
    for (int i = 0; i < 7; i++) { g_fit_obst[i] = obst[i]; g_fit_box[i] = box[i]; }
    for (int i = 0; i < 9; i++) g_fit_self[i] = self[i];
    g_fit_obst_h = obst_h;
    g_fit_init = 1;
}

void orc_set_flags(int prim_margin_mode, double hull_margin, double gjk_rel_tol, int gjk_max_iter) {
    g_flags.prim_margin_mode = prim_margin_mode;
    g_flags.hull_margin = hull_margin;
    g_flags.gjk_rel_tol = gjk_rel_tol;
    g_flags.gjk_max_iter = gjk_max_iter;
}

/* ------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void matvec(const double *R, const double *v, double *o) {   /* o = R v, R row-major */
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
static inline void matTvec(const double *R, const double *v, double *o) {  /* o = R^T v */
    o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
static inline void matmul(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* ------------------------------------------------------------------ quaternion / Euler (PyBullet semantics) */
/* pybullet getQuaternionFromEuler: q = Rz(yaw) Ry(pitch) Rx(roll), (x,y,z,w)          pyb_setup.py:152 */
void orc_quat_from_euler(const double e[3], double q[4]) {
    double phi = e[0] / 2.0, the = e[1] / 2.0, psi = e[2] / 2.0;
    q[0] = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
    q[1] = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
    q[2] = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
    q[3] = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= n;
}

/* pybullet getEulerFromQuaternion -> (roll, pitch, yaw) with the gimbal branch at |sarg| >= 0.99999
 * pyb_setup.py:190,248 */
void orc_euler_from_quat(const double q[4], double e[3]) {
    double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    double sarg = -2.0 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= -0.99999) {
        e[1] = -0.5 * M_PI; e[0] = 0.0; e[2] = 2.0 * atan2(q[0], -q[1]);
    } else if (sarg >= 0.99999) {
        e[1] = 0.5 * M_PI; e[0] = 0.0; e[2] = 2.0 * atan2(-q[0], q[1]);
    } else {
        e[1] = asin(sarg);
        e[0] = atan2(2.0 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        e[2] = atan2(2.0 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}

static void quat_mul(const double a[4], const double b[4], double o[4]) {   /* Hamilton, (x,y,z,w) */
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* pybullet getDifferenceQuaternion(start, end) = nearest(end) * start^-1            pyb_setup.py:359 */
void orc_quat_difference(const double s[4], const double e[4], double o[4]) {
    double dm = 0, dp = 0, e1[4], sinv[4];
    for (int i = 0; i < 4; i++) { dm += (s[i] - e[i]) * (s[i] - e[i]); dp += (s[i] + e[i]) * (s[i] + e[i]); }
    for (int i = 0; i < 4; i++) e1[i] = (dm < dp) ? e[i] : -e[i];
    sinv[0] = -s[0]; sinv[1] = -s[1]; sinv[2] = -s[2]; sinv[3] = s[3];
    quat_mul(e1, sinv, o);
}

/* pybullet getAxisAngleFromQuaternion: angle = 2 acos(w); axis = xyz / sqrt(1-w^2), (1,0,0) if degenerate
 * pyb_setup.py:363 */
void orc_axis_angle(const double q[4], double axis[3], double *angle) {
    double w = q[3] > 1.0 ? 1.0 : (q[3] < -1.0 ? -1.0 : q[3]);
    *angle = 2.0 * acos(w);
    double s2 = 1.0 - q[3] * q[3];
    if (s2 < 10.0 * 2.2204460492503131e-16) { axis[0] = 1; axis[1] = 0; axis[2] = 0; return; }
    double s = 1.0 / sqrt(s2);
    axis[0] = q[0] * s; axis[1] = q[1] * s; axis[2] = q[2] * s;
}

static void quat_to_mat(const double q[4], double R[9]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = x * x + y * y + z * z + w * w, s = 2.0 / n;
    R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z);     R[2] = s * (x * z + w * y);
    R[3] = s * (x * y + w * z);     R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
    R[6] = s * (x * z - w * y);     R[7] = s * (y * z + w * x);     R[8] = 1 - s * (x * x + y * y);
}

static void mat_to_quat(const double R[9], double q[4]) {   /* btMatrix3x3::getRotation */
    double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) {
        double s = sqrt(tr + 1.0);
        q[3] = s * 0.5; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = s * 0.5; s = 0.5 / s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
}

/* Kinematic motion of a mass-0 createMultiBody base under resetBaseVelocity over n substeps of dt
 * (btMultiBody::stepPositionsMultiDof: position explicit Euler, orientation by the exponential map of the
 * world-frame angular velocity, renormalised every substep).  [RECALLED: SURVEY App. B-5]
 * pyb_setup.py:340-349 (set_velocity) + pyb_setup.py:52-55 (step) */
void orc_integrate_base(double pos[3], double quat[4], const double v[3], const double w[3], double dt, int nsub) {
    for (int s = 0; s < nsub; s++) {
        for (int i = 0; i < 3; i++) pos[i] += dt * v[i];
        double fAngle = sqrt(dot3(w, w)), ax[3];
        if (fAngle * dt > 0.5 * (M_PI / 2.0)) fAngle = 0.5 * (M_PI / 2.0) / dt;   /* ANGULAR_MOTION_THRESHOLD */
        double k;
        if (fAngle < 0.001) k = 0.5 * dt - dt * dt * dt * 0.020833333333 * fAngle * fAngle;
        else k = sin(0.5 * fAngle * dt) / fAngle;
        for (int i = 0; i < 3; i++) ax[i] = w[i] * k;
        double dq[4] = {ax[0], ax[1], ax[2], cos(fAngle * dt * 0.5)}, o[4];
        quat_mul(dq, quat, o);   /* world-frame increment applied on the left */
        double n = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
        for (int i = 0; i < 4; i++) quat[i] = o[i] / n;
    }
}

/* ------------------------------------------------------------------ forward kinematics
 * World pose of PyBullet link frames 0..6 for joint angles q (joints 1..6), base at the origin.
 * T_i = T_{i-1} * Trans(xyz_i) * Rz(y)Ry(p)Rx(r) * Rz(q_i)     (ur5e.urdf:232-279; UR5.py:258)
 * Link 7 (ee_link) = link 6 frame (identity fixed joint, no <inertial>): UR5.py:263,334-340.          */
void orc_fk(const double q[6], double pos[7][3], double rot[7][9]) {
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(rot[0], I, sizeof(I));
    pos[0][0] = pos[0][1] = pos[0][2] = 0.0;
    for (int i = 0; i < 6; i++) {
        double t[3], Rf[9], Rq[9], Rl[9];
        matvec(rot[i], &UR5E_JOINT_XYZ[3 * i], t);
        for (int k = 0; k < 3; k++) pos[i + 1][k] = pos[i][k] + t[k];
        matmul(rot[i], &UR5E_JOINT_ROT[9 * i], Rf);
        double c = cos(q[i]), s = sin(q[i]);
        Rq[0] = c; Rq[1] = -s; Rq[2] = 0; Rq[3] = s; Rq[4] = c; Rq[5] = 0; Rq[6] = 0; Rq[7] = 0; Rq[8] = 1;
        matmul(Rf, Rq, Rl);
        memcpy(rot[i + 1], Rl, sizeof(Rl));
    }
}

/* EE pose as the reference reads it: position + PyBullet Euler triple of link 7.   UR5.py:334-340 */
void orc_ee_pose(const double q[6], double p[3], double e[3]) {
    double pos[7][3], rot[7][9], qt[4];
    orc_fk(q, pos, rot);
    memcpy(p, pos[6], 3 * sizeof(double));
    mat_to_quat(rot[6], qt);
    orc_euler_from_quat(qt, e);
}

/* ------------------------------------------------------------------ convex shapes + GJK */
enum { SH_HULL = 0, SH_BOX = 1, SH_CYLZ = 2, SH_POINT = 3 };
typedef struct {
    int type;
    const double *verts; int nverts;   /* SH_HULL: local-frame vertices */
    double he[3];                      /* SH_BOX: core half extents; SH_CYLZ: (r, r, half height) of the core */
    double R[9], t[3];                 /* local -> world */
    double margin;
} shape_t;

static void support(const shape_t *s, const double dw[3], double out[3]) {
    double d[3], l[3];
    matTvec(s->R, dw, d);
    switch (s->type) {
    case SH_HULL: {
        int best = 0; double bd = -1e300;
        for (int i = 0; i < s->nverts; i++) {
            double v = dot3(&s->verts[3 * i], d);
            if (v > bd) { bd = v; best = i; }
        }
        l[0] = s->verts[3 * best]; l[1] = s->verts[3 * best + 1]; l[2] = s->verts[3 * best + 2];
        break;
    }
    case SH_BOX:
        l[0] = d[0] >= 0 ? s->he[0] : -s->he[0];
        l[1] = d[1] >= 0 ? s->he[1] : -s->he[1];
        l[2] = d[2] >= 0 ? s->he[2] : -s->he[2];
        break;
    case SH_CYLZ: {
        double sl = sqrt(d[0] * d[0] + d[1] * d[1]);
        if (sl != 0.0) { l[0] = d[0] * s->he[0] / sl; l[1] = d[1] * s->he[0] / sl; }
        else { l[0] = s->he[0]; l[1] = 0.0; }
        l[2] = d[2] < 0.0 ? -s->he[2] : s->he[2];
        break;
    }
    default: l[0] = l[1] = l[2] = 0.0;
    }
    matvec(s->R, l, out);
    out[0] += s->t[0]; out[1] += s->t[1]; out[2] += s->t[2];
}

/* closest point to the origin on a triangle (Ericson, Real-Time Collision Detection 5.1.5);
 * returns barycentric weights */
static void closest_tri(const double *a, const double *b, const double *c, double lam[3]) {
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    sub3(b, a, ab); sub3(c, a, ac);
    ap[0] = -a[0]; ap[1] = -a[1]; ap[2] = -a[2];
    double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; return; }
    bp[0] = -b[0]; bp[1] = -b[1]; bp[2] = -b[2];
    double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); lam[0] = 1 - v; lam[1] = v; lam[2] = 0; return; }
    cp[0] = -c[0]; cp[1] = -c[1]; cp[2] = -c[2];
    double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); lam[0] = 1 - w; lam[1] = 0; lam[2] = w; return; }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); lam[0] = 0; lam[1] = 1 - w; lam[2] = w; return;
    }
    double sum = va + vb + vc;
    if (!(sum > 0.0)) {
        /* degenerate (collinear) triangle: the closest point is on one of the edges */
        const double *P[3] = {a, b, c};
        double best = 1e300;
        for (int e = 0; e < 3; e++) {
            const double *p = P[e], *q = P[(e + 1) % 3];
            double pq[3]; sub3(q, p, pq);
            double den = dot3(pq, pq), t = den > 0 ? -dot3(p, pq) / den : 0.0;
            t = t < 0 ? 0 : (t > 1 ? 1 : t);
            double x[3] = {p[0] + t * pq[0], p[1] + t * pq[1], p[2] + t * pq[2]};
            double dd = dot3(x, x);
            if (dd < best) { best = dd; lam[0] = lam[1] = lam[2] = 0; lam[e] = 1 - t; lam[(e + 1) % 3] = t; }
        }
        return;
    }
    double den = 1.0 / sum, v = vb * den, w = vc * den;
    lam[0] = 1 - v - w; lam[1] = v; lam[2] = w;
}

/* origin strictly on the other side of plane(a,b,c) from d?  (degenerate -> treated as outside) */
static int origin_outside(const double *a, const double *b, const double *c, const double *d) {
    double ab[3], ac[3], n[3], ad[3];
    sub3(b, a, ab); sub3(c, a, ac); cross3(ab, ac, n); sub3(d, a, ad);
    double signp = -dot3(a, n), signd = dot3(ad, n);
    return signp * signd <= 0.0 ? 1 : 0;
}

/* Reduce simplex W (n points) to the sub-simplex supporting the point closest to the origin; v = that point.
 * Returns 1 if the origin is enclosed (n == 4 and inside). */
static int closest_simplex(double W[4][3], int *n, double v[3]) {
    double lam[4] = {0, 0, 0, 0};
    if (*n == 1) { lam[0] = 1; }
    else if (*n == 2) {
        double ab[3]; sub3(W[1], W[0], ab);
        double den = dot3(ab, ab), t = den > 0 ? -dot3(W[0], ab) / den : 0.0;
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        lam[0] = 1 - t; lam[1] = t;
    } else if (*n == 3) {
        closest_tri(W[0], W[1], W[2], lam);
    } else {
        static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
        double best = 1e300; int any = 0;
        /* a (numerically) flat tetrahedron encloses nothing: look at all four faces */
        double e1[3], e2[3], e3[3], cr[3];
        sub3(W[1], W[0], e1); sub3(W[2], W[0], e2); sub3(W[3], W[0], e3); cross3(e1, e2, cr);
        double det = dot3(cr, e3), L2 = fmax(fmax(dot3(e1, e1), dot3(e2, e2)), dot3(e3, e3));
        int flat = det * det <= 1e-24 * L2 * L2 * L2;
        for (int f = 0; f < 4; f++) {
            const int *id = F[f];
            if (!flat && !origin_outside(W[id[0]], W[id[1]], W[id[2]], W[id[3]])) continue;
            double l3[3], p[3];
            closest_tri(W[id[0]], W[id[1]], W[id[2]], l3);
            for (int k = 0; k < 3; k++) p[k] = l3[0] * W[id[0]][k] + l3[1] * W[id[1]][k] + l3[2] * W[id[2]][k];
            double dd = dot3(p, p);
            if (dd < best) {
                best = dd; any = 1;
                lam[0] = lam[1] = lam[2] = lam[3] = 0;
                lam[id[0]] = l3[0]; lam[id[1]] = l3[1]; lam[id[2]] = l3[2];
            }
        }
        if (!any) return 1;
    }
    double nv[3] = {0, 0, 0}; int m = 0; double Wn[4][3];
    for (int i = 0; i < *n; i++) {
        if (lam[i] > 0.0) {
            for (int k = 0; k < 3; k++) { nv[k] += lam[i] * W[i][k]; Wn[m][k] = W[i][k]; }
            m++;
        }
    }
    memcpy(W, Wn, sizeof(double) * 3 * m);
    *n = m; v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
    return 0;
}

/* distance between the cores of A and B; *deep = 1 when they intersect (returns 0). */
static double gjk_core_distance(const shape_t *A, const shape_t *B, int *deep, int *iters) {
    double W[4][3]; int n = 0;
    double v[3], sa[3], sb[3], w[3], d[3];
    d[0] = A->t[0] - B->t[0]; d[1] = A->t[1] - B->t[1]; d[2] = A->t[2] - B->t[2];
    if (dot3(d, d) < 1e-20) { d[0] = 1; d[1] = 0; d[2] = 0; }
    /* first point: support of A-B towards -d */
    double nd[3] = {-d[0], -d[1], -d[2]};
    support(A, nd, sa); support(B, d, sb); sub3(sa, sb, v);
    memcpy(W[0], v, sizeof(v)); n = 1;
    double vv = dot3(v, v), lb = 0.0;    /* lb: best proven lower bound (separating axis v: v.w / |v|) */
    *deep = 0;
    int it = 0;
    for (; it < g_flags.gjk_max_iter; it++) {
        if (vv < 1e-24) { *deep = 1; vv = 0; break; }
        double nv[3] = {-v[0], -v[1], -v[2]};
        support(A, nv, sa); support(B, v, sb); sub3(sa, sb, w);
        double delta = dot3(v, w);
        if (delta > 0 && delta / sqrt(vv) > lb) lb = delta / sqrt(vv);
        if (vv - delta <= g_flags.gjk_rel_tol * vv) break;            /* lower bound met */
        int dup = 0;
        for (int i = 0; i < n; i++) {
            double e[3]; sub3(w, W[i], e);
            if (dot3(e, e) <= 1e-30) dup = 1;
        }
        if (dup) break;
        memcpy(W[n], w, sizeof(w)); n++;
        double vnew[3];
        if (closest_simplex(W, &n, vnew)) {
            if (lb > 1e-9) break;          /* a separating axis exists: the enclosure is round-off */
            *deep = 1; vv = 0; break;
        }
        double vvn = dot3(vnew, vnew);
        if (vvn >= vv) break;                                          /* no progress (round-off) */
        memcpy(v, vnew, sizeof(v)); vv = vvn;
    }
    if (iters) *iters = it;
    return sqrt(vv);
}

/* ------------------------------------------------------------------ scene */
typedef struct {
    double q[6];
    double obs_pos[3], obs_quat[4];     /* obstacle cylinder base pose */
    double tgt_pos[3], tgt_quat[4];     /* target body pose */
    int tgt_type;                        /* 0 none/ghost (Ori), 1 sphere r=0.02 (Obs), 2 box he=0.025 (Sta/Dyn) */
    int has_obstacle;                    /* keys[5] == 'obstacle'  pyb_setup.py:398-399 */
    int geom;                            /* ORC_GEOM_HULL / ORC_GEOM_CAPSULE */
} orc_scene_t;

static double prim_margin(double he0, double he1, double he2) {
    if (g_flags.prim_margin_mode == 0) return 0.001;
    double m = he0 < he1 ? he0 : he1; m = m < he2 ? m : he2;
    double safe = 0.1 * m;
    return safe < 0.04 ? safe : 0.04;
}

static void make_box(shape_t *s, double cx, double cy, double cz, double hx, double hy, double hz) {
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double m = prim_margin(hx, hy, hz);
    s->type = SH_BOX; s->margin = m;
    s->he[0] = hx - m; s->he[1] = hy - m; s->he[2] = hz - m;
    memcpy(s->R, I, sizeof(I)); s->t[0] = cx; s->t[1] = cy; s->t[2] = cz;
}
/* create_table(1.1, 1.8, 0.92, x_offset=0.5, z_offset=-0.12)   reach.py:169, pyb_setup.py:802-811 */
static void make_table(shape_t *s) { make_box(s, 0.5, 0.0, -0.12 - 0.46, 0.55, 0.9, 0.46); }
/* create_track(0.2, 1.1, 0.12, x_offset=0, z_offset=0)          reach.py:170, pyb_setup.py:835-844 */
static void make_track(shape_t *s) { make_box(s, 0.0, 0.0, -0.06, 0.1, 0.55, 0.06); }

/* obstacle: cylinder radius 0.05, height 0.4, axis = local z    reach.py:279-283,427-431,626-630 */
static double g_obst_cap_verts[6] = {0, 0, -0.2, 0, 0, 0.2};
static void make_obstacle(const orc_scene_t *sc, shape_t *s) {
    if (sc->geom == ORC_GEOM_CAPSULE) {
        fit_init();
        g_obst_cap_verts[2] = -g_fit_obst_h; g_obst_cap_verts[5] = g_fit_obst_h;
        s->type = SH_HULL; s->verts = g_obst_cap_verts; s->nverts = 2; s->margin = 0.0;   /* pair margins: cap_pair_margin */
        quat_to_mat(sc->obs_quat, s->R); memcpy(s->t, sc->obs_pos, 3 * sizeof(double));
        return;
    }
    double m = prim_margin(0.05, 0.05, 0.2);
    s->type = SH_CYLZ; s->margin = m;
    s->he[0] = 0.05 - m; s->he[1] = 0.05 - m; s->he[2] = 0.2 - m;
    quat_to_mat(sc->obs_quat, s->R); memcpy(s->t, sc->obs_pos, 3 * sizeof(double));
}

static void make_target(const orc_scene_t *sc, shape_t *s) {
    if (sc->geom == ORC_GEOM_CAPSULE) {
        static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        s->type = SH_POINT; memcpy(s->R, I, sizeof(I));
        s->margin = (sc->tgt_type == 1 ? 0.02 : 0.025 * sqrt(3.0)) + 0.05;     /* target sphere + obstacle radius */
        memcpy(s->t, sc->tgt_pos, 3 * sizeof(double));
        return;
    }
    if (sc->tgt_type == 1) {               /* sphere radius 0.02: point core, margin = radius  reach.py:270-277 */
        static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        s->type = SH_POINT; s->margin = 0.02; memcpy(s->R, I, sizeof(I));
    } else {                               /* box half extents 0.025  reach.py:418-426,617-625 */
        double m = prim_margin(0.025, 0.025, 0.025);
        s->type = SH_BOX; s->margin = m; s->he[0] = s->he[1] = s->he[2] = 0.025 - m;
        quat_to_mat(sc->tgt_quat, s->R);
    }
    memcpy(s->t, sc->tgt_pos, 3 * sizeof(double));
}

static double g_cap_verts[7][6];
static int g_cap_init = 0;
static void make_link(const orc_scene_t *sc, int link, double pos[7][3], double rot[7][9], shape_t *s) {
    s->type = SH_HULL;
    if (sc->geom == ORC_GEOM_CAPSULE) {
        if (!g_cap_init) {
            for (int l = 0; l < 7; l++)
                for (int k = 0; k < 3; k++) {
                    g_cap_verts[l][k] = UR5E_CAPSULE_P0[3 * l + k];
                    g_cap_verts[l][3 + k] = UR5E_CAPSULE_P1[3 * l + k];
                }
            g_cap_init = 1;
        }
        s->verts = g_cap_verts[link]; s->nverts = 2; s->margin = 0.0;     /* pair margins: cap_pair_margin */
    } else {
        s->verts = &UR5E_HULL_VERTS[3 * UR5E_HULL_OFFSET[link]];
        s->nverts = UR5E_HULL_OFFSET[link + 1] - UR5E_HULL_OFFSET[link];
        s->margin = g_flags.hull_margin;
    }
    memcpy(s->R, rot[link], 9 * sizeof(double)); memcpy(s->t, pos[link], 3 * sizeof(double));
}

/* capsule geometry: margin of a pair class.  kind 0 = link vs obstacle, 1 = link vs table/track, 2 = self pair index */
static double cap_pair_margin(int kind, int l, int p) {
    fit_init();
    if (kind == 0) return g_fit_obst[l];
    if (kind == 1) return g_fit_box[l];
    return g_fit_self[p];
}
static double pair_distance(const shape_t *A, const shape_t *B, int *deep) {
    int dp = 0;
    double d = gjk_core_distance(A, B, &dp, 0);
    if (deep) *deep |= dp;
    return d - A->margin - B->margin;
}

/* getClosestPoints(UR5, obstacle, linkIndexA=link+2, distance=5.0)[0][8] for the five links 2..6
 * pyb_setup.py:439-456 */
int orc_link_distances(const orc_scene_t *sc, double out[5]) {
    double pos[7][3], rot[7][9]; shape_t L, O; int deep = 0;
    orc_fk(sc->q, pos, rot); make_obstacle(sc, &O);
    for (int i = 0; i < 5; i++) {
        int dp = 0;
        make_link(sc, i + 2, pos, rot, &L);
        out[i] = pair_distance(&L, &O, &dp);
        if (sc->geom == ORC_GEOM_CAPSULE) out[i] -= cap_pair_margin(0, i + 2, 0);
        if (dp) deep |= (1 << i);
    }
    return deep;
}

/* PyBullet.check_collision, same pair order and early-out        pyb_setup.py:382-429
 * returns 0 = no collision, else 1 + index of the first pair that reported distance <= 0.01:
 *   1..5   links 2..6 vs obstacle      6..10  links 2..6 vs table     11..15 links 2..6 vs track
 *   16..24 self pairs (1:3,1:4,1:5,1:6,2:4,2:5,2:6,3:5,3:6)
 * min_out (optional) receives the smallest distance seen among the pairs evaluated before returning;
 * all_out (optional, 24 doubles) receives every pair distance with no early-out. */
int orc_check_collision(const orc_scene_t *sc, double *all_out) {
    double pos[7][3], rot[7][9]; shape_t L, M, T;
    orc_fk(sc->q, pos, rot);
    int first = 0, idx = 0;
    const double thr = 0.01;
    if (sc->has_obstacle) {
        make_obstacle(sc, &T);
        for (int l = 2; l < 7; l++) {
            make_link(sc, l, pos, rot, &L);
            double d = pair_distance(&L, &T, 0);
            if (sc->geom == ORC_GEOM_CAPSULE) d -= cap_pair_margin(0, l, 0);
            if (all_out) all_out[idx] = d;
            if (d <= thr && !first) { first = 1 + idx; if (!all_out) return first; }
            idx++;
        }
    } else {
        if (all_out) for (int i = 0; i < 5; i++) all_out[i] = 1e9;
        idx = 5;
    }
    for (int o = 0; o < 2; o++) {
        if (o == 0) make_table(&T); else make_track(&T);
        for (int l = 2; l < 7; l++) {
            make_link(sc, l, pos, rot, &L);
            double d = pair_distance(&L, &T, 0);
            if (sc->geom == ORC_GEOM_CAPSULE) d -= cap_pair_margin(1, l, 0);
            if (all_out) all_out[idx] = d;
            if (d <= thr && !first) { first = 1 + idx; if (!all_out) return first; }
            idx++;
        }
    }
    int start = 3;
    for (int a = 1; a < 4; a++) {
        for (int b = start; b < 7; b++) {
            make_link(sc, a, pos, rot, &L); make_link(sc, b, pos, rot, &M);
            double d = pair_distance(&L, &M, 0);
            if (sc->geom == ORC_GEOM_CAPSULE) d -= cap_pair_margin(2, 0, idx - 15);
            if (all_out) all_out[idx] = d;
            if (d <= thr && !first) { first = 1 + idx; if (!all_out) return first; }
            idx++;
        }
        start++;
    }
    return first;
}

/* getClosestPoints(target, obstacle, distance=5)[0][8]            pyb_setup.py:431-437 */
double orc_target_obstacle_distance(const orc_scene_t *sc, int *deep) {
    shape_t T, O; int dp = 0;
    make_target(sc, &T); make_obstacle(sc, &O);
    double d = pair_distance(&T, &O, &dp);
    if (deep) *deep = dp;
    return d;
}

/* generic pair query used by unit tests: link (0..6) of the robot at q vs table(0)/track(1)/obstacle(2)/link(3+l) */
double orc_pair_distance(const orc_scene_t *sc, int link, int other, int *deep, int *iters) {
    double pos[7][3], rot[7][9]; shape_t L, O; int dp = 0;
    orc_fk(sc->q, pos, rot); make_link(sc, link, pos, rot, &L);
    if (other == 0) make_table(&O); else if (other == 1) make_track(&O);
    else if (other == 2) make_obstacle(sc, &O); else make_link(sc, other - 3, pos, rot, &O);
    double d = gjk_core_distance(&L, &O, &dp, iters) - L.margin - O.margin;
    if (deep) *deep = dp;
    return d;
}

/* calibration helper: the 24 pair distances of orc_check_collision between the SEGMENT CORES (no pair margins; the box
 * pairs keep the box margin), for the current capsule-fit obstacle half length */
void orc_capsule_core_distances(const orc_scene_t *sc_in, double out[24]) {
    orc_scene_t sc = *sc_in;
    sc.geom = ORC_GEOM_CAPSULE;
    double so[7], sb[7], ss[9], sh;
    fit_init();
    memcpy(so, g_fit_obst, sizeof(so)); memcpy(sb, g_fit_box, sizeof(sb)); memcpy(ss, g_fit_self, sizeof(ss)); sh = g_fit_obst_h;
    double z7[7] = {0}, z9[9] = {0};
    orc_set_capsule_fit(z7, z7, z9, sh);
    orc_check_collision(&sc, out);
    orc_set_capsule_fit(so, sb, ss, sh);
}

int orc_scene_sizeof(void) { return (int)sizeof(orc_scene_t); }
