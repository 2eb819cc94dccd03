/* ur_motor_oracle.c -- CPU restatement (double precision) of what the reference's MOTOR-DRIVEN robots ask of Bullet.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's cpu_baseline; the product never links or calls it.
 *
 * Reference path restated (UR_gym/envs/robots/UR5.py:10-118 `UR5`, :121-240 `UR5Reg`; UR_gym/pyb_setup.py):
 *   set_action           UR5.py:44-50,76-90    target = current joint angles + 0.1 * pi * clip(action)
 *   control_joints       pyb_setup.py:365-380  setJointMotorControlArray(POSITION_CONTROL, targetPositions, forces)
 *   sim.step()           pyb_setup.py:52-55    20 x stepSimulation at dt = 1/500, gravity (0, 0, -9.81) (pyb_setup.py:40-44)
 *   get_ee_position      UR5.py:112-114        getLinkState(ee_link = 6)[0]
 *   get_ee_velocity      UR5.py:116-118        getLinkState(6, computeLinkVelocity=True)[6]
 *
 * PARITY UNPINNED.  The arithmetic lives in the third-party `pybullet` wheel (setup.py:22, unpinned), which is neither in
 * /root/reference nor installable here; the reference ships no tests or golden vectors for this path and -- unlike the four
 * teleporting tasks -- no trained policy with published statistics either (Trained_Models/ holds Ori/Obs/Sta/Dyn only).
 * Everything below marked [RECALLED] is Bullet's published algorithm as remembered, not checked against a PyBullet run:
 *   - btMultiBodyDynamicsWorld substep: forward dynamics without motor torques (gravity, velocity-product terms, link
 *     damping) -> v_free = v + dt * qdd; constraint rows solved at velocity level by projected Gauss-Seidel
 *     (numSolverIterations = 50); then q += dt * v (semi-implicit Euler).                                   [RECALLED]
 *   - btMultiBodyJointMotor row (POSITION_CONTROL as pybullet sets it up: kp = 0.1, kd = 1.0, target velocity 0, erp 1,
 *     max impulse = force * dt):  desired joint velocity = kp * (target - q) / dt + v + kd * (0 - v),
 *     accumulated impulse clamped to +- force * dt.                                                         [RECALLED]
 *   - link damping: every link gets force -m v (k1 + k2 |v|) at its centre of mass and torque -(I w)(k1 + k2 |w|), with
 *     k1 = k2 = 0.04 (pybullet's default linearDamping / angularDamping).                                   [RECALLED]
 *   - a link without <inertial> (ee_link) is given mass 1 and inertia diag(1, 1, 1) by the URDF importer.  [RECALLED]
 * NOT restated: joint-limit rows (+-pi in ur5.urdf), contact rows (robot vs table / track).
 *
 * Formulation: world-frame Jacobians (mass matrix and bias assembled body by body).  The CUDA kernel uses link-frame
 * recursive Newton-Euler instead; agreement between the two is a check of the algebra, not of Bullet.
 */
#include <math.h>
#include <string.h>

#include "../ur-gym_b200/csrc/ur5_motor_data.h"

#define NB 7 /* bodies: links 1..6 and ee_link (fixed to link 6) */
#define NJ 6

typedef struct { double x, y, z; } v3;
static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 scl(double s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static double norm(v3 a) { return sqrt(dot(a, a)); }
static v3 mv(const double *R, v3 a) {
    return V(R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z, R[6] * a.x + R[7] * a.y + R[8] * a.z);
}
static void mm(const double *A, const double *B, double *C) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
static void axis_angle(v3 a, double q, double *R) { /* Rodrigues */
    double c = cos(q), s = sin(q), t = 1.0 - c;
    R[0] = t * a.x * a.x + c;       R[1] = t * a.x * a.y - s * a.z; R[2] = t * a.x * a.z + s * a.y;
    R[3] = t * a.x * a.y + s * a.z; R[4] = t * a.y * a.y + c;       R[5] = t * a.y * a.z - s * a.x;
    R[6] = t * a.x * a.z - s * a.y; R[7] = t * a.y * a.z + s * a.x; R[8] = t * a.z * a.z + c;
}

typedef struct {
    double Rw[NB][9];   /* world rotation of the link frame */
    v3 pw[NB];          /* world position of the link frame origin */
    v3 cw[NB];          /* world centre of mass */
    v3 aw[NJ];          /* world joint axis */
} Kin;

static void kinematics(const double *q, Kin *K) {
    double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    v3 pp = V(0, 0, 0);
    for (int i = 0; i < NB; i++) {
        double A[9], Rq[9];
        mm(Rp, &UR5M_JOINT_ROT[9 * i], A);
        K->pw[i] = add(pp, mv(Rp, V(UR5M_JOINT_XYZ[3 * i], UR5M_JOINT_XYZ[3 * i + 1], UR5M_JOINT_XYZ[3 * i + 2])));
        if (i < NJ) {
            v3 a = V(UR5M_JOINT_AXIS[3 * i], UR5M_JOINT_AXIS[3 * i + 1], UR5M_JOINT_AXIS[3 * i + 2]);
            axis_angle(a, q[i], Rq);
            mm(A, Rq, K->Rw[i]);
            K->aw[i] = mv(K->Rw[i], a);
        } else {
            memcpy(K->Rw[i], A, sizeof(A));
        }
        K->cw[i] = add(K->pw[i], mv(K->Rw[i], V(UR5M_LINK_COM[3 * i], UR5M_LINK_COM[3 * i + 1], UR5M_LINK_COM[3 * i + 2])));
        memcpy(Rp, K->Rw[i], sizeof(Rp));
        pp = K->pw[i];
    }
}
/* world inertia tensor times a vector: R diag(I) R^T a */
static v3 inertia_mul(const Kin *K, int i, v3 a) {
    const double *R = K->Rw[i];
    v3 l = V(R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z, R[2] * a.x + R[5] * a.y + R[8] * a.z);
    l = V(UR5M_LINK_INERTIA[3 * i] * l.x, UR5M_LINK_INERTIA[3 * i + 1] * l.y, UR5M_LINK_INERTIA[3 * i + 2] * l.z);
    return mv(R, l);
}
static int joint_of(int body) { return body < NJ ? body : NJ - 1; } /* last joint that moves the body */

void orm_mass_matrix(const double *q, double *M) {
    Kin K;
    kinematics(q, &K);
    memset(M, 0, sizeof(double) * 36);
    for (int i = 0; i < NB; i++)
        for (int j = 0; j <= joint_of(i); j++) {
            v3 jvj = cross(K.aw[j], sub(K.cw[i], K.pw[j]));
            v3 Iaj = inertia_mul(&K, i, K.aw[j]);
            for (int k = 0; k <= joint_of(i); k++) {
                v3 jvk = cross(K.aw[k], sub(K.cw[i], K.pw[k]));
                M[6 * j + k] += UR5M_LINK_MASS[i] * dot(jvj, jvk) + dot(K.aw[k], Iaj);
            }
        }
}
/* generalized bias forces: tau such that M qdd + tau = (applied joint torques); gravity, velocity products, link damping */
void orm_bias(const double *q, const double *qd, double gravity_z, double k_damp, double *tau) {
    Kin K;
    kinematics(q, &K);
    v3 w[NB], al[NB], ap[NB], vp[NB], F[NB], N[NB];
    v3 wp = V(0, 0, 0), alp = V(0, 0, 0), app = V(0, 0, 0), vpp = V(0, 0, 0), pp = V(0, 0, 0);
    for (int i = 0; i < NB; i++) {
        v3 d = sub(K.pw[i], pp);
        vp[i] = add(vpp, cross(wp, d));
        ap[i] = add(app, add(cross(alp, d), cross(wp, cross(wp, d))));
        if (i < NJ) {
            v3 rel = scl(qd[i], K.aw[i]);
            w[i] = add(wp, rel);
            al[i] = add(alp, cross(wp, rel));
        } else {
            w[i] = wp; al[i] = alp;
        }
        v3 r = sub(K.cw[i], K.pw[i]);
        v3 vc = add(vp[i], cross(w[i], r));
        v3 ac = add(ap[i], add(cross(al[i], r), cross(w[i], cross(w[i], r))));
        double m = UR5M_LINK_MASS[i];
        v3 Iw = inertia_mul(&K, i, w[i]);
        F[i] = sub(scl(m, ac), V(0, 0, m * gravity_z));
        N[i] = add(inertia_mul(&K, i, al[i]), cross(w[i], Iw));
        if (k_damp > 0.0) {     /* link damping [RECALLED]: force -m v (k + k |v|), torque -(I w)(k + k |w|) */
            F[i] = add(F[i], scl(m * (k_damp + k_damp * norm(vc)), vc));
            N[i] = add(N[i], scl(k_damp + k_damp * norm(w[i]), Iw));
        }
        wp = w[i]; alp = al[i]; app = ap[i]; vpp = vp[i]; pp = K.pw[i];
    }
    for (int j = 0; j < NJ; j++) {
        double t = 0.0;
        for (int i = j; i < NB; i++) t += dot(cross(K.aw[j], sub(K.cw[i], K.pw[j])), F[i]) + dot(K.aw[j], N[i]);
        tau[j] = t;
    }
}
static void invert6(const double *M, double *Minv) { /* Gauss-Jordan with partial pivoting (M is SPD) */
    double A[6][12];
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) { A[r][c] = M[6 * r + c]; A[r][6 + c] = r == c ? 1.0 : 0.0; }
    for (int c = 0; c < 6; c++) {
        int p = c;
        for (int r = c + 1; r < 6; r++) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (p != c) for (int k = 0; k < 12; k++) { double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
        double inv = 1.0 / A[c][c];
        for (int k = 0; k < 12; k++) A[c][k] *= inv;
        for (int r = 0; r < 6; r++) if (r != c) { double f = A[r][c]; for (int k = 0; k < 12; k++) A[r][k] -= f * A[c][k]; }
    }
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) Minv[6 * r + c] = A[r][6 + c];
}
/* n_sub x stepSimulation with position motors (see the header).  q, qd in / out.  force == NULL: no motors (free motion). */
void orm_substeps(double *q, double *qd, const double *target, const double *force, int n_sub, double dt, double kp, double kd,
                  int iters, double gravity_z, double k_damp) {
    for (int s = 0; s < n_sub; s++) {
        double M[36], Minv[36], tau[6], v[6];
        orm_mass_matrix(q, M);
        invert6(M, Minv);
        orm_bias(q, qd, gravity_z, k_damp, tau);
        for (int i = 0; i < 6; i++) {
            double a = 0.0;
            for (int k = 0; k < 6; k++) a -= Minv[6 * i + k] * tau[k];
            v[i] = qd[i] + dt * a;
        }
        if (force) {
            double want[6], lam[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 6; i++) want[i] = kp * (target[i] - q[i]) / dt + v[i] + kd * (0.0 - v[i]);
            for (int it = 0; it < iters; it++)
                for (int i = 0; i < 6; i++) {
                    double dl = (want[i] - v[i]) / Minv[6 * i + i], l0 = lam[i], lim = force[i] * dt;
                    double l1 = l0 + dl;
                    l1 = l1 > lim ? lim : (l1 < -lim ? -lim : l1);
                    dl = l1 - l0; lam[i] = l1;
                    for (int k = 0; k < 6; k++) v[k] += Minv[6 * k + i] * dl;
                }
        }
        for (int i = 0; i < 6; i++) { qd[i] = v[i]; q[i] += dt * v[i]; }
    }
}
/* ee_link world position and linear velocity */
void orm_ee_state(const double *q, const double *qd, double *pos, double *vel) {
    Kin K;
    kinematics(q, &K);
    v3 w = V(0, 0, 0), vp = V(0, 0, 0), pp = V(0, 0, 0);
    for (int i = 0; i < NB; i++) {
        vp = add(vp, cross(w, sub(K.pw[i], pp)));
        if (i < NJ) w = add(w, scl(qd[i], K.aw[i]));
        pp = K.pw[i];
    }
    pos[0] = K.pw[NB - 1].x; pos[1] = K.pw[NB - 1].y; pos[2] = K.pw[NB - 1].z;
    vel[0] = vp.x; vel[1] = vp.y; vel[2] = vp.z;
}
/* total mechanical energy (tests: conserved by free motion without damping as dt -> 0) */
double orm_energy(const double *q, const double *qd, double gravity_z) {
    double M[36], e = 0.0;
    Kin K;
    orm_mass_matrix(q, M);
    kinematics(q, &K);
    for (int i = 0; i < 6; i++) for (int k = 0; k < 6; k++) e += 0.5 * qd[i] * M[6 * i + k] * qd[k];
    for (int i = 0; i < NB; i++) e -= UR5M_LINK_MASS[i] * gravity_z * K.cw[i].z;
    return e;
}
