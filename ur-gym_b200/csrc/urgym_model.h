// urgym_model.h -- host-side construction of the constants the kernels read from __constant__ memory, from the
// generated robot description (ur5e_model_data.h) and the scene the reach tasks create (reach.py / pyb_setup.py).
#pragma once
#include <math.h>
#include <string.h>

#include "ur5e_model_data.h"
#include "ur5e_hull_adjacency.h"
#include "urgym_capsule_fit.h"
#include "urgym_device.cuh"

namespace urgym {

// double-precision FK on the host, only to derive the constants of the (fixed) reset pose
static void host_fk(const double q[6], double pos[7][3], double R[7][9]) {
    for (int k = 0; k < 9; k++) R[0][k] = (k % 4 == 0) ? 1.0 : 0.0;
    pos[0][0] = pos[0][1] = pos[0][2] = 0.0;
    for (int j = 0; j < 6; j++) {
        const double *X = &UR5E_JOINT_XYZ[3 * j], *F = &UR5E_JOINT_ROT[9 * j];
        double A[9];
        for (int r = 0; r < 3; r++) {
            pos[j + 1][r] = pos[j][r] + R[j][3 * r] * X[0] + R[j][3 * r + 1] * X[1] + R[j][3 * r + 2] * X[2];
            for (int c = 0; c < 3; c++) A[3 * r + c] = R[j][3 * r] * F[c] + R[j][3 * r + 1] * F[3 + c] + R[j][3 * r + 2] * F[6 + c];
        }
        double c = cos(q[j]), s = sin(q[j]);
        for (int r = 0; r < 3; r++) {
            R[j + 1][3 * r] = A[3 * r] * c + A[3 * r + 1] * s;
            R[j + 1][3 * r + 1] = A[3 * r + 1] * c - A[3 * r] * s;
            R[j + 1][3 * r + 2] = A[3 * r + 2];
        }
    }
}

// the hull blob of urgym_device.cuh: vertices, adjacency offsets, adjacency (out: URGYM_HULL_BLOB_F4 float4)
static void build_hull_blob(float4 *out) {
    static_assert(UR5E_NUM_HULL_VERTS == URGYM_HULL_NV && UR5E_NUM_HULL_ADJ == URGYM_HULL_NADJ, "regenerate the hull tables");
    memset(out, 0, sizeof(float4) * URGYM_HULL_BLOB_F4);
    for (int i = 0; i < URGYM_HULL_NV; i++)
        out[i] = make_float4((float)UR5E_HULL_VERTS[3 * i], (float)UR5E_HULL_VERTS[3 * i + 1], (float)UR5E_HULL_VERTS[3 * i + 2], 0.0f);
    unsigned short *off = reinterpret_cast<unsigned short *>(out + URGYM_HULL_NV), *adj = off + URGYM_HULL_OFF_U16;
    for (int i = 0; i <= URGYM_HULL_NV; i++) off[i] = UR5E_HULL_ADJ_OFF[i];
    for (int i = 0; i < URGYM_HULL_NADJ; i++) adj[i] = UR5E_HULL_ADJ[i];
    // start table of the support function: exact support vertex of every cell-centre direction, per link
    unsigned short *dir = adj + URGYM_HULL_ADJ_U16;
    for (int l = 0; l < 7; l++)
        for (int cell = 0; cell < URGYM_HULL_DIR_CELLS; cell++) {
            double d[3];
            hull_cell_dir(cell, d);
            int best = 0; double bv = -1e300;
            for (int i = UR5E_HULL_OFFSET[l]; i < UR5E_HULL_OFFSET[l + 1]; i++) {
                const double v = UR5E_HULL_VERTS[3 * i] * d[0] + UR5E_HULL_VERTS[3 * i + 1] * d[1] + UR5E_HULL_VERTS[3 * i + 2] * d[2];
                if (v > bv) { bv = v; best = i - UR5E_HULL_OFFSET[l]; }
            }
            dir[l * URGYM_HULL_DIR_CELLS + cell] = (unsigned short)best;
        }
}

static void build_model_const(ModelConst &M) {
    memset(&M, 0, sizeof(M));
    for (int j = 0; j < 6; j++) {
        for (int k = 0; k < 3; k++) M.joint_xyz[j][k] = (float)UR5E_JOINT_XYZ[3 * j + k];
        for (int k = 0; k < 9; k++) M.joint_rot[j][k] = (float)UR5E_JOINT_ROT[9 * j + k];
    }
    for (int j = 0; j < 6; j++)
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++) M.joint_rot_p[j][k][c] = (float)UR5E_JOINT_ROT[9 * j + 3 * k + c];
    for (int l = 0; l < 7; l++)
        for (int k = 0; k < 3; k++) {
            M.cap_pp[l][k][0] = (float)UR5E_CAPSULE_P0[3 * l + k];
            M.cap_pp[l][k][1] = (float)UR5E_CAPSULE_P1[3 * l + k];
        }
    const double hull_margin = 0.001;                       // URDF mesh links, Bullet default collision margin
    for (int l = 0; l < 7; l++) {
        for (int k = 0; k < 3; k++) {
            M.cap_p0[l][k] = (float)UR5E_CAPSULE_P0[3 * l + k];
            M.cap_p1[l][k] = (float)UR5E_CAPSULE_P1[3 * l + k];
        }
        M.cap_m[l] = (float)(UR5E_CAPSULE_R[l] + hull_margin);
        double hl = 0.0;
        for (int k = 0; k < 3; k++) hl += (UR5E_CAPSULE_P1[3 * l + k] - UR5E_CAPSULE_P0[3 * l + k]) * (UR5E_CAPSULE_P1[3 * l + k] - UR5E_CAPSULE_P0[3 * l + k]);
        M.cap_hl[l] = (float)(0.5 * sqrt(hl) + 1e-5);
        M.cap_ia[l] = (float)(1.0 / hl);
    }
    for (int l = 0; l < 8; l++) M.hull_off[l] = UR5E_HULL_OFFSET[l];
    M.hull_margin = (float)hull_margin;
    const double pm = 0.001;                                // createCollisionShape primitives: margin 0.001, core shrunk
    // create_table(1.1, 1.8, 0.92, x_offset=0.5, z_offset=-0.12)   reach.py:169; pyb_setup.py:802-811
    const double tc[3] = {0.5, 0.0, -0.12 - 0.46}, th[3] = {0.55, 0.9, 0.46};
    // create_track(0.2, 1.1, 0.12, x_offset=0, z_offset=0)         reach.py:170; pyb_setup.py:835-844
    const double kc[3] = {0.0, 0.0, -0.06}, kh[3] = {0.1, 0.55, 0.06};
    for (int k = 0; k < 3; k++) {
        M.box_c[0][k] = (float)tc[k]; M.box_he[0][k] = (float)(th[k] - pm);
        M.box_c[1][k] = (float)kc[k]; M.box_he[1][k] = (float)(kh[k] - pm);
    }
    M.box_margin[0] = M.box_margin[1] = (float)pm;
    // obstacle cylinder radius 0.05, height 0.4                    reach.py:279-283,427-431,626-630
    M.obst_r = (float)(0.05 - pm); M.obst_h = (float)(0.2 - pm); M.obst_margin = (float)pm;
    M.tgt_box_he = (float)(0.025 - pm); M.tgt_box_margin = (float)pm;   // reach.py:418-426
    M.tgt_sphere_margin = 0.02f;                                         // reach.py:270-277
    // capsule mode: smallest bounding capsule of the cylinder (segment half length = half height, radius = radius);
    // target: the Obs sphere itself, the bounding sphere of the Sta/Dyn cube
    M.obst_cap_h = 0.2f; M.obst_cap_m = 0.05f;
    M.obst_cap_ie = (float)(1.0 / (0.4 * 0.4));
    for (int l = 0; l < 7; l++) { M.fit_obst[l] = (float)URGYM_FIT_OBST[l]; M.fit_box[l] = (float)URGYM_FIT_BOX[l]; }
    for (int k = 0; k < 9; k++) M.fit_self[k] = (float)URGYM_FIT_SELF[k];
    {
        const int L1[9] = {1, 1, 1, 1, 2, 2, 2, 3, 3}, L2[9] = {3, 4, 5, 6, 4, 5, 6, 5, 6};
        for (int p = 0; p < 9; p++) {
            // broad-phase threshold: midpoint-midpoint for the short-short pairs (1,4) (1,5) (1,6), midpoint of the short
            // member against the long member's segment for the others (robot_pass_capsule)
            const bool spheres = L1[p] == 1 && L2[p] != 3;
            const int ls = L1[p] == 1 ? 1 : L2[p];
            const float reach = 0.01f + M.fit_self[p];                                       // 0.01 = URGYM_COLLISION_MARGIN
            const float far = spheres ? reach + M.cap_hl[L1[p]] + M.cap_hl[L2[p]] : reach + M.cap_hl[ls];
            M.self_far2[p] = far * far;
            M.self_reach2[p] = reach * reach;
        }
    }
    // both boxes are centred on y = 0 and share one margin (reach.py:169-170): box_tests relies on it
    for (int l = 0; l < 8; l++) M.box_reach[l] = l < 7 ? 0.01f + M.fit_box[l] + M.box_margin[0] : 0.0f;   // 0.01 = URGYM_COLLISION_MARGIN
    for (int b = 0; b < 2; b++) {
        const float cx = M.box_c[b][0], hx = M.box_he[b][0], hy = M.box_he[b][1];
        float *K = M.box_k[b];
        K[0] = M.box_c[b][2] + M.box_he[b][2]; K[1] = cx; K[2] = hx; K[3] = hy;
        K[4] = cx - hx; K[5] = cx + hx; K[6] = -hy; K[7] = hy;
    }
    for (int b = 0; b < 2; b++) {
        float *S = M.sat2[b];
        S[0] = M.box_c[b][0]; S[1] = M.box_c[b][2]; S[2] = M.box_he[b][0]; S[3] = M.box_he[b][2];
        S[4] = 0.01f + M.fit_box[2] + M.box_margin[b];
    }
    M.fit_obst_h = (float)URGYM_FIT_OBST_H;
    M.fit_obst_ie = (float)(1.0 / (4.0 * URGYM_FIT_OBST_H * URGYM_FIT_OBST_H));
    M.box_top = (float)fmax(tc[2] + th[2] - pm, kc[2] + kh[2] - pm);
    M.tgt_cap_m[0] = 0.0f; M.tgt_cap_m[1] = 0.02f;
    M.tgt_cap_m[2] = M.tgt_cap_m[3] = (float)(0.025 * 1.7320508075688772);
    // reset pose                                                   UR5.py:262
    const double qn[6] = {0.0, -1.5708, 0.0, -1.5708, 0.0, 0.0};
    double pos[7][3], R[7][9];
    host_fk(qn, pos, R);
    for (int j = 0; j < 6; j++) M.neutral_q[j] = (float)qn[j];
    for (int l = 0; l < 7; l++) {
        for (int r = 0; r < 3; r++) {
            double a = pos[l][r], b = pos[l][r];
            for (int c = 0; c < 3; c++) {
                a += R[l][3 * r + c] * UR5E_CAPSULE_P0[3 * l + c];
                b += R[l][3 * r + c] * UR5E_CAPSULE_P1[3 * l + c];
            }
            M.neutral_ca[l][r] = (float)a; M.neutral_cb[l][r] = (float)b;
            M.neutral_p[l][r] = (float)pos[l][r];
        }
        for (int k = 0; k < 9; k++) M.neutral_R[l][k] = (float)R[l][k];
    }
    // EE pose = link 6 frame (ee_link: identity fixed joint, no inertial)   UR5.py:263,334-340
    {
        const double *E = R[6];
        double sarg = -E[6], roll, pitch, yaw;
        // |sarg| is far from the gimbal branch at the reset pose; regular getEulerFromQuaternion formulas
        pitch = asin(sarg); roll = atan2(E[7], E[8]); yaw = atan2(E[3], E[0]);
        M.neutral_ee[0] = (float)pos[6][0]; M.neutral_ee[1] = (float)pos[6][1]; M.neutral_ee[2] = (float)pos[6][2];
        M.neutral_ee[3] = (float)roll; M.neutral_ee[4] = (float)pitch; M.neutral_ee[5] = (float)yaw;
    }
}


}  // namespace urgym
