#!/usr/bin/env python3
"""Derive the numeric model of the UR5e reach scenes from the reference's asset files.

Reads (read-only, in THIS container only -- the GPU box has no /root/reference):
  UR_gym/envs/robots/urdf/ur5e.urdf                 joint origins  (ur5e.urdf:222-298)
                                                    collision-mesh placements (ur5e.urdf:75-213)
  UR_gym/envs/robots/meshes/ur5/collision/*.stl     collision meshes (binary STL, all convex)

Writes (committed, so the product and the oracle never touch /root/reference at run time):
  ur-gym_b200/csrc/ur5e_model_data.h   C arrays (double literals) shared by the CUDA library and the C oracle
  ur-gym_b200/assets/ur5e_model.npz    the same numbers for Python (tests, capsule/hull disagreement tools)

Nothing here is reference *source*: it is the robot description (numbers) that any implementation of
these tasks needs.  Every emitted table names the reference line it came from.

Usage:  python tools/extract_constants.py [/root/reference]
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull
from scipy.optimize import minimize

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
URDF = os.path.join(REF, "UR_gym/envs/robots/urdf/ur5e.urdf")

# PyBullet link index -> URDF link name (depth-first over joints in file order; UR5.py:258,263 use 1..6 and 7)
LINKS = ["base_link_inertia", "shoulder_link", "upper_arm_link", "forearm_link",
         "wrist_1_link", "wrist_2_link", "wrist_3_link"]
JOINTS = ["shoulder_pan_joint", "shoulder_lift_joint", "elbow_joint",
          "wrist_1_joint", "wrist_2_joint", "wrist_3_joint"]


def rpy_matrix(r, p, y):
    """URDF fixed-axis roll/pitch/yaw -> R = Rz(y) Ry(p) Rx(r)."""
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def read_stl(path):
    b = open(path, "rb").read()
    n = struct.unpack("<I", b[80:84])[0]
    assert len(b) == 84 + 50 * n, "not a binary STL: " + path
    rec = np.frombuffer(b[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return rec["v"].reshape(-1, 3)


def vec(s):
    return np.array([float(x) for x in s.split()], dtype=np.float64)


def min_enclosing_circle(pts):
    """Smallest circle containing 2-D points (tiny problem: Nelder-Mead polish of the midrange start)."""
    c0 = 0.5 * (pts.min(0) + pts.max(0))
    f = lambda c: np.max(np.hypot(pts[:, 0] - c[0], pts[:, 1] - c[1]))
    res = minimize(f, c0, method="Nelder-Mead", options={"xatol": 1e-9, "fatol": 1e-12, "maxiter": 4000})
    c = res.x
    return c, f(c)


def bounding_capsule(v):
    """Smallest-volume *bounding* capsule among a few candidate axes: every hull vertex lies inside it."""
    cands = [np.eye(3)[i] for i in range(3)]
    w, e = np.linalg.eigh(np.cov((v - v.mean(0)).T))
    cands.append(e[:, -1])
    best = None
    for ax in cands:
        ax = ax / np.linalg.norm(ax)
        # orthonormal frame (u, w, ax)
        tmp = np.eye(3)[np.argmin(np.abs(ax))]
        u = np.cross(ax, tmp); u /= np.linalg.norm(u)
        w2 = np.cross(ax, u)
        t = v @ ax
        p2 = np.stack([v @ u, v @ w2], 1)
        c, r = min_enclosing_circle(p2)
        r = r * (1 + 1e-9) + 1e-9
        rho = np.hypot(p2[:, 0] - c[0], p2[:, 1] - c[1])
        s = np.sqrt(np.maximum(r * r - rho * rho, 0.0))
        a, b = np.min(t + s), np.max(t - s)
        if a > b:
            a = b = 0.5 * (a + b)
            # grow the radius until the degenerate (sphere) capsule bounds everything
            centre = c[0] * u + c[1] * w2 + a * ax
            r = np.max(np.linalg.norm(v - centre, axis=1)) * (1 + 1e-9)
        vol = np.pi * r * r * (b - a) + 4.0 / 3.0 * np.pi * r ** 3
        p0 = c[0] * u + c[1] * w2 + a * ax
        p1 = c[0] * u + c[1] * w2 + b * ax
        if best is None or vol < best[0]:
            best = (vol, p0, p1, r)
    _, p0, p1, r = best
    # verify
    d = p1 - p0
    L2 = d @ d
    tt = np.clip(((v - p0) @ d) / L2, 0, 1) if L2 > 0 else np.zeros(len(v))
    dist = np.linalg.norm(v - (p0 + tt[:, None] * d), axis=1)
    assert dist.max() <= r + 1e-12
    return p0, p1, r


def main():
    root = ET.parse(URDF).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {j.get("name"): j for j in root.findall("joint")}

    jxyz, jrpy, jrot, jlim = [], [], [], []
    for jn in JOINTS:
        j = joints[jn]
        assert j.get("type") == "revolute" and np.allclose(vec(j.find("axis").get("xyz")), [0, 0, 1])
        o = j.find("origin")
        xyz, rpy = vec(o.get("xyz")), vec(o.get("rpy"))
        jxyz.append(xyz); jrpy.append(rpy); jrot.append(rpy_matrix(*rpy))
        lim = j.find("limit")
        jlim.append([float(lim.get("lower")), float(lim.get("upper"))])
    # fixed joints on the chain must be identity for the kinematics used here (ur5e.urdf:222-231, 294-298)
    for fj in ["base_link-base_link_inertia", "ee_fixed_joint"]:
        o = joints[fj].find("origin")
        assert np.allclose(vec(o.get("xyz")), 0) and np.allclose(vec(o.get("rpy")), 0)
    # ee_link has no <inertial> (ur5e.urdf:299-306): its reported frame is the link frame
    assert links["ee_link"].find("inertial") is None

    hull_v, hull_off, hull_faces, caps = [], [0], [], []
    for ln in LINKS:
        c = links[ln].find("collision")
        o = c.find("origin")
        mesh = c.find("geometry").find("mesh").get("filename")
        stl = os.path.normpath(os.path.join(os.path.dirname(URDF), mesh))
        v32 = np.unique(read_stl(stl), axis=0)                  # float32 vertices, de-duplicated
        v = v32.astype(np.float64)
        hull = ConvexHull(v)
        assert len(hull.vertices) == len(v), "mesh is not convex: " + ln
        R, t = rpy_matrix(*vec(o.get("rpy"))), vec(o.get("xyz"))
        vl = v @ R.T + t                                        # mesh frame -> link frame
        hull_v.append(vl)
        hull_off.append(hull_off[-1] + len(vl))
        hull_faces.append(hull.simplices.copy())
        caps.append(bounding_capsule(vl))
        print(f"{ln:20s} {os.path.basename(stl):14s} verts={len(vl):4d} capsule r={caps[-1][2]:.4f} "
              f"len={np.linalg.norm(caps[-1][1]-caps[-1][0]):.4f}")

    jxyz, jrpy, jrot, jlim = map(np.array, (jxyz, jrpy, jrot, jlim))
    allv = np.concatenate(hull_v)
    cap_p0 = np.array([c[0] for c in caps]); cap_p1 = np.array([c[1] for c in caps]); cap_r = np.array([c[2] for c in caps])

    os.makedirs(os.path.join(ROOT, "ur-gym_b200/assets"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "ur-gym_b200/assets/ur5e_model.npz"),
                        joint_xyz=jxyz, joint_rpy=jrpy, joint_rot=jrot, joint_limits=jlim,
                        hull_vertices=allv, hull_offsets=np.array(hull_off, dtype=np.int32),
                        capsule_p0=cap_p0, capsule_p1=cap_p1, capsule_r=cap_r,
                        **{f"hull_faces_{i}": f.astype(np.int32) for i, f in enumerate(hull_faces)})

    def arr(name, a, per_line=3, ctype="double"):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        out = [f"static const {ctype} {name}[{len(a)}] = {{"]
        for i in range(0, len(a), per_line):
            out.append("  " + ", ".join(repr(float(x)) for x in a[i:i + per_line]) + ",")
        out.append("};")
        return "\n".join(out)

    h = []
    h.append("/* GENERATED by tools/extract_constants.py -- do not edit.\n"
             " * Numeric description of the UR5e used by UR5{Ori,Obs,Sta,Dyn}Reach-v1, derived from the reference's\n"
             " * asset files (not from its source code):\n"
             " *   joint origins        UR_gym/envs/robots/urdf/ur5e.urdf:232-279 (revolute, axis z)\n"
             " *   collision placement  UR_gym/envs/robots/urdf/ur5e.urdf:75-213  (<collision><origin>)\n"
             " *   hull vertices        UR_gym/envs/robots/meshes/ur5/collision/<link>.stl (convex; de-duplicated float32\n"
             " *                        vertices, moved into the LINK frame in double precision)\n"
             " * Link index = PyBullet link index: 0 base_link_inertia, 1 shoulder, 2 upper_arm, 3 forearm,\n"
             " * 4 wrist_1, 5 wrist_2, 6 wrist_3 (7 = ee_link, same frame as 6: ur5e.urdf:294-298).\n"
             " * UR5E_CAPSULE_*: smallest bounding capsule of each hull found by the extractor (capsule geometry mode\n"
             " * and broad phase); not reference data.\n */")
    h.append("#ifndef UR5E_MODEL_DATA_H\n#define UR5E_MODEL_DATA_H\n")
    h.append("#define UR5E_NUM_JOINTS 6\n#define UR5E_NUM_LINKS 7")
    h.append(f"#define UR5E_NUM_HULL_VERTS {len(allv)}")
    h.append(f"#define UR5E_MAX_HULL_VERTS {max(len(v) for v in hull_v)}")
    h.append("static const int UR5E_HULL_OFFSET[8] = {" + ", ".join(str(x) for x in hull_off) + "};")
    h.append(arr("UR5E_JOINT_XYZ", jxyz))
    h.append(arr("UR5E_JOINT_RPY", jrpy))
    h.append("/* R = Rz(yaw) Ry(pitch) Rx(roll) of each joint origin, row-major 3x3 */")
    h.append(arr("UR5E_JOINT_ROT", jrot))
    h.append(arr("UR5E_JOINT_LIMITS", jlim, per_line=2))
    h.append(arr("UR5E_CAPSULE_P0", cap_p0))
    h.append(arr("UR5E_CAPSULE_P1", cap_p1))
    h.append(arr("UR5E_CAPSULE_R", cap_r, per_line=7))
    h.append("/* hull vertices, link frame, xyz interleaved; link L owns [UR5E_HULL_OFFSET[L], UR5E_HULL_OFFSET[L+1]) */")
    h.append(arr("UR5E_HULL_VERTS", allv))
    h.append("\n#endif")
    with open(os.path.join(ROOT, "ur-gym_b200/csrc/ur5e_model_data.h"), "w") as f:
        f.write("\n".join(h) + "\n")
    print("wrote header + npz;", len(allv), "hull vertices")


if __name__ == "__main__":
    main()
