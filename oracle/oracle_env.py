"""oracle/oracle_env.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of UR-gym's reach-task step/reset path, used only as the *checker* by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
(ur-gym_b200/) never imports this module.

Three provenance tiers, kept visibly apart:

  T0  reference code executed as-is: UR_gym/utils.py (numpy + scipy only).  It cannot travel to the GPU box, so
      tests/golden/make_golden.py runs it HERE and commits its outputs; `distance` / `angular_distance` /
      the Euler samplers below are checked against those vectors (tests/test_oracle_golden.py).
  T1  restatement of the reference's own *Python* (this file): step ordering, observation layout, rewards,
      success, termination, TimeLimit, samplers, rejection rules, obstacle schedule.  Each method names the
      reference lines it follows.
  T2  restatement of *Bullet / PyBullet* behaviour (oracle/ur_oracle_sim.c): FK, Euler extraction, quaternion
      difference / axis-angle, convex distance with margins, kinematic base integration.
      **Parity unpinned**: PyBullet is absent from this environment and the reference ships no golden vectors,
      so T2 is self-consistent but anchored to no PyBullet run.  Its assumptions are flags (OracleSim kwargs and
      orc_set_flags).

The random draws of `reset()` come from a `UniformStream`.  The reference mixes two numpy generators
(`task.np_random` for positions, the global `np.random` for Euler angles: reach.py:207-208, utils.py:82-98) whose
streams cannot be reproduced on a GPU; parity with the CUDA path therefore uses `PhiloxStream`, the same
counter-based stream the kernels use, while `NumpyStream` draws like the reference does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libur_oracle.so")

GEOM_HULL, GEOM_CAPSULE = 0, 1
# what PyBullet.get_link_distances measures (pyb_setup.py:439-456).  LD_OBSTACLE: the code as the reference ships it --
# links 2..6 against the obstacle only.  LD_WORKBENCH: what the method's own docstring describes ("Check the distance
# between workbench, obstacle and UR5 ... Array of robot links to any obstacle"): per link the smallest of the
# distances to the obstacle, the table and the track.  The shipped UR5ObsReach / UR5StaReach policies (September 2023,
# nine months older than the shipped code and the UR5DynReach policy) only reproduce their published success rates
# with the second definition; see DESIGN.md section 2.
LD_OBSTACLE, LD_WORKBENCH = 0, 1


def build_oracle(force: bool = False) -> str:
    if force and os.path.exists(_LIB_PATH):
        os.remove(_LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", _HERE])        # no-op when up to date (sources + generated headers)
    return _LIB_PATH


class _Scene(ctypes.Structure):
    _fields_ = [("q", ctypes.c_double * 6),
                ("obs_pos", ctypes.c_double * 3), ("obs_quat", ctypes.c_double * 4),
                ("tgt_pos", ctypes.c_double * 3), ("tgt_quat", ctypes.c_double * 4),
                ("tgt_type", ctypes.c_int), ("has_obstacle", ctypes.c_int), ("geom", ctypes.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        assert _lib.orc_scene_sizeof() == ctypes.sizeof(_Scene)
        _lib.orc_target_obstacle_distance.restype = ctypes.c_double
        _lib.orc_pair_distance.restype = ctypes.c_double
        _lib.orc_set_flags.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        _lib.orc_set_capsule_fit.argtypes = [ctypes.POINTER(ctypes.c_double)] * 3 + [ctypes.c_double]
    return _lib


def _d(n):
    return (ctypes.c_double * n)()


# --------------------------------------------------------------------------------------------------------------
# T0-pinned math: restatement of UR_gym/utils.py
# --------------------------------------------------------------------------------------------------------------
def distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils.py:5-31 -- L2 norm of the first three components; result reshaped to (n,)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    diff = a[..., :3] - b[..., :3]
    return np.sqrt((diff ** 2).sum(-1)).reshape(-1)


def _quat_ZYX(e: np.ndarray) -> np.ndarray:
    """scipy Rotation.from_euler('ZYX', e).as_quat(): intrinsic z-y'-x'' with e = (z, y, x) angles, i.e.
    R = Rz(e0) Ry(e1) Rx(e2); quaternion (x, y, z, w).  utils.py:48-54 (quirk Q2: the triple fed in is
    PyBullet's (roll, pitch, yaw), so roll is used as the z angle)."""
    e = np.asarray(e, dtype=np.float64)
    hz, hy, hx = e[..., 0] / 2, e[..., 1] / 2, e[..., 2] / 2
    cz, sz, cy, sy, cx, sx = np.cos(hz), np.sin(hz), np.cos(hy), np.sin(hy), np.cos(hx), np.sin(hx)
    return np.stack([cz * cy * sx - sz * sy * cx,
                     cz * sy * cx + sz * cy * sx,
                     sz * cy * cx - cz * sy * sx,
                     cz * cy * cx + sz * sy * sx], -1)


def angular_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils.py:34-69 -- 2*arccos(|<qa, qb>|), dot clipped to [-1, 1]."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    dot = np.sum(_quat_ZYX(a[..., 3:]) * _quat_ZYX(b[..., 3:]), axis=-1)
    return (2 * np.arccos(np.abs(np.clip(dot, -1.0, 1.0)))).reshape(-1)


def euler_constrained_from_uniform(u_roll: float, u_yaw: float) -> np.ndarray:
    """utils.py:81-86 -- np.random.uniform(-90,-180), 0, np.random.uniform(0,-180), then deg2rad.
    numpy's uniform(low, high) is low + (high-low)*u for u in [0,1), also when low > high."""
    roll = -90.0 + (-180.0 - -90.0) * u_roll
    yaw = 0.0 + (-180.0 - 0.0) * u_yaw
    return np.deg2rad([roll, 0, yaw])


def euler_obstacle_from_uniform(u_sign: float, u_roll: float, u_pitch: float) -> np.ndarray:
    """utils.py:88-101 -- np.random.choice(['negative','positive'], p=[.5,.5]) takes 'negative' when its
    uniform draw is < 0.5; roll in -/+ U(30,150) deg; pitch in -U(30,150) if |roll| > 90 else +U(30,150)."""
    if u_sign < 0.5:
        roll = -30.0 + (-150.0 - -30.0) * u_roll
    else:
        roll = 30.0 + (150.0 - 30.0) * u_roll
    if roll < -90 or roll > 90:
        pitch = -30.0 + (-150.0 - -30.0) * u_pitch
    else:
        pitch = 30.0 + (150.0 - 30.0) * u_pitch
    return np.deg2rad([roll, pitch, 0])


# --------------------------------------------------------------------------------------------------------------
# uniform streams
# --------------------------------------------------------------------------------------------------------------
_PH_M0, _PH_M1, _PH_W0, _PH_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox-4x32-10 (Salmon et al., SC'11).  counter: 4 u32, key: 2 u32 -> 4 u32.  Same function as
    ur-gym_b200/csrc (philox4x32_10 in urgym_kernels.cu)."""
    c0, c1, c2, c3 = (int(x) & _M32 for x in counter)
    k0, k1 = (int(x) & _M32 for x in key)
    for r in range(10):
        if r:
            k0 = (k0 + _PH_W0) & _M32
            k1 = (k1 + _PH_W1) & _M32
        p0, p1 = _PH_M0 * c0, _PH_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
    return c0, c1, c2, c3


class UniformStream:
    """u(slot) for rejection iteration k of the reset that env i performs at reset event e (the product counts its
    step/reset calls; every reset that happens inside call number e draws from position e)."""

    def begin(self, env_index: int, event: int, blocks_per_iter: int) -> None:
        raise NotImplementedError

    def iteration(self, k: int) -> None:
        raise NotImplementedError

    def u(self, slot: int) -> float:
        raise NotImplementedError


class PhiloxStream(UniformStream):
    """The kernels' stream: counter = (k*blocks_per_iter + slot//4, event, env_index lo, hi), key = seed (lo, hi);
    u = (x >> 8) * 2**-24, exactly representable in float32."""

    def __init__(self, seed: int):
        self.key = (seed & _M32, (seed >> 32) & _M32)

    def begin(self, env_index, event, blocks_per_iter):
        self.env, self.ep, self.bpi, self.k, self._cache = env_index, event, blocks_per_iter, 0, {}

    def iteration(self, k):
        self.k, self._cache = k, {}

    def u(self, slot):
        b = slot // 4
        if b not in self._cache:
            self._cache[b] = philox4x32_10((self.k * self.bpi + b, self.ep, self.env & _M32, (self.env >> 32) & _M32),
                                           self.key)
        return float(self._cache[b][slot % 4] >> 8) * 2.0 ** -24


class NumpyStream(UniformStream):
    """Draws the way the reference does: a fresh double per call, in call order (slots are ignored)."""

    def __init__(self, seed: Optional[int] = None):
        self.rng = np.random.default_rng(seed)

    def begin(self, env_index, event, blocks_per_iter):
        pass

    def iteration(self, k):
        pass

    def u(self, slot):
        return float(self.rng.random())


# --------------------------------------------------------------------------------------------------------------
# T2: stand-in for `class PyBullet` (pyb_setup.py:15)
# --------------------------------------------------------------------------------------------------------------
class OracleSim:
    """Holds the bodies the reach scenes create and answers the queries pyb_setup.PyBullet answers on the hot path.

    T2 flags (each is an assumption about Bullet that a PyBullet run could overturn):
      hold_pose=True       after resetJointState the arm does not move during the 20 substeps: default velocity
                           motors (max impulse 1.0/substep) hold it against gravity; contact response on a
                           penetrating teleport is NOT modelled
      joint_limits=False   URDF limits are not enforced by the teleport path
    """

    def __init__(self, geom: int = GEOM_HULL, n_substeps: int = 20, link_dist_mode: int = LD_OBSTACLE):
        self.geom = geom
        self.link_dist_mode = link_dist_mode
        self.n_substeps = n_substeps                  # pyb_setup.py:25
        self.timestep = 1.0 / 500                     # pyb_setup.py:40
        self._bodies: Dict[str, dict] = {}            # insertion order == _bodies_idx order
        self.q = np.zeros(6)
        self.hold_pose = True
        self.joint_limits = False
        self.last_collision_pair = 0

    @property
    def dt(self):                                     # pyb_setup.py:47-50
        return self.timestep * self.n_substeps

    # -- body bookkeeping (create_* at pyb_setup.py:510-905 only matter through order, shape and pose)
    def add_body(self, name: str, kind: str, position=(0.0, 0.0, 0.0)) -> None:
        self._bodies[name] = dict(kind=kind, pos=np.array(position, dtype=np.float64),
                                  quat=np.array([0.0, 0.0, 0.0, 1.0]), v=np.zeros(3), w=np.zeros(3))

    def step(self) -> None:
        """pyb_setup.py:52-55 -- 20 x stepSimulation.  Robot: hold_pose.  Mass-0 bodies with a base velocity move
        kinematically."""
        for b in self._bodies.values():
            if np.any(b["v"] != 0) or np.any(b["w"] != 0):
                p, qt = _d(3), _d(4)
                p[:] = b["pos"]; qt[:] = b["quat"]
                v, w = _d(3), _d(3)
                v[:] = b["v"]; w[:] = b["w"]
                lib().orc_integrate_base(p, qt, v, w, ctypes.c_double(self.timestep), self.n_substeps)
                b["pos"], b["quat"] = np.array(p[:]), np.array(qt[:])

    # -- joints / links
    def set_joint_angles(self, angles) -> None:       # pyb_setup.py:319-338 (resetJointState x6)
        self.q = np.array(angles, dtype=np.float64)

    def get_joint_angle(self, joint: int) -> float:   # pyb_setup.py:281-291, joint indices 1..6
        return float(self.q[joint - 1])

    def _ee(self):
        q, p, e = _d(6), _d(3), _d(3)
        q[:] = self.q
        lib().orc_ee_pose(q, p, e)
        return np.array(p[:]), np.array(e[:])

    def get_link_position(self, link: int) -> np.ndarray:          # pyb_setup.py:221-232
        assert link == 7, "the reach path only reads the ee_link (UR5.py:263)"
        return self._ee()[0]

    def get_link_orientation(self, link: int) -> np.ndarray:       # pyb_setup.py:234-253, type="euler"
        assert link == 7
        return self._ee()[1]

    # -- bases
    def euler_to_quaternion(self, euler) -> np.ndarray:            # pyb_setup.py:151-152
        e, q = _d(3), _d(4)
        e[:] = euler
        lib().orc_quat_from_euler(e, q)
        return np.array(q[:])

    def set_base_pose(self, body: str, position, orientation) -> None:   # pyb_setup.py:305-317
        if len(orientation) == 3:
            orientation = self.euler_to_quaternion(orientation)
        b = self._bodies[body]
        b["pos"] = np.array(position, dtype=np.float64)
        b["quat"] = np.array(orientation, dtype=np.float64)

    def get_base_position(self, body: str) -> np.ndarray:          # pyb_setup.py:154-164
        return self._bodies[body]["pos"].copy()

    def get_base_rotation(self, body: str) -> np.ndarray:          # pyb_setup.py:179-195, type="euler"
        q, e = _d(4), _d(3)
        q[:] = self._bodies[body]["quat"]
        lib().orc_euler_from_quat(q, e)
        return np.array(e[:])

    def set_velocity(self, body: str, linear_velocity, angular_velocity) -> None:   # pyb_setup.py:340-349
        self._bodies[body]["v"] = np.array(linear_velocity, dtype=np.float64)
        self._bodies[body]["w"] = np.array(angular_velocity, dtype=np.float64)

    def get_quaternion_difference(self, start_quaternion, end_quaternion) -> np.ndarray:   # pyb_setup.py:351-359
        s, e, o = _d(4), _d(4), _d(4)
        s[:] = start_quaternion; e[:] = end_quaternion
        lib().orc_quat_difference(s, e, o)
        return np.array(o[:])

    def get_axis_angle(self, relative_rotation):                   # pyb_setup.py:361-363
        q, ax, ang = _d(4), _d(3), ctypes.c_double()
        q[:] = relative_rotation
        lib().orc_axis_angle(q, ax, ctypes.byref(ang))
        return np.array(ax[:]), ang.value

    # -- closest-point queries
    def _scene(self) -> _Scene:
        sc = _Scene()
        sc.q[:] = self.q
        keys = list(self._bodies)
        # pyb_setup.py:398-399: the obstacle loop runs only when the 6th registered body (UR5 is the first) is it
        sc.has_obstacle = int(len(keys) >= 5 and keys[4] == "obstacle")
        if "obstacle" in self._bodies:
            sc.obs_pos[:] = self._bodies["obstacle"]["pos"]; sc.obs_quat[:] = self._bodies["obstacle"]["quat"]
        else:
            sc.obs_quat[:] = [0, 0, 0, 1]
        t = self._bodies["target"]
        sc.tgt_pos[:] = t["pos"]; sc.tgt_quat[:] = t["quat"]
        sc.tgt_type = {"ghost": 0, "sphere": 1, "box": 2}[t["kind"]]
        sc.geom = self.geom
        return sc

    def check_collision(self) -> bool:                             # pyb_setup.py:382-429
        self.last_collision_pair = lib().orc_check_collision(ctypes.byref(self._scene()), None)
        return self.last_collision_pair != 0

    def all_pair_distances(self) -> np.ndarray:
        out = _d(24)
        lib().orc_check_collision(ctypes.byref(self._scene()), out)
        return np.array(out[:])

    def capsule_core_distances(self) -> np.ndarray:
        """segment-core distances of the 24 pairs (capsule geometry without its pair margins): calibration input"""
        out = _d(24)
        lib().orc_capsule_core_distances(ctypes.byref(self._scene()), out)
        return np.array(out[:])

    def get_target_to_obstacle_distance(self) -> float:            # pyb_setup.py:431-437
        deep = ctypes.c_int()
        return lib().orc_target_obstacle_distance(ctypes.byref(self._scene()), ctypes.byref(deep))

    def get_link_distances(self) -> np.ndarray:                    # pyb_setup.py:439-456
        out = _d(5)
        self.last_deep_mask = lib().orc_link_distances(ctypes.byref(self._scene()), out)
        d = np.array(out[:])
        if self.link_dist_mode == LD_WORKBENCH:         # links 2..6 vs table (pairs 5..9) and track (pairs 10..14)
            a = self.all_pair_distances()
            d = np.minimum(d, np.minimum(a[5:10], a[10:15]))
        return d


# --------------------------------------------------------------------------------------------------------------
# T1: robot (UR5.py:243-351)
# --------------------------------------------------------------------------------------------------------------
class UR5Ori:
    def __init__(self, sim: OracleSim):
        self.sim = sim
        self.action = np.zeros(6)
        self.action_low = -np.ones(6, dtype=np.float32)            # spaces.Box(-1, 1, (6,), float32)  UR5.py:251
        self.action_high = np.ones(6, dtype=np.float32)
        self.joint_indices = np.array([1, 2, 3, 4, 5, 6])          # UR5.py:258
        self.neutral_joint_values = np.array([0.0, -1.5708, 0.0, -1.5708, 0.0, 0.0])   # UR5.py:262
        self.ee_link = 7                                           # UR5.py:263

    def set_action(self, action: np.ndarray) -> None:
        """UR5.py:273-279.  dtype follows numpy: a float32 action stays float32 through clip, *pi and *0.1, and is
        widened only when added to the float64 joint angles (UR5.py:275-276,314-317)."""
        action = np.clip(np.array(action, copy=True), self.action_low, self.action_high)
        self.action = action[:6] * np.pi
        self.sim.set_joint_angles(self.get_joint_angles() + self.action * 0.1)

    def get_obs(self) -> np.ndarray:                               # UR5.py:320-325
        return np.concatenate((self.get_ee_position(), self.get_ee_orientation(), self.get_joint_angles()))

    def reset(self) -> None:                                       # UR5.py:327-332
        self.sim.set_joint_angles(self.neutral_joint_values)

    def set_joint_angles(self, angles) -> None:                    # core.py:161-167
        self.sim.set_joint_angles(angles)

    def get_ee_position(self) -> np.ndarray:                       # UR5.py:334-336
        return self.sim.get_link_position(self.ee_link)

    def get_ee_orientation(self) -> np.ndarray:                    # UR5.py:338-340
        return self.sim.get_link_orientation(self.ee_link)

    def get_joint_angles(self) -> np.ndarray:                      # UR5.py:346-351
        return np.array([self.sim.get_joint_angle(j) for j in self.joint_indices])


# --------------------------------------------------------------------------------------------------------------
# T1: tasks (reach.py:141-785).  One class, four parameter sets; every branch names its reference lines.
# --------------------------------------------------------------------------------------------------------------
_W = np.array([8, 2.4, 1.2, 1.2, 0.2])
TASKS = {
    "UR5OriReach-v1": dict(kind="Ori", goal_low=[0.3, -0.5, 0.0], goal_high=[0.75, 0.5, 0.2],      # reach.py:151-152
                           goal_dim=6, bpi=2),
    "UR5ObsReach-v1": dict(kind="Obs", goal_low=[0.3, -0.5, -0.1], goal_high=[0.75, 0.5, 0.2],     # reach.py:248-249
                           obs_low=[0.5, -0.5, 0.25], obs_high=[1.0, 0.5, 0.55],                   # reach.py:250-251
                           goal_dim=3, bpi=3),
    "UR5StaReach-v1": dict(kind="Sta", goal_low=[0.3, -0.5, 0.0], goal_high=[0.75, 0.5, 0.2],      # reach.py:385-386
                           obs_low=[0.5, -0.5, 0.25], obs_high=[1.0, 0.5, 0.55],                   # reach.py:387-388
                           goal_dim=6, bpi=3),
    "UR5DynReach-v1": dict(kind="Dyn", goal_low=[0.4, -0.5, 0.0], goal_high=[0.75, 0.5, 0.2],      # reach.py:584-585
                           obs_low=[0.5, -0.8, 0.25], obs_high=[1.2, 0.8, 0.75],                   # reach.py:586-587
                           goal_dim=6, bpi=5),
}
# uniform-draw slots inside one rejection iteration (shared with the kernels; see DESIGN.md "reset stream")
SLOTS = {
    "Ori": dict(goal=0, goal_roll=3, goal_yaw=4),
    "Obs": dict(goal=0, obs=3, obs_sign=6, obs_roll=7, obs_pitch=8),
    "Sta": dict(goal=0, goal_roll=3, goal_yaw=4, obs=5, obs_sign=8, obs_roll=9, obs_pitch=10),
    "Dyn": dict(start=0, end=3, goal=6, goal_roll=9, goal_yaw=10,
                start_sign=11, start_roll=12, start_pitch=13, end_sign=14, end_roll=15, end_pitch=16),
}
MAX_RESET_ITERS = 256   # the kernels give up after this many rejections and keep the last draw (P ~ 1e-21)


class ReachTask:
    def __init__(self, sim: OracleSim, robot: UR5Ori, env_id: str, stream: UniformStream, env_index: int = 0):
        p = TASKS[env_id]
        self.sim, self.robot, self.kind, self.p = sim, robot, p["kind"], p
        self.stream, self.env_index, self.episode = stream, env_index, 0
        self.goal_range_low, self.goal_range_high = np.array(p["goal_low"]), np.array(p["goal_high"])
        if self.kind != "Ori":
            self.obs_range_low, self.obs_range_high = np.array(p["obs_low"]), np.array(p["obs_high"])
        self.distance_threshold = 0.05                                 # reach.py:148,247,391,590
        self.ori_distance_threshold = 0.0873                           # reach.py:149,392,591
        self.collision_weight, self.success_weight = -500, 200         # reach.py:154-155,253-254,394,398
        self.distance_weight = -100 if self.kind == "Obs" else -70     # reach.py:156,255,395,594
        self.orientation_weight = -30                                  # reach.py:157,396,595
        self.obs_distance_weight = 100                                 # reach.py:256
        self.dist_change_weight = _W / np.sum(_W) * 50                 # reach.py:397-398,596-597
        self.goal = None
        self.obstacle = np.zeros(6)
        self.obstacle_start, self.obstacle_end = np.zeros(6), np.zeros(6)
        self.velocity = np.zeros(6)
        self.collision = False
        self.link_dist, self.last_dist = np.zeros(5), np.zeros(5)
        self.step_num = 0
        self.reset_iterations = 0
        self._create_scene()

    def _create_scene(self) -> None:
        """reach.py:167-187, 266-305, 414-452, 613-651.  Order matters: check_collision looks at the 6th body."""
        s = self.sim
        s.add_body("plane", "box"); s.add_body("table", "box"); s.add_body("track", "box")
        if self.kind == "Ori":
            s.add_body("target", "ghost", (0, 0, 1.0)); s.add_body("zone_goal", "ghost")
        else:
            s.add_body("target", "sphere" if self.kind == "Obs" else "box", (0, 0, 0) if self.kind == "Obs" else (0, 0, 1.0))
            s.add_body("obstacle", "cylinder", (0, 0, 1.0))
            s.add_body("zone_goal", "ghost"); s.add_body("zone_obs", "ghost")

    # ---- observations
    def get_obs(self) -> np.ndarray:
        if self.kind == "Ori":                                          # reach.py:189-190
            return np.array(self.goal)
        if self.kind == "Obs":                                          # reach.py:307-308
            return np.concatenate((self.goal, self.obstacle, self.link_dist))
        cur = np.concatenate((self.sim.get_base_position("obstacle"), self.sim.get_base_rotation("obstacle")))
        if self.kind == "Sta":                                          # reach.py:454-458
            return np.concatenate((self.goal, cur, self.link_dist))
        return np.concatenate((self.goal, cur, self.velocity, self.link_dist))   # reach.py:653-657

    def get_achieved_goal(self) -> np.ndarray:
        if self.kind == "Obs":                                          # reach.py:310-311
            return np.array(self.robot.get_ee_position())
        return np.concatenate((self.robot.get_ee_position(), self.robot.get_ee_orientation()))   # :192-195 etc.

    def get_goal(self) -> np.ndarray:                                   # core.py:206-211
        if self.goal is None:
            raise RuntimeError("No goal yet, call reset() first")
        return self.goal.copy()

    # ---- sampling
    def _uniform(self, low, high, slot0):
        """np_random.uniform(low, high) on 3-vectors: low + (high - low) * u, elementwise."""
        u = np.array([self.stream.u(slot0 + i) for i in range(3)])
        return low + (high - low) * u

    def _sample_goal(self, S) -> np.ndarray:
        pos = self._uniform(self.goal_range_low, self.goal_range_high, S["goal"])
        if self.kind == "Obs":                                          # reach.py:337-340
            return pos
        rot = euler_constrained_from_uniform(self.stream.u(S["goal_roll"]), self.stream.u(S["goal_yaw"]))
        return np.concatenate((pos, rot))                               # reach.py:206-210,505-510,715-720

    def _sample_obstacle(self, S, which="obs") -> np.ndarray:           # reach.py:342-346,512-516,722-726
        pos = self._uniform(self.obs_range_low, self.obs_range_high, S[which])
        rot = euler_obstacle_from_uniform(self.stream.u(S[which + "_sign"]), self.stream.u(S[which + "_roll"]),
                                          self.stream.u(S[which + "_pitch"]))
        return np.concatenate((pos, rot))

    # ---- reset
    def reset(self, event: Optional[int] = None) -> None:
        S = SLOTS[self.kind]
        self.episode += 1
        self.stream.begin(self.env_index, self.episode if event is None else event, self.p["bpi"])
        self.min_reject_margin = np.inf        # how close any rejection decision of this reset was to its threshold
        self.collision = False
        if self.kind == "Ori":                                          # reach.py:197-200
            self.stream.iteration(0)
            self.goal = self._sample_goal(S)
            self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
            return
        if self.kind == "Dyn":
            self.step_num = 0                                           # reach.py:666  (velocity is NOT cleared: Q4)
        k = 0
        while True:
            self.stream.iteration(k)
            if self.kind == "Dyn":                                      # reach.py:668-675
                self.goal = self._sample_goal(S)
                self.obstacle_start = self._sample_obstacle(S, "start")
                self.obstacle_end = self._sample_obstacle(S, "end")
                self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
                self.sim.set_base_pose("obstacle", self.obstacle_end[:3], self.obstacle_end[3:])
                start_end = distance(self.obstacle_end, self.obstacle_start)
                t2o = self.sim.get_target_to_obstacle_distance()
                self.min_reject_margin = min(self.min_reject_margin, abs(t2o - 0.1), abs(float(start_end[0]) - 1.0))
                fail = (t2o < 0.1) or bool(start_end < 1)
            else:                                                       # reach.py:316-321, 468-473
                self.goal = self._sample_goal(S)
                self.obstacle = self._sample_obstacle(S)
                if self.kind == "Obs":
                    self.sim.set_base_pose("target", self.goal, np.array([0.0, 0.0, 0.0, 1.0]))
                else:
                    self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
                self.sim.set_base_pose("obstacle", self.obstacle[:3], self.obstacle[3:])
                t2o = self.sim.get_target_to_obstacle_distance()
                self.min_reject_margin = min(self.min_reject_margin, abs(t2o - 0.1))
                fail = t2o < 0.1
            k += 1
            if not fail or k >= MAX_RESET_ITERS:
                break
        self.reset_iterations = k
        if self.kind == "Dyn":                                          # reach.py:677-678
            self.sim.set_base_pose("obstacle", self.obstacle_start[:3], self.obstacle_start[3:])
        elif self.kind == "Sta":                                        # reach.py:475-476
            self.sim.set_base_pose("obstacle", self.obstacle[:3], self.obstacle[3:])
        self.collision = self.sim.check_collision()                     # reach.py:322,477,679
        self.link_dist = self.sim.get_link_distances()                  # reach.py:323,478,680
        self.last_dist = self.link_dist                                 # reach.py:324,479,681 (alias)

    # ---- injection hooks
    def set_goal(self, test_goal) -> None:                              # reach.py:202-204
        self.goal = np.array(test_goal, dtype=np.float64)
        self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])

    def set_goal_and_obstacle(self, test_data) -> None:
        t = np.array(test_data, dtype=np.float64)
        if self.kind == "Obs":                                          # reach.py:328-335
            self.goal, self.obstacle = t[:3], t[3:]
            self.sim.set_base_pose("target", self.goal, np.array([0.0, 0.0, 0.0]))
            self.sim.set_base_pose("obstacle", self.obstacle[:3], self.obstacle[3:])
        elif self.kind == "Sta":                                        # reach.py:483-503
            if len(t) == 12:
                self.goal, self.obstacle = t[:6], t[6:]
                self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
                self.sim.set_base_pose("obstacle", self.obstacle[:3], self.obstacle[3:])
            else:
                self.goal = t[:6]
                self.obstacle_start, self.obstacle, self.obstacle_end = t[6:12], t[6:12], t[12:]
                self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
                self.sim.set_base_pose("obstacle", self.obstacle_start[:3], self.obstacle_start[3:])
        else:                                                           # reach.py:702-713
            self.goal, self.obstacle_start, self.obstacle_end = t[:6], t[6:12], t[12:]
            self.sim.set_base_pose("target", self.goal[:3], self.goal[3:])
            self.sim.set_base_pose("obstacle", self.obstacle_start[:3], self.obstacle_start[3:])
        self.collision = self.sim.check_collision()
        self.link_dist = self.sim.get_link_distances()
        self.last_dist = self.link_dist

    # ---- obstacle motion
    def _twist(self, time_duration: float):
        lin = (self.obstacle_end[:3] - self.obstacle_start[:3]) / time_duration
        rot_end = self.sim.euler_to_quaternion(self.obstacle_end[3:])
        rot_start = self.sim.euler_to_quaternion(self.obstacle_start[3:])
        axis, angle = self.sim.get_axis_angle(self.sim.get_quaternion_difference(rot_start, rot_end))
        return lin, np.array(axis) * angle / time_duration

    def set_velocity(self) -> None:
        if self.kind == "Dyn":                                          # reach.py:728-753
            if self.step_num < 25:
                lin, ang = self._twist(2)
            else:
                lin, ang = np.zeros(3), np.zeros(3)
            self.sim.set_velocity("obstacle", lin, ang)
            self.velocity = np.concatenate((lin, ang))
            self.step_num += 1
        else:                                                           # reach.py:518-541 (Sta, injection only)
            if np.linalg.norm(self.obstacle_end[:3] - self.sim.get_base_position("obstacle"), axis=-1) > 0.05:
                lin, ang = self._twist(1)
            else:
                lin, ang = np.zeros(3), np.zeros(3)
            self.sim.set_velocity("obstacle", lin, ang)
            self.velocity = np.concatenate((lin, ang))

    # ---- success / collision / reward
    def is_success(self, achieved_goal, desired_goal) -> np.ndarray:
        ok = distance(achieved_goal, desired_goal) < self.distance_threshold
        if self.kind != "Obs":                                          # reach.py:212-215,543-546,755-758
            ok = ok & (angular_distance(achieved_goal, desired_goal) < self.ori_distance_threshold)
        return np.array(ok, dtype=np.bool_)                             # Obs: reach.py:348-350

    def check_collision(self) -> bool:                                  # reach.py:217-219,352-354,548-550,760-762
        self.collision = self.sim.check_collision()
        return self.collision

    def compute_reward(self, achieved_goal, desired_goal, info=None) -> np.ndarray:
        if self.kind == "Ori":                                          # reach.py:221-236 (shape (1,), quirk Q7)
            r = np.float64(0.0)
            r = r + np.where(self.is_success(achieved_goal, desired_goal), self.success_weight, 0)
            r = r + distance(achieved_goal, desired_goal) * self.distance_weight
            r = r + angular_distance(achieved_goal, desired_goal) * self.orientation_weight
            r = r + (self.collision_weight if self.collision else 0)
            return r
        if self.kind == "Obs":                                          # reach.py:356-374
            self.link_dist = self.sim.get_link_distances()
            change = self.link_dist - self.last_dist
            self.last_dist = self.link_dist
            r = np.float64(0.0)
            r = r + np.where(self.is_success(achieved_goal, desired_goal), self.success_weight, 0)
            r = r + (self.collision_weight if self.collision else 0)
            r = r + self.distance_weight * distance(achieved_goal, desired_goal)
            r = r + np.where(self.link_dist < 0.2, self.obs_distance_weight * change, 0).sum()
            return r
        # Sta / Dyn                                                      reach.py:552-573, 764-785
        if self.collision:
            return np.float64(self.collision_weight)
        if self.is_success(achieved_goal, desired_goal):
            return np.float64(self.success_weight)
        r = np.float64(0.0)
        r = r + self.distance_weight * distance(achieved_goal, desired_goal)
        r = r + self.orientation_weight * angular_distance(achieved_goal, desired_goal)
        self.link_dist = self.sim.get_link_distances()
        change = self.link_dist - self.last_dist
        self.last_dist = self.link_dist
        r = r + np.where(self.link_dist < 0.2, self.dist_change_weight * change, 0).sum()
        return r


# --------------------------------------------------------------------------------------------------------------
# T1: env junction (core.py:222-317) + gymnasium TimeLimit(100) from the registration (UR_gym/__init__.py:19-42)
# --------------------------------------------------------------------------------------------------------------
class OracleEnv:
    max_episode_steps = 100

    def __init__(self, env_id: str, geom: int = GEOM_HULL, stream: Optional[UniformStream] = None,
                 env_index: int = 0, first_event: Optional[int] = None, link_dist_mode: int = LD_OBSTACLE):
        assert env_id in TASKS, env_id
        self.spec_id = env_id
        self.sim = OracleSim(geom=geom, link_dist_mode=link_dist_mode)
        self.robot = UR5Ori(self.sim)
        self.task = ReachTask(self.sim, self.robot, env_id, stream or NumpyStream(0), env_index)
        self._elapsed_steps = 0
        self.reset(event=first_event)                                   # core.py:237

    def _get_obs(self) -> Dict[str, np.ndarray]:                        # core.py:252-261
        robot_obs = self.robot.get_obs().astype(np.float32)
        task_obs = self.task.get_obs().astype(np.float32)
        return {"observation": np.concatenate([robot_obs, task_obs]),
                "achieved_goal": self.task.get_achieved_goal().astype(np.float32),
                "desired_goal": self.task.get_goal().astype(np.float32)}

    def reset(self, seed=None, options=None, event: Optional[int] = None):   # core.py:263-273
        self._elapsed_steps = 0
        self.robot.reset()
        self.task.reset(event)
        obs = self._get_obs()
        return obs, {"is_success": self.task.is_success(obs["achieved_goal"], self.task.get_goal())}

    def step(self, action):                                             # core.py:303-317
        self.robot.set_action(action)
        if self.spec_id == "UR5DynReach-v1":
            self.task.set_velocity()
        elif self.spec_id == "UR5StaReach-v1" and not np.array_equal(self.task.obstacle_end, np.zeros(6)):
            self.task.set_velocity()
        self.sim.step()
        collision = self.task.check_collision()
        obs = self._get_obs()
        terminated = bool(self.task.is_success(obs["achieved_goal"], self.task.get_goal()) or collision)
        info = {"is_success": (not collision) if terminated else terminated}
        reward = float(np.asarray(self.task.compute_reward(obs["achieved_goal"], self.task.get_goal(), info)).reshape(-1)[0])
        self._elapsed_steps += 1                                        # gymnasium TimeLimit.step
        truncated = self._elapsed_steps >= self.max_episode_steps
        return obs, reward, terminated, truncated, info


def make(env_id: str, **kw) -> OracleEnv:
    return OracleEnv(env_id, **kw)
