"""CPU oracle of the reference's motor-driven env `UR5IAIReach-v1` (UR_gym/envs/ur_tasks.py:10-21: robot `UR5`, task
`ReachIAI`).  TEST INFRASTRUCTURE ONLY -- the product never imports this module.

Restated line by line from the reference's Python (tier T1):
  RobotTaskEnv.step / reset / _get_obs   UR_gym/envs/core.py:252-273,303-317  (+ TimeLimit(100), UR_gym/__init__.py:7-11)
  UR5.set_action / get_obs / reset       UR_gym/envs/robots/UR5.py:44-50,76-103
  ReachIAI                               UR_gym/envs/tasks/reach.py:9-66
and, for what the reference asks of Bullet (tier T2, PARITY UNPINNED -- see the header of ur_motor_oracle.c),
  oracle/ur_motor_oracle.c               setJointMotorControlArray(POSITION_CONTROL) + 20 x stepSimulation.
Reset randomness: the same counter-based Philox stream as the four teleporting tasks (oracle_env.PhiloxStream)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libur_motor_oracle.so")
_lib = None

N_SUBSTEPS, DT = 20, 1.0 / 500            # pyb_setup.py:23-25,50
KP, KD, SOLVER_ITERS = 0.1, 1.0, 50       # pybullet's POSITION_CONTROL defaults, numSolverIterations  [RECALLED]
GRAVITY_Z, LINK_DAMPING = -9.81, 0.04     # pyb_setup.py:44; pybullet's default linear / angular damping  [RECALLED]
JOINT_FORCES = np.array([150.0, 150.0, 150.0, 28.0, 28.0, 28.0])      # UR5.py:36
NEUTRAL = np.array([0.0, -1.5708, 0.0, 0.0, 0.0, 0.0])                # UR5.py:39
GOAL_LOW, GOAL_HIGH = np.array([0.2, -0.4, 0.0]), np.array([0.6, 0.4, 0.8])      # reach.py:20-21 (goal_range = 0.8)
DISTANCE_THRESHOLD = 0.005                                            # reach.py:14
MAX_EPISODE_STEPS = 100                                               # UR_gym/__init__.py:10


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orm_mass_matrix.argtypes = [dp, dp]
        L.orm_bias.argtypes = [dp, dp, ctypes.c_double, ctypes.c_double, dp]
        L.orm_substeps.argtypes = [dp, dp, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                   ctypes.c_double, ctypes.c_double]
        L.orm_ee_state.argtypes = [dp, dp, dp, dp]
        L.orm_energy.argtypes = [dp, dp, ctypes.c_double]
        L.orm_energy.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def mass_matrix(q):
    q, M = np.ascontiguousarray(q, np.float64), np.zeros(36)
    lib().orm_mass_matrix(_p(q), _p(M))
    return M.reshape(6, 6)


def bias(q, qd, gravity_z=GRAVITY_Z, damping=LINK_DAMPING):
    q, qd, t = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64), np.zeros(6)
    lib().orm_bias(_p(q), _p(qd), gravity_z, damping, _p(t))
    return t


def substeps(q, qd, target, force=JOINT_FORCES, n_sub=N_SUBSTEPS, dt=DT, kp=KP, kd=KD, iters=SOLVER_ITERS, gravity_z=GRAVITY_Z,
             damping=LINK_DAMPING):
    q, qd = np.array(q, np.float64), np.array(qd, np.float64)
    target = np.ascontiguousarray(target, np.float64)
    fp = _p(np.ascontiguousarray(force, np.float64)) if force is not None else None
    lib().orm_substeps(_p(q), _p(qd), _p(target), fp, n_sub, dt, kp, kd, iters, gravity_z, damping)
    return q, qd


def ee_state(q, qd):
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    pos, vel = np.zeros(3), np.zeros(3)
    lib().orm_ee_state(_p(q), _p(qd), _p(pos), _p(vel))
    return pos, vel


def energy(q, qd, gravity_z=GRAVITY_Z):
    q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
    return lib().orm_energy(_p(q), _p(qd), gravity_z)


class UR5IAIReachOracle:
    """One env.  `stream`: an oracle_env.UniformStream (PhiloxStream for parity with the kernels)."""
    OBS, GOAL = 6, 3

    def __init__(self, stream, env_index=0, first_event=1):
        self.stream, self.env_index = stream, env_index
        self.q, self.qd = NEUTRAL.copy(), np.zeros(6)
        self.goal = np.zeros(3)
        self.elapsed = 0
        self.reset(event=first_event)

    # core.py:252-261
    def _get_obs(self):
        pos, vel = ee_state(self.q, self.qd)
        return {"observation": np.concatenate([pos, vel]).astype(np.float32),        # UR5.get_obs, UR5.py:92-97
                "achieved_goal": pos.astype(np.float32),                               # reach.py:44-46
                "desired_goal": self.goal.astype(np.float32)}

    # core.py:263-273: robot.reset (set_joint_neutral -> resetJointState: angles, zero velocities), task.reset
    def reset(self, event):
        self.q, self.qd = NEUTRAL.copy(), np.zeros(6)
        self.stream.begin(self.env_index, event, 1)                      # one Philox block per iteration: slots 0-2 = goal
        self.stream.iteration(0)
        u = np.array([self.stream.u(0), self.stream.u(1), self.stream.u(2)])
        self.goal = GOAL_LOW + (GOAL_HIGH - GOAL_LOW) * u                # np_random.uniform(low, high), reach.py:54-57
        self.elapsed = 0
        return self._get_obs()

    # core.py:303-317 + TimeLimit
    def step(self, action):
        a = np.clip(np.asarray(action, np.float32), -1.0, 1.0)          # UR5.py:45-46
        ctrl = (a * np.float32(np.pi)) * np.float32(0.1)                # UR5.py:48,85  (float32 like the numpy expression)
        target = self.q + ctrl.astype(np.float64)                       # UR5.py:87-89: current angles + control
        self.q, self.qd = substeps(self.q, self.qd, target)             # control_joints + sim.step()
        obs = self._get_obs()
        diff = obs["achieved_goal"] - obs["desired_goal"]                # utils.py:5-31 on the float32 arrays of _get_obs
        d = np.sqrt((diff ** 2).sum(-1))
        success = bool(d < DISTANCE_THRESHOLD)                           # reach.py:59-61
        terminated = bool(success)                                       # check_collision() returns None, core.py:310-313
        reward = float(np.float32(-d))                                   # reach.py:66-68
        self.elapsed += 1
        truncated = self.elapsed >= MAX_EPISODE_STEPS
        return obs, reward, terminated, truncated, {"is_success": terminated}
