"""ctypes wrapper over tests/hostcheck (the product's per-env __host__ __device__ functions compiled for the host)
plus a Python re-statement of the step kernel's orchestration (terminal observation, auto-reset, reset-event
counter), so the CPU test tier can drive the product's arithmetic the way the CUDA kernel does.  Test
infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HC_DIR = os.path.join(HERE, "hostcheck")
TASK_ID = {"UR5OriReach-v1": 0, "UR5ObsReach-v1": 1, "UR5StaReach-v1": 2, "UR5DynReach-v1": 3}
OBS_DIM = [18, 26, 29, 35]
GOAL_DIM = [6, 3, 6, 6]
_libs = {}


def lib(geom):
    if geom not in _libs:
        target = os.path.join(HC_DIR, f"libhostcheck_g{geom}.so")
        subprocess.check_call(["make", "-s", "-C", HC_DIR, target])
        _libs[geom] = ctypes.CDLL(target)
    return _libs[geom]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class HostCheckSim:
    def __init__(self, env_id, geom, n, seed=0, offset=0, autoreset=True, link_dist_mode=0):
        self.task, self.geom, self.n, self.seed, self.offset = TASK_ID[env_id], geom, n, seed, offset
        self.link_dist_mode = link_dist_mode
        self.D, self.G = OBS_DIM[self.task], GOAL_DIM[self.task]
        self.q = np.zeros((n, 6), np.float32); self.elapsed = np.zeros(n, np.int32); self.ep_ret = np.zeros(n, np.float32)
        self.ld = np.zeros((n, 5), np.float32); self.E = np.zeros((n, 24), np.float32)
        self.stale_vel = np.zeros((n, 6), np.float32)
        self.event = 0
        self.autoreset = autoreset
        self.L = lib(geom)
        self.L.hc_set_ld_mode(link_dist_mode)       # library-global: one HostCheckSim per mode at a time

    def _state_ptrs(self, idx=None):
        return [_p(self.q), _p(self.elapsed), _p(self.ep_ret), _p(self.ld), _p(self.E)]

    def _reset_rows(self, rows, obs):
        """reset envs `rows`; obs rows must already hold the velocity columns to keep (quirk Q4)"""
        m = len(rows)
        q = np.zeros((m, 6), np.float32); el = np.zeros(m, np.int32); er = np.zeros(m, np.float32)
        ld = np.zeros((m, 5), np.float32); E = np.ascontiguousarray(self.E[rows])
        o = np.ascontiguousarray(obs[rows]); it = np.zeros(m, np.int32)
        env_index = (np.asarray(rows, np.int64) + self.offset).astype(np.int64)
        rc = self.L.hc_reset(self.task, self.geom, ctypes.c_int64(m), _p(q), _p(el), _p(er), _p(ld), _p(E),
                             ctypes.c_uint64(self.seed), ctypes.c_uint32(self.event), _p(env_index), _p(o), _p(it))
        assert rc == 0
        self.q[rows], self.elapsed[rows], self.ep_ret[rows], self.ld[rows], self.E[rows] = q, el, er, ld, E
        obs[rows] = o
        return it

    def reset(self):
        self.event += 1
        obs = np.zeros((self.n, self.D), np.float32)
        if self.task == 3:
            obs[:, 24:30] = self.stale_vel
        self._reset_rows(np.arange(self.n), obs)
        return obs

    def step(self, actions):
        self.event += 1
        n, D = self.n, self.D
        act = np.ascontiguousarray(actions, np.float32)
        obs = np.zeros((n, D), np.float32); rew = np.zeros(n, np.float32); flags = np.zeros((n, 4), np.uint8)
        vel = np.zeros((n, 6), np.float32)
        rc = self.L.hc_step(self.task, self.geom, ctypes.c_int64(n), *self._state_ptrs(), _p(act), _p(obs), _p(rew),
                            _p(flags), _p(vel))
        assert rc == 0
        term, trunc, succ = flags[:, 0].astype(bool), flags[:, 1].astype(bool), flags[:, 2].astype(bool)
        done = term | trunc
        out = dict(reward=rew, terminated=term, truncated=trunc, is_success=succ, collision=flags[:, 3].astype(bool),
                   terminal_obs=obs.copy())
        rows = np.nonzero(done)[0]
        if self.autoreset and len(rows):
            if self.task == 3:
                self.stale_vel[rows] = obs[rows, 24:30]
            self._reset_rows(rows, obs)
        out["obs"] = obs
        return out

    def observe(self):
        obs = np.zeros((self.n, self.D), np.float32)
        self.L.hc_observe(self.task, ctypes.c_int64(self.n), *self._state_ptrs(), _p(self.stale_vel), _p(obs))
        return obs

    def refresh(self):
        coll = np.zeros(self.n, np.uint8)
        rc = self.L.hc_refresh(self.task, self.geom, ctypes.c_int64(self.n), *self._state_ptrs(), _p(coll))
        assert rc == 0
        return coll.astype(bool)

    # injection (API layout), mirroring urgym_set_state
    def set_q(self, q):
        self.q[:] = q

    def set_goal(self, g):
        self.E[:, :self.G] = g

    def set_obstacle(self, o):
        off = 3 if self.task == 1 else 6
        self.E[:, off:off + 6] = o

    def set_obstacle_end(self, o):
        self.E[:, 12:18] = o

    def set_obstacle_start(self, o):            # ReachSta only (Dyn's start is set_obstacle)
        self.E[:, 18:24] = o


def ee_pose(q):
    q = np.ascontiguousarray(q, np.float32)
    ee = np.zeros((len(q), 6), np.float32)
    lib(1).hc_ee_pose(ctypes.c_int64(len(q)), _p(q), _p(ee))
    return ee


def philox(counter, key):
    c = (ctypes.c_uint32 * 4)(*counter); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
    lib(1).hc_philox(c, k, o)
    return tuple(o)
