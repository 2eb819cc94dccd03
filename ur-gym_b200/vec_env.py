"""Batched (vectorised) reach-task simulator on one GPU: the env side of train.py's SAC loop
(/root/reference train.py:39-60 sees the env through SB3's DummyVecEnv: auto-reset, terminal observation kept,
TimeLimit.truncated) for N environments at once.

All arrays are torch tensors on the simulator's device; the C library writes into them in place on the current
torch CUDA stream.  `desired_goal` and `achieved_goal` are views of the observation buffer (goal = columns
12:12+G, EE pose = columns 0:G), the same numbers RobotTaskEnv._get_obs returns (core.py:252-261)."""
import ctypes
from typing import Dict, Optional

import torch

from . import _native as nat


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class UR5VecEnv:
    """N independent `UR5{Ori,Obs,Sta,Dyn}Reach-v1` environments with global indices
    [env_index_offset, env_index_offset + num_envs).

    geometry: "capsule" (bounding capsules, closed-form distances: the throughput path) or "hull" (the reference's
    convex-hull link meshes and cylinder obstacle through GJK: the reference-geometry path).
    link_dist: what the `link_dist` observation columns and reward term measure -- "obstacle" (links vs the obstacle,
    pyb_setup.py:439-456 as shipped; default) or "workbench" (per link the smallest distance to obstacle, table and
    track: the method's docstring, and the definition the shipped Obs / Sta policies were trained with)."""

    def __init__(self, env_id: str, num_envs: int, device: int = 0, seed: int = 0, env_index_offset: int = 0,
                 geometry: str = "capsule", auto_reset: bool = True, goal_buffers: bool = False,
                 link_dist: str = "obstacle"):
        if env_id not in nat.TASK_IDS:
            raise ValueError(f"unknown env id {env_id!r}; known: {sorted(nat.TASK_IDS)}")
        if geometry not in nat.GEOMS:
            raise ValueError("geometry must be 'capsule' or 'hull'")
        if link_dist not in nat.LINK_DIST_MODES:
            raise ValueError("link_dist must be 'obstacle' or 'workbench'")
        if not torch.cuda.is_available():
            raise nat.UrgymError("no CUDA device: ur-gym_b200 has no CPU path")
        self.env_id, self.task = env_id, nat.TASK_IDS[env_id]
        self.num_envs, self.device_index, self.seed, self.offset = int(num_envs), int(device), int(seed), int(env_index_offset)
        self.device = torch.device("cuda", self.device_index)
        self.geometry = geometry
        self.L = nat.lib()
        self.obs_dim, self.goal_dim = self.L.urgym_obs_dim(self.task), self.L.urgym_goal_dim(self.task)
        h = ctypes.c_void_p()
        rc = self.L.urgym_create(ctypes.byref(h), self.task, nat.GEOMS[geometry], self.num_envs, self.offset,
                                 ctypes.c_uint64(self.seed & (2 ** 64 - 1)), self.device_index)
        nat.check(None, rc)
        self.h = h
        if not auto_reset:
            nat.check(self.h, self.L.urgym_set_autoreset(self.h, 0))
        self.link_dist_mode = link_dist
        if link_dist != "obstacle" and self.task != 0:
            nat.check(self.h, self.L.urgym_set_link_dist_mode(self.h, nat.LINK_DIST_MODES[link_dist]))
        n, D, G = self.num_envs, self.obs_dim, self.goal_dim
        kw = dict(device=self.device)
        self.obs = torch.zeros((n, D), dtype=torch.float32, **kw)
        self.terminal_obs = torch.zeros((n, D), dtype=torch.float32, **kw)
        self.reward = torch.zeros(n, dtype=torch.float32, **kw)
        self.terminated = torch.zeros(n, dtype=torch.uint8, **kw)
        self.truncated = torch.zeros(n, dtype=torch.uint8, **kw)
        self.is_success = torch.zeros(n, dtype=torch.uint8, **kw)
        # separate achieved_goal / terminal achieved_goal arrays (the reference returns them as arrays of their own);
        # off by default because they are views of the observation buffers here
        self.achieved = torch.zeros((n, G), dtype=torch.float32, **kw) if goal_buffers else None
        self.terminal_achieved = torch.zeros((n, G), dtype=torch.float32, **kw) if goal_buffers else None
        self._step_ptrs = [_ptr(self.obs), _ptr(self.achieved), None, _ptr(self.reward), _ptr(self.terminated),
                           _ptr(self.truncated), _ptr(self.is_success), _ptr(self.terminal_obs), _ptr(self.terminal_achieved)]
        self._obs = self._obs_dict(self.obs)
        self._info = {"is_success": self.is_success, "terminal_observation": self.terminal_obs}
        # RobotTaskEnv.__init__ ends with reset() (core.py:237); a batch constructor that drew 1 Mi episodes nobody asked
        # for would be wasteful, so the first reset stays explicit -- but nothing may run on the zero-filled pool
        self._initialised = False

    # ---- plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.L.urgym_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _obs_dict(self, obs: torch.Tensor) -> Dict[str, torch.Tensor]:
        G = self.goal_dim
        return {"observation": obs, "achieved_goal": obs[:, :G], "desired_goal": obs[:, 12:12 + G]}

    # ---- gym-style batch API
    def reset(self, mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """reset all envs (mask None) or those with a non-zero mask byte; returns the observation dict"""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if mask is None:
            self._initialised = True
        self._need_reset()
        nat.check(self.h, self.L.urgym_reset(self.h, _ptr(mask), _ptr(self.obs), _ptr(self.achieved), None, self._stream()))
        return self._obs

    def _need_reset(self):
        if not self._initialised:
            raise nat.UrgymError("reset() (all envs) or load_state_dict() must be called before step / observe / "
                                 "reset(mask): the state pool is zero-filled (q = 0, no goal)")

    def step(self, actions: torch.Tensor):
        """actions float32 [N,6] on the device.  Returns (obs dict, reward, terminated, truncated, info); for envs that
        finished, obs is the first observation of the next episode and info["terminal_observation"] rows hold the last
        one (DummyVecEnv semantics; `time_limit_truncated()` gives its "TimeLimit.truncated").  Two kernel launches
        on the current torch stream (step, then the dense auto-reset of the finished envs), no synchronisation."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.num_envs, 6):
            raise ValueError(f"actions must have shape ({self.num_envs}, 6)")
        if not self._initialised:
            self._need_reset()
        rc = self.L.urgym_step(self.h, actions.data_ptr(), *self._step_ptrs,
                               torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            nat.check(self.h, rc)
        return self._obs, self.reward, self.terminated, self.truncated, self._info

    def time_limit_truncated(self) -> torch.Tensor:
        """DummyVecEnv's info["TimeLimit.truncated"]: truncated and not terminated"""
        return self.truncated & (1 - self.terminated)

    def capture_steps(self, action_buffers, chains: int = 1) -> "torch.cuda.CUDAGraph":
        """Capture one step per tensor of `action_buffers` (persistent device tensors the caller refills between
        replays) into a CUDA graph; graph.replay() then advances len(action_buffers) env steps with one launch from the
        host.  The reset-event counters live on the device, so replays draw fresh episodes.

        chains > 1 splits the envs into that many contiguous ranges, each advanced by its own branch of the graph
        (urgym_step_range): the branches are independent, so one range's auto-reset kernel overlaps another range's
        step kernel.  Results are identical to chains = 1."""
        self._need_reset()
        for a in action_buffers:
            if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous() or a.shape != (self.num_envs, 6):
                raise ValueError("capture_steps needs contiguous float32 [N,6] tensors on the simulator's device")
        if not 1 <= chains <= 8:
            raise ValueError("chains must be 1..8")
        per = -(-self.num_envs // chains)
        per = -(-per // 256) * 256                      # whole reset groups per chain
        ranges = [(f, min(per, self.num_envs - f)) for f in range(0, self.num_envs, per)]
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            main = torch.cuda.current_stream(self.device)
            if len(ranges) == 1:
                for a in action_buffers:
                    self.step(a)
            else:
                # chains count their own reset events: start from a common position whatever was stepped before
                nat.check(self.h, self.L.urgym_sync_events(self.h, main.cuda_stream))
                side = [torch.cuda.Stream(self.device) for _ in ranges[1:]]
                fork = torch.cuda.Event()
                fork.record(main)
                for c, (first, count) in enumerate(ranges):
                    st = main if c == 0 else side[c - 1]
                    if c:
                        st.wait_event(fork)
                    for a in action_buffers:
                        rc = self.L.urgym_step_range(self.h, first, count, c, a.data_ptr(), *self._step_ptrs, st.cuda_stream)
                        if rc != 0:
                            nat.check(self.h, rc)
                for st in side:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    main.wait_event(ev)
        return graph

    def reseed(self, seed: int) -> None:
        """re-key the counter-based reset stream (RobotTaskEnv.reset(seed=...), core.py:263-267)"""
        self.seed = int(seed)
        nat.check(self.h, self.L.urgym_set_seed(self.h, ctypes.c_uint64(self.seed & (2 ** 64 - 1))))

    def observe(self) -> Dict[str, torch.Tensor]:
        self._need_reset()
        with torch.cuda.device(self.device):
            nat.check(self.h, self.L.urgym_observe(self.h, _ptr(self.obs), None, None, self._stream()))
        return self._obs_dict(self.obs)

    # ---- state access: injection hooks of the reference + checkpointing
    _FIELDS = {"q": (nat.F_Q, 6, torch.float32), "goal": (nat.F_GOAL, None, torch.float32),
               "obstacle": (nat.F_OBSTACLE, 6, torch.float32), "obstacle_end": (nat.F_OBSTACLE_END, 6, torch.float32),
               "obstacle_start": (nat.F_OBSTACLE_START, 6, torch.float32),
               "link_dist": (nat.F_LINK_DIST, 5, torch.float32), "elapsed": (nat.F_ELAPSED, 0, torch.int32),
               "ep_return": (nat.F_EP_RETURN, 0, torch.float32), "velocity": (nat.F_VELOCITY, 6, torch.float32),
               "hot": (nat.F_HOT, 24, torch.float32)}

    def _field(self, name):
        fid, k, dt = self._FIELDS[name]
        if k is None:
            k = self.goal_dim
        shape = (self.num_envs,) if k == 0 else (self.num_envs, k)
        return fid, shape, dt

    def get_state(self, name: str) -> torch.Tensor:
        fid, shape, dt = self._field(name)
        out = torch.empty(shape, dtype=dt, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.h, self.L.urgym_get_state(self.h, fid, _ptr(out), self._stream()))
        return out

    def set_state(self, name: str, value) -> None:
        fid, shape, dt = self._field(name)
        v = torch.as_tensor(value, dtype=dt, device=self.device).expand(shape).contiguous()
        with torch.cuda.device(self.device):
            nat.check(self.h, self.L.urgym_set_state(self.h, fid, _ptr(v), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()     # v may be a temporary

    def refresh(self) -> torch.Tensor:
        """after injecting goal / obstacle / joints: recompute link_dist (= last_dist) and return the collision flags
        (tail of set_goal_and_obstacle, reach.py:333-335,501-503,711-713)"""
        coll = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.h, self.L.urgym_refresh(self.h, _ptr(coll), self._stream()))
        return coll

    def _state_names(self):
        # restore order: episode constants first (setting one re-derives the hot planes), then the per-step state, then
        # the raw hot words -- restored last they make a restore bit-exact whatever the dict's own order is
        names = ["goal"]
        if self.task != 0:
            names += ["obstacle"]
        if self.task in (2, 3):
            names += ["obstacle_end", "obstacle_start"] if self.task == 2 else ["obstacle_end"]
        names += ["q", "elapsed", "ep_return"]
        if self.task != 0:
            names += ["link_dist"]
        if self.task == 3:
            names += ["velocity"]
        return names + ["hot"]

    def state_dict(self) -> Dict[str, object]:
        ev = ctypes.c_uint32()
        nat.check(self.h, self.L.urgym_get_event(self.h, ctypes.byref(ev)))
        d = {k: self.get_state(k) for k in self._state_names()}
        d["event"] = int(ev.value)
        d["meta"] = {"env_id": self.env_id, "num_envs": self.num_envs, "seed": self.seed, "env_index_offset": self.offset,
                     "geometry": self.geometry, "link_dist": self.link_dist_mode}
        return d

    def load_state_dict(self, d: Dict[str, object], strict: bool = True) -> None:
        """restore a state_dict().  strict: the dict must come from a simulator with the same task, size, seed, env index
        offset, geometry and link-distance mode (the reset stream is keyed by seed and global env index: restoring into
        another configuration would silently continue with different episodes)."""
        meta = d.get("meta")
        if strict:
            mine = {"env_id": self.env_id, "num_envs": self.num_envs, "seed": self.seed, "env_index_offset": self.offset,
                    "geometry": self.geometry, "link_dist": self.link_dist_mode}
            if meta != mine:
                raise nat.UrgymError(f"state_dict was taken from {meta}, this simulator is {mine}")
        missing = [k for k in self._state_names() + ["event"] if k not in d]
        if missing:
            raise nat.UrgymError(f"state_dict lacks {missing}")
        for k in self._state_names():
            self.set_state(k, d[k])
        nat.check(self.h, self.L.urgym_set_event(self.h, int(d["event"])))
        self._initialised = True

    def stats(self, reset: bool = True) -> Dict[str, float]:
        """per-shard episode statistics accumulated on the device since the last reset=True call"""
        out = (ctypes.c_double * nat.STATS_COUNT)()
        with torch.cuda.device(self.device):
            nat.check(self.h, self.L.urgym_stats(self.h, out, int(reset), self._stream()))
        return dict(zip(nat.STAT_NAMES, list(out)))

    @property
    def launch_count(self) -> int:
        return int(self.L.urgym_launch_count(self.h))

    # ---- host-buffer path (what a CPU-side RL library uses): numpy / pinned tensors in, numpy out
    def step_host(self, actions, out: Dict[str, torch.Tensor]):
        """actions and the tensors in `out` (keys obs, reward, terminated, truncated, is_success and optionally
        terminal_obs) are HOST tensors, ideally pinned; copies in, steps, copies out, synchronises."""
        self._need_reset()
        nat.check(self.h, self.L.urgym_step_host(self.h, _ptr(actions), _ptr(out["obs"]), None, None, _ptr(out["reward"]),
                                                 _ptr(out["terminated"]), _ptr(out["truncated"]), _ptr(out["is_success"]),
                                                 _ptr(out.get("terminal_obs")), None))
        return out

    def step_host_async(self, slot: int, actions, out: Dict[str, torch.Tensor]) -> None:
        """enqueue a host-buffer step into output slot 0 / 1 and return at once (urgym_step_host_async); `out` and
        `actions` belong to the slot until host_wait(slot) returns"""
        self._need_reset()
        nat.check(self.h, self.L.urgym_step_host_async(self.h, int(slot), _ptr(actions), _ptr(out["obs"]), None, None,
                                                       _ptr(out["reward"]), _ptr(out["terminated"]), _ptr(out["truncated"]),
                                                       _ptr(out["is_success"]), _ptr(out.get("terminal_obs")), None))

    def host_wait(self, slot: int) -> None:
        nat.check(self.h, self.L.urgym_host_wait(self.h, int(slot)))

    def alloc_host_buffers(self, terminal_obs: bool = True, pin: bool = True) -> Dict[str, torch.Tensor]:
        n, D = self.num_envs, self.obs_dim
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=pin)
        out = {"actions": mk((n, 6), torch.float32), "obs": mk((n, D), torch.float32), "reward": mk((n,), torch.float32),
               "terminated": mk((n,), torch.uint8), "truncated": mk((n,), torch.uint8), "is_success": mk((n,), torch.uint8)}
        if terminal_obs:
            out["terminal_obs"] = mk((n, D), torch.float32)
        return out
