"""ctypes binding of liburgym_b200.so (C ABI: include/urgym_b200.h).

There is no CPU implementation behind this module: if the shared library is missing or the machine has no
sm_100 GPU, creating a simulator raises.  The library is built in-tree by `make -C ur-gym_b200/csrc` (or
`__graft_entry__.build()`)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liburgym_b200.so")
if os.environ.get("URGYM_B200_LIB"):       # A/B builds while tuning kernels (tools/kernel_time.py): never silent
    LIB_PATH = os.environ["URGYM_B200_LIB"]
    import sys as _sys
    print(f"[urgym_b200] WARNING: URGYM_B200_LIB overrides the library under test: {LIB_PATH}", file=_sys.stderr)

TASK_IDS = {"UR5OriReach-v1": 0, "UR5ObsReach-v1": 1, "UR5StaReach-v1": 2, "UR5DynReach-v1": 3}
GEOM_HULL, GEOM_CAPSULE = 0, 1
GEOMS = {"hull": GEOM_HULL, "capsule": GEOM_CAPSULE}
(F_Q, F_GOAL, F_OBSTACLE, F_OBSTACLE_END, F_LINK_DIST, F_ELAPSED, F_EP_RETURN, F_VELOCITY, F_HOT,
 F_OBSTACLE_START) = range(10)
LD_OBSTACLE, LD_WORKBENCH = 0, 1
LINK_DIST_MODES = {"obstacle": LD_OBSTACLE, "workbench": LD_WORKBENCH}
STATS_COUNT = 8
STAT_NAMES = ("episodes", "return_sum", "length_sum", "successes", "collisions", "truncations", "env_steps",
              "reset_iterations")

EXPORTS = ["urgym_step_range", "urgym_create", "urgym_destroy", "urgym_last_error", "urgym_obs_dim", "urgym_goal_dim", "urgym_num_envs",
           "urgym_step", "urgym_reset", "urgym_observe", "urgym_refresh", "urgym_get_state", "urgym_set_state",
           "urgym_stats", "urgym_step_host", "urgym_reset_host", "urgym_step_host_async", "urgym_host_wait", "urgym_replay_write", "urgym_set_autoreset", "urgym_get_event",
           "urgym_set_event", "urgym_set_seed", "urgym_set_link_dist_mode", "urgym_sync_events", "urgym_launch_count", "urgym_profile_enable", "urgym_profile_read",
           "urgym_motor_create", "urgym_motor_destroy", "urgym_motor_last_error", "urgym_motor_reset", "urgym_motor_step",
           "urgym_motor_get_state", "urgym_motor_set_state", "urgym_motor_stats", "urgym_motor_launch_count"]
MOTOR_F_Q, MOTOR_F_QD, MOTOR_F_GOAL, MOTOR_F_ELAPSED = range(4)
MOTOR_ENV_ID = "UR5IAIReach-v1"


class UrgymError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UrgymError(f"{LIB_PATH} is missing: build it with `make -C ur-gym_b200/csrc` "
                             "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, u64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
        L.urgym_create.argtypes = [ctypes.POINTER(vp), i32, i32, i64, i64, u64, i32]
        L.urgym_destroy.argtypes = [vp]
        L.urgym_last_error.argtypes = [vp]; L.urgym_last_error.restype = ctypes.c_char_p
        L.urgym_obs_dim.argtypes = [i32]; L.urgym_goal_dim.argtypes = [i32]
        L.urgym_num_envs.argtypes = [vp]; L.urgym_num_envs.restype = i64
        L.urgym_step.argtypes = [vp] + [vp] * 10 + [vp]
        L.urgym_reset.argtypes = [vp, vp, vp, vp, vp, vp]
        L.urgym_step_range.argtypes = [vp, i64, i64, i32] + [vp] * 10 + [vp]
        L.urgym_observe.argtypes = [vp, vp, vp, vp, vp]
        L.urgym_refresh.argtypes = [vp, vp, vp]
        L.urgym_get_state.argtypes = [vp, i32, vp, vp]
        L.urgym_set_state.argtypes = [vp, i32, vp, vp]
        L.urgym_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), i32, vp]
        L.urgym_step_host.argtypes = [vp] + [vp] * 10
        L.urgym_reset_host.argtypes = [vp, vp, vp, vp, vp]
        L.urgym_step_host_async.argtypes = [vp, i32] + [vp] * 10
        L.urgym_host_wait.argtypes = [vp, i32]
        L.urgym_replay_write.argtypes = [vp] + [vp] * 13 + [i64, vp, vp]
        L.urgym_set_autoreset.argtypes = [vp, i32]
        L.urgym_get_event.argtypes = [vp, ctypes.POINTER(u32)]
        L.urgym_set_event.argtypes = [vp, u32]
        L.urgym_set_seed.argtypes = [vp, u64]
        L.urgym_set_link_dist_mode.argtypes = [vp, i32]
        L.urgym_sync_events.argtypes = [vp, vp]
        L.urgym_launch_count.argtypes = [vp]; L.urgym_launch_count.restype = i64
        L.urgym_profile_enable.argtypes = [vp, i32]
        L.urgym_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
        try:
            L.urgym_motor_create.argtypes = [ctypes.POINTER(vp), i64, i64, u64, i32]
            L.urgym_motor_destroy.argtypes = [vp]
            L.urgym_motor_last_error.argtypes = [vp]; L.urgym_motor_last_error.restype = ctypes.c_char_p
            L.urgym_motor_reset.argtypes = [vp, vp, vp, vp, vp, vp]
            L.urgym_motor_step.argtypes = [vp] + [vp] * 9 + [vp]
            L.urgym_motor_get_state.argtypes = [vp, i32, vp, vp]
            L.urgym_motor_set_state.argtypes = [vp, i32, vp, vp]
            L.urgym_motor_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_double), i32]
            L.urgym_motor_launch_count.argtypes = [vp]; L.urgym_motor_launch_count.restype = i64
        except AttributeError:
            if not os.environ.get("URGYM_B200_LIB"):      # only an A/B build of an older revision may lack the motor entry points
                raise
        _lib = L
    return _lib


def check_motor(handle, rc):
    if rc != 0:
        msg = lib().urgym_motor_last_error(handle)
        raise UrgymError(f"urgym motor error {rc}: {msg.decode() if msg else '?'}")


def check(handle, rc):
    if rc != 0:
        msg = lib().urgym_last_error(handle)
        raise UrgymError(f"urgym error {rc}: {msg.decode() if msg else '?'}")
