"""ur-gym_b200: B200-native batched simulator for UR-gym's reach tasks (UR5{Ori,Obs,Sta,Dyn}Reach-v1).

    from urgym_b200 import make, UR5VecEnv            # `urgym_b200.py` at the repo root aliases this package,
    env = make("UR5DynReach-v1", render=False)        # whose directory name (ur-gym_b200) is not an identifier
    vec = UR5VecEnv("UR5DynReach-v1", num_envs=1 << 20, device=0)
"""
from ._native import GEOM_CAPSULE, GEOM_HULL, LIB_PATH, STAT_NAMES, TASK_IDS, UrgymError  # noqa: F401
from .envs import ENV_IDS, RobotTaskEnv, make, register_with_gymnasium  # noqa: F401
from .sharding import allreduce_stats, shard_range, summarize  # noqa: F401
from .sb3_vec_env import SB3VecEnvAdapter  # noqa: F401
from .vec_env import UR5VecEnv  # noqa: F401
from .rollout import DeviceReplayRing, Rollout, mlp_policy  # noqa: F401
from .motor_env import MotorTaskEnv, UR5MotorVecEnv  # noqa: F401

register_with_gymnasium()
