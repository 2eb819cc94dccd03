"""Host-side mirrors of UR_gym/utils.py (distance, angular_distance, Euler samplers) for the Python API surface:
`task.is_success(achieved, desired)` and `env.compute_reward(...)` accept arbitrary arrays (SB3's HER relabelling
calls them on replay batches), so they are plain numpy functions here as they are in the reference.  The hot path
never calls these -- the step kernel evaluates the same formulas per env on the GPU."""
import numpy as np


def distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """UR_gym/utils.py:5-31: L2 distance of the first three components, shape (n,)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    d = a[..., :3] - b[..., :3]
    return np.sqrt((d * d).sum(-1)).reshape(-1)


def _quat_ZYX(e: np.ndarray) -> np.ndarray:
    # scipy Rotation.from_euler('ZYX', e).as_quat(): R = Rz(e0) Ry(e1) Rx(e2), (x, y, z, w)   UR_gym/utils.py:48-54
    e = np.asarray(e, dtype=np.float64)
    hz, hy, hx = e[..., 0] / 2, e[..., 1] / 2, e[..., 2] / 2
    cz, sz, cy, sy, cx, sx = np.cos(hz), np.sin(hz), np.cos(hy), np.sin(hy), np.cos(hx), np.sin(hx)
    return np.stack([cz * cy * sx - sz * sy * cx, cz * sy * cx + sz * cy * sx,
                     sz * cy * cx - cz * sy * sx, cz * cy * cx + sz * sy * sx], -1)


def angular_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """UR_gym/utils.py:34-69: 2*arccos(|<qa, qb>|) of the Euler parts (columns 3:6), with the reference's 'ZYX' reading
    of the (roll, pitch, yaw) triple."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    dot = np.sum(_quat_ZYX(a[..., 3:]) * _quat_ZYX(b[..., 3:]), axis=-1)
    return (2 * np.arccos(np.abs(np.clip(dot, -1.0, 1.0)))).reshape(-1)


def sample_euler_constrained() -> np.ndarray:
    """UR_gym/utils.py:81-86 (global numpy generator, like the reference)."""
    return np.deg2rad([np.random.uniform(-90, -180), 0, np.random.uniform(0, -180)])


def sample_euler_obstacle() -> np.ndarray:
    """UR_gym/utils.py:88-101."""
    if np.random.choice(["negative", "positive"], p=[0.5, 0.5]) == "negative":
        roll = np.random.uniform(-30, -150)
    else:
        roll = np.random.uniform(30, 150)
    pitch = np.random.uniform(-30, -150) if (roll < -90 or roll > 90) else np.random.uniform(30, 150)
    return np.deg2rad([roll, pitch, 0])
