"""CPU tier for the host layer: the C-ABI library loads and exports every symbol include/urgym_b200.h declares,
fails loudly without a GPU (no CPU fallback), the shard arithmetic, and the multi-rank statistics reduction over
gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "urgym_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(urgym_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import urgym_b200 as ug
    if not os.path.exists(ug.LIB_PATH):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "ur-gym_b200", "csrc")])
    L = ctypes.CDLL(ug.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/urgym_b200.h but not exported"
    assert sorted(ug._native.EXPORTS) == declared
    assert [L.urgym_obs_dim(t) for t in range(4)] == [18, 26, 29, 35]
    assert [L.urgym_goal_dim(t) for t in range(4)] == [6, 3, 6, 6]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import urgym_b200 as ug
    with pytest.raises(ug.UrgymError):
        ug.UR5VecEnv("UR5OriReach-v1", 4)
    with pytest.raises(ug.UrgymError):
        ug.make("UR5OriReach-v1")
    L = ug._native.lib()
    h = ctypes.c_void_p()
    assert L.urgym_create(ctypes.byref(h), 0, 1, 4, 0, 0, 0) == -2          # URGYM_ENODEVICE
    assert b"no CPU path" in L.urgym_last_error(None)
    assert L.urgym_create(ctypes.byref(h), 0, 1, ctypes.c_int64(1 << 31), 0, 0, 0) == -1     # URGYM_EINVAL: 32-bit env indices per handle
    assert b"2^31" in L.urgym_last_error(None)
    # the motor-driven env has no CPU path either
    with pytest.raises(ug.UrgymError):
        ug.UR5MotorVecEnv(4)
    with pytest.raises(ug.UrgymError):
        ug.make("UR5IAIReach-v1")
    assert L.urgym_motor_create(ctypes.byref(h), 4, 0, ctypes.c_uint64(0), 0) == -2   # URGYM_ENODEVICE
    assert b"no CPU path" in L.urgym_motor_last_error(None)
    with pytest.raises(NotImplementedError, match="IndexError"):
        ug.make("UR5RegReach-v1")                      # the reference's env cannot step
    with pytest.raises(ValueError):
        ug.make("UR5NoSuchReach-v1")


def test_product_never_imports_the_oracle():
    """no import / include / dlopen of anything under oracle/ from the product (comments may name the calibration tool)"""
    pkg = os.path.join(ROOT, "ur-gym_b200")
    pat = re.compile(r"(^|\s)(from|import)\s+oracle\b|#\s*include\s+\"[^\"]*oracle|libur_oracle|libur_motor_oracle|oracle_env|import\s+motor_oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and "build" not in dp:
                assert not pat.search(open(os.path.join(dp, f)).read()), f
    for f in ("urgym_b200.py",):
        assert not pat.search(open(os.path.join(ROOT, f)).read()), f


def test_shard_range_partitions():
    import urgym_b200 as ug
    for total in (1, 7, 1 << 20, (1 << 23) + 5):
        for w in (1, 2, 3, 4, 8):
            if total < w:
                continue
            r = [ug.shard_range(total, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ug.shard_range(8, 2, 2)


def test_utils_match_reference_golden(golden):
    """the host-side mirrors of UR_gym/utils.py against outputs of the reference module itself"""
    import urgym_b200 as ug
    from importlib import import_module
    u = import_module("ur-gym_b200.utils")
    np.testing.assert_allclose(u.distance(golden["a"], golden["b"]), golden["dist_batched"], atol=1e-15)
    got, want = u.angular_distance(golden["a"], golden["b"]), golden["ang_batched"]
    np.testing.assert_allclose(np.cos(got / 2), np.cos(want / 2), atol=4e-16)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import urgym_b200 as ug
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lo, hi = ug.shard_range(1001, rank, world)
stats = dict(zip(ug.STAT_NAMES, [float(hi - lo), -2.5 * (rank + 1), 10.0 * (hi - lo), 1.0 + rank, 2.0, 3.0, 7.0 * (hi - lo), 4.0]))
tot = ug.allreduce_stats(stats, device=torch.device("cpu"))
assert tot["episodes"] == 1001.0 and tot["return_sum"] == -7.5 and tot["successes"] == 3.0 and tot["env_steps"] == 7007.0, tot
s = ug.summarize(tot)
assert abs(s["mean_length"] - 10.0) < 1e-12
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_stats_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()
