#!/usr/bin/env python3
"""Generate tier-T0 golden vectors by EXECUTING the reference's own UR_gym/utils.py (numpy + scipy only).

Run in the build container (where /root/reference is mounted); the GPU box has no reference, so the outputs
are committed:   python tests/golden/make_golden.py  ->  tests/golden/utils_golden.npz

Pinned functions (reference file:line):
  distance                  UR_gym/utils.py:5-31
  angular_distance          UR_gym/utils.py:34-69
  sample_euler_constrained  UR_gym/utils.py:81-86
  sample_euler_obstacle     UR_gym/utils.py:88-101
For the two samplers the uniform doubles numpy's legacy global generator hands out under the same seed are
recorded next to the outputs, so a restatement written as a pure function of those uniforms can be checked.
"""
import importlib.util
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "UR_gym/utils.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(20261018)
N = 512
a = np.concatenate([rng.uniform(-1.2, 1.2, (N, 3)), rng.uniform(-np.pi, np.pi, (N, 3))], 1)
b = np.concatenate([rng.uniform(-1.2, 1.2, (N, 3)), rng.uniform(-np.pi, np.pi, (N, 3))], 1)
# edge cases: identical poses, antipodal quaternions, gimbal pitch, tiny differences around the success thresholds
a[0] = b[0]
a[1, 3:] = b[1, 3:] + [2 * np.pi, 0, 0]
a[2, 3:] = [0.3, np.pi / 2, -0.2]; b[2, 3:] = [0.1, np.pi / 2, 0.4]
a[3, 3:] = b[3, 3:] + [0.0873, 0, 0]
a[4, 3:] = b[4, 3:] + [0, 0, 1e-7]
a[5, :3] = b[5, :3] + [0.05, 0, 0]
a[6] = 0.0; b[6] = 0.0
a32 = a.astype(np.float32)      # the reference feeds float32 achieved goals and float64 goals (core.py:256,313)

out = dict(a=a, b=b,
           dist_batched=ref.distance(a, b), ang_batched=ref.angular_distance(a, b),
           dist_rows=np.array([ref.distance(a[i], b[i])[0] for i in range(N)]),
           ang_rows=np.array([ref.angular_distance(a[i], b[i])[0] for i in range(N)]),
           dist_f32=np.array([ref.distance(a32[i], b[i])[0] for i in range(N)]),
           ang_f32=np.array([ref.angular_distance(a32[i], b[i])[0] for i in range(N)]),
           dist3=np.array([ref.distance(a[i, :3], b[i, :3])[0] for i in range(N)]))

M = 4096
seed = 12345
np.random.seed(seed)
cons = np.array([ref.sample_euler_constrained() for _ in range(M)])
np.random.seed(seed)
cons_u = np.random.random_sample(2 * M).reshape(M, 2)          # the doubles the two uniform() calls consumed
np.random.seed(seed + 1)
obst = np.array([ref.sample_euler_obstacle() for _ in range(M)])
np.random.seed(seed + 1)
obst_u = np.random.random_sample(3 * M).reshape(M, 3)          # choice(), uniform(roll), uniform(pitch)
out.update(cons=cons, cons_u=cons_u, obst=obst, obst_u=obst_u)

np.savez_compressed(os.path.join(HERE, "utils_golden.npz"), **out)
print("wrote", os.path.join(HERE, "utils_golden.npz"), {k: v.shape for k, v in out.items()})
