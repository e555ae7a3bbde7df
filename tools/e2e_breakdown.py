"""Where does an end-to-end step (ppn_step_host, pinned buffers, zero-copy results) spend its time?  IEEE-14 x 4096.
Variants: full call; no observation rows; do-nothing without the action upload; device-pointer ppn_step + one sync.
    python tools/e2e_breakdown.py      (GPU box)"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pypownet_b200.vec_env import VecRunEnv, _ptr  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
case, cfg, chronics, imaps = bench.build_workload(grid)
sc, sr = bench.shard_starts(B, 0, 1)


def fresh():
    return VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps,
                     start_chronics=sc, start_rows=sr)


def timeit(fn, n=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


nd = case.obs_dynamic_length
act = torch.zeros((B, case.action_length), dtype=torch.uint8).pin_memory()
env = fresh()
print('%s x %d' % (grid, B))
print('  step_pinned (actions up, float64 rows down)        %.4f ms' % timeit(lambda: env.step_pinned(act)))
print('  step_pinned (float32 rows)                         %.4f ms' % timeit(lambda: env.step_pinned(act, obs_dtype=torch.float32)))
env = fresh()
full, po, pr, pd, pf = torch.empty((B, (nd + 1) & ~1), dtype=torch.float64).pin_memory(), None, torch.empty((B, 5), dtype=torch.float64).pin_memory(), \
    torch.empty((B,), dtype=torch.uint8).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()
lib, h = env.lib, env.handle
print('  ppn_step_host, no observation rows                 %.4f ms' % timeit(
    lambda: lib.ppn_step_host(h, _ptr(act), None, 0, _ptr(pr), _ptr(pd), _ptr(pf), None, 1)))
env = fresh(); lib, h = env.lib, env.handle
print('  ppn_step_host, no rows, no action upload           %.4f ms' % timeit(
    lambda: lib.ppn_step_host(h, None, None, 0, _ptr(pr), _ptr(pd), _ptr(pf), None, 1)))
env = fresh(); lib, h = env.lib, env.handle
print('  ppn_step_host, rows, no action upload              %.4f ms' % timeit(
    lambda: lib.ppn_step_host(h, None, _ptr(full), full.shape[1], _ptr(pr), _ptr(pd), _ptr(pf), None, 1)))
env = fresh()


def dev_step():
    env.step(None, auto_reset=True)
    torch.cuda.synchronize()


print('  ppn_step (device buffers) + synchronize            %.4f ms' % timeit(dev_step))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env = fresh()
for _ in range(10):
    env.step(None, auto_reset=True)
a.record()
for _ in range(100):
    env.step(None, auto_reset=True)
b.record()
torch.cuda.synchronize()
print('  ppn_step back to back, device time                 %.4f ms' % (a.elapsed_time(b) / 100))
for stride in (nd, (nd + 1) & ~1, (nd + 15) & ~15, (nd + 31) & ~31, (nd + 63) & ~63, 512):
    env = fresh(); lib, h = env.lib, env.handle
    buf = torch.empty((B, stride), dtype=torch.float64).pin_memory()
    ms = timeit(lambda: lib.ppn_step_host(h, None, _ptr(buf), stride, _ptr(pr), _ptr(pd), _ptr(pf), None, 1))
    print('  rows with stride %4d doubles (%5d B, %s)            %.4f ms' % (stride, 8 * stride, 'x256 B' if (8 * stride) % 256 == 0 else ('x128 B' if (8 * stride) % 128 == 0 else 'unaligned'), ms))
