"""Where does the end-to-end step time go?  Every variant replays the SAME 150 steps (state restored).  (GPU box)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench
from pypownet_b200 import _lib
from pypownet_b200.vec_env import VecRunEnv, _ptr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
case, cfg, chronics, imaps = bench.build_workload('case14')
sc, sr = bench.env_starts(B)
env = VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps, start_chronics=sc, start_rows=sr)
act = torch.zeros((B, case.action_length), dtype=torch.uint8).pin_memory()
dact = act.cuda()
for _ in range(50): env.step(dact, auto_reset=True)
torch.cuda.synchronize()
saved = [env.get_state(f).clone() for f in (_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS)]
def restore():
    for f, v in zip((_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS), saved): env.set_state(f, v)
    torch.cuda.synchronize()
def timeit(name, f, n=150):
    restore()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); print('%-46s %.1f us/step' % (name, 1e6 * (time.perf_counter() - t0) / n))
def sync_step():
    env.step(dact, auto_reset=True); torch.cuda.synchronize()
def sync_step_noobs():
    env.step(dact, auto_reset=True, want_obs=False); torch.cuda.synchronize()
env.step_pinned(act)
po, pr, pd, pf = env._pin
def host_noobs():
    env._check(env.lib.ppn_step_host(env.handle, _ptr(act), None, env.obs_dynamic_length, _ptr(pr), _ptr(pd), _ptr(pf), None, 1))
for rep in range(1):
    timeit('step (device ptrs, async loop)', lambda: env.step(dact, auto_reset=True))
    timeit('step (device ptrs, sync each)', sync_step)
    timeit('step no obs (device ptrs, sync each)', sync_step_noobs)
    timeit('step_pinned (zero-copy)', lambda: env.step_pinned(act))
    timeit('step_pinned no actions', lambda: env.step_pinned(None))
    timeit('ppn_step_host without observation', host_noobs)
    os.environ['PPN_HOST_STAGED'] = '1'
    timeit('step_pinned (staged, 2 chunks)', lambda: env.step_pinned(act))
    del os.environ['PPN_HOST_STAGED']
restore()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(151)]
ev[0].record()
for k in range(150):
    env.step(dact, auto_reset=True); ev[k + 1].record()
torch.cuda.synchronize()
t = np.array([ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(150)])
print('device time per step back to back: mean %.1f us min %.1f max %.1f' % (t.mean(), t.min(), t.max()))
print(' '.join('%.0f' % x for x in t[:60]))
