"""Debug build (-DPPN_TIMING) of the library + clock64() phase timings of one load-flow of env 0.
    python tools/phase_timing.py [grid] [envs]      (GPU box)"""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
lib_dbg = os.path.join(ROOT, 'pypownet_b200', 'libpypownet_b200_timing.so')   # built in the container; *.so travel with the snapshot
import __graft_entry__ as g
if not os.path.exists(lib_dbg) or '--build' in sys.argv:
    os.makedirs(os.path.dirname(lib_dbg), exist_ok=True)
    subprocess.run(['/usr/local/cuda/bin/nvcc'] + g.NVCC_FLAGS + ['-DPPN_TIMING', '-o', lib_dbg] + [os.path.join(g.CSRC, s) for s in g.SOURCES], check=True)
if '--build' in sys.argv:
    sys.exit(0)
from pypownet_b200 import _lib
_lib.LIB_PATH = lib_dbg
import bench
from pypownet_b200.vec_env import VecRunEnv
grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
case, cfg, chronics, imaps = bench.build_workload(grid)
sc, sr = bench.shard_starts(B, 0, 1)
env = VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps, start_chronics=sc, start_rows=sr)
lib = env.lib
lib.ppn_debug_timing.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
buf = (ctypes.c_longlong * 64)()
names = ['start', 'types+scan', 'connectivity', 'entries', 'sbus', 'V0', 'matrices built', 'first mismatch', "B' inverted (sparse: factor)", "B'' inverted (sparse: inverses)", 'iterations done', 'pfsoln', 'checks']
for step in range(8):
    torch.cuda.synchronize(); lib.ppn_debug_timing(buf, 1)
    env.step(None, auto_reset=True); torch.cuda.synchronize()
    lib.ppn_debug_timing(buf, 0)
    t = np.array(buf[:13], dtype=np.int64)
    if t[12] == 0 or t[10] == 0:
        print('step', step, 'first load-flow of env 0 ended early'); continue
    d = np.diff(t)
    halves = buf[30]
    print('step %d: n1=%d n2=%d half-iterations=%d total %d cycles' % (step, buf[31], buf[32], halves, t[12] - t[0]))
    print('   ' + ', '.join('%s %d' % (names[i + 1], d[i]) for i in range(12)))
    print('   in the loop: P updates %d, Q updates %d, mismatches %d  -> per half-iteration: update %.0f + mismatch %.0f cycles'
          % (buf[21], buf[20], buf[22], (buf[20] + buf[21]) / max(halves, 1), buf[22] / (halves + 1)))
    if buf[23]:
        print('   triangular solves %d cycles in total, %.0f per half-iteration' % (buf[23], buf[23] / max(halves, 1)))
