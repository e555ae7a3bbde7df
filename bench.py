"""Benchmark of the batched pypownet step path on B200 (BASELINE.json metric: env steps/s on batched grids).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--grid case14] [--envs 4096]

N=1 workload = BASELINE.json configs[1]: default14 AC, 4096 batched envs, do-nothing agent, one B200; envs start on
different chronics/rows, games that end are restarted in the same step (Runner semantics).  The reference's chronics
do not travel with the repo: chronics are synthetic with the shipped ones' statistics (pypownet_b200/synthetic.py).
A "step" is one env-step of every env of the batch (one fused kernel launch).  One JSON line on stdout (rank 0).
`--impl reference` times the UNMODIFIED reference package (installed in baseline/_ref by tools/install_reference.sh; it
travels with the snapshot) on all host cores for the same workload, on a bounded sample per step: one RunEnv per process
on oracle/shims (PYPOWER / gym restated).  Without baseline/_ref -- or with PPN_CPU_BASELINE=port -- it times the numpy
restatement of the same path (oracle/flat.py) and says kind "port".
Other switches: --cascade (synthetic thermal limits that trip), --agent nothing|random, --no-secondary (only the primary
workload), --no-cpu, --sharding spread|blocks|strided, --emulate-shard R/W, --profile-ranks FILE (per-rank table at N>1).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIDS = {'case14': 'default14', 'case30': 'default30', 'case118': 'default118'}
N_CHRONICS, N_ROWS = 12, 720


def algorithmic_bytes(case, with_obs=True):
    """bytes one env-step must move through HBM."""
    G, L, N, S = case.n_gen, case.n_load, case.n_line, case.n_sub
    chronic_in = 4 * (2 * G + 2 * L + 2 * N)
    planned_in = 4 * (2 * G + 2 * L)
    action_in = G + L + 3 * N
    state = 36 * S + 15 * N + G + L + 16
    dyn_obs = 8 * (7 * L + 7 * G + 13 * N + S + 6)
    return chronic_in + planned_in + action_in + 2 * state + (dyn_obs if with_obs else 0) + 48


def build_workload(grid, seed=0, cascade=False):
    """Grid, configuration, synthetic chronics and thermal limits of a bench workload.  cascade=True swaps the shipped
    limits (never exceeded on IEEE-30 / IEEE-118) for the synthetic ones of tools/make_cascade_limits.py
    (1.05 x p90 of each line's do-nothing flow), so that the cascading-failure loop fires."""
    from pypownet_b200.case import Case
    from pypownet_b200 import synthetic
    case = Case.builtin(grid)
    with open(os.path.join(ROOT, 'pypownet_b200', 'data', grid + '.json')) as f:
        d = json.load(f)
    imaps = np.array(d['imaps_cascade' if cascade and 'imaps_cascade' in d else 'imaps'], dtype=np.float64)
    chronics = synthetic.make_chronics(case, N_CHRONICS, N_ROWS, seed=seed, thermal_limits=imaps)
    return case, synthetic.default_config(grid), chronics, imaps


def env_starts(n_envs, offset=0):
    """env e plays chronic e mod 12 from row (e // 12) mod (T - 1); e is the GLOBAL index."""
    from pypownet_b200.sharding import env_starts as starts
    return starts(N_CHRONICS, N_ROWS, offset, offset + n_envs)


def random_action_bank(case, n_envs, n_batches=16, seed=1234):
    """RandomNodeSplitting + RandomLineSwitch (agent.py:78-158): per env and step one substation
    with its element bits i.i.d. Bernoulli(1/2) plus one line switch.  uint8 [n_batches, n_envs, action_length]; the
    bench cycles through the pre-drawn batches."""
    rng = np.random.default_rng(seed)
    elem_sub = np.asarray(case.elem_sub)
    nt_ = len(elem_sub)
    bank = np.zeros((n_batches, n_envs, case.action_length), dtype=np.uint8)
    for k in range(n_batches):
        subs = rng.integers(0, case.n_sub, size=n_envs)
        bits = rng.integers(0, 2, size=(n_envs, nt_), dtype=np.uint8)
        bank[k, :, :nt_] = bits * (elem_sub[None, :] == subs[:, None])
        bank[k, np.arange(n_envs), nt_ + rng.integers(0, case.n_line, size=n_envs)] = 1
    return bank


# ------------------------------------------------------------------------------------- CPU baseline (reference / port)
# kind "reference": the UNMODIFIED reference package installed in baseline/_ref (tools/install_reference.sh; it travels
# with the snapshot) on oracle/shims (PYPOWER 5.1.4 slice + gym.spaces restated: neither is installable), one RunEnv per
# process on the synthetic bench workload written in the reference's on-disk format.  kind "port": oracle/flat.py, the
# numpy restatement of the same path (5-10x faster per core than the reference), when baseline/_ref is absent.
_WORKER = {}
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, 'pypownet')) and os.environ.get('PPN_CPU_BASELINE', '') != 'port'


def _reference_folder(grid):
    """The synthetic workload as an environment folder of the reference, written once per (grid, machine)."""
    import tempfile
    from oracle.ref_folder import write_environment_folder
    root = os.path.join(tempfile.gettempdir(), 'pypownet_b200_bench_%s_%d' % (grid, os.getuid()))
    marker = os.path.join(root, '.complete')
    if not os.path.exists(marker):
        case, cfg, chronics, imaps = build_workload(grid)
        for ch in chronics:
            ch.imaps = np.asarray(imaps, dtype=np.float64)
        write_environment_folder(root, case, cfg, chronics)      # no reward_signal.py: the reference's own default
        open(marker, 'w').close()
    return root


def _cpu_init(grid, counter, folder):
    """Pool initializer: every process builds ONE env once (as a reference process would)."""
    sys.path.insert(0, ROOT)
    with counter.get_lock():
        env_id = counter.value
        counter.value += 1
    if folder is not None:                                # the unmodified reference
        import logging
        import warnings
        logging.disable(logging.CRITICAL)
        warnings.simplefilter('ignore')
        sys.path[:0] = [os.path.join(ROOT, 'oracle', 'shims'), REF_DIR]
        os.chdir(folder)                                  # the reference writes tmp/ and logs into the cwd
        from pypownet.environment import RunEnv
        env = RunEnv(folder, 'level0', start_id=env_id % N_CHRONICS)
        a = env.action_space.get_do_nothing_action()
        step = lambda: env.step(a)[2]                     # noqa: E731
        over = env.process_game_over
    else:
        from oracle.flat import FlatEnv, Config
        case, cfg, chronics, imaps = build_workload(grid)
        c, r = shard_starts(4096, 0, 1)                   # one env of the 4096-env GPU batch per process
        k = (257 * env_id) % 4096
        fenv = FlatEnv(case, Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub), chronics,
                       start_id=int(c[k]), thermal_limits=imaps, start_row=int(r[k]))
        a = np.zeros(case.action_length, dtype=np.uint8)
        step = lambda: fenv.step(a)[2]                    # noqa: E731
        over = fenv.process_game_over
    for _ in range(5):
        if step():
            over()
    _WORKER['step'], _WORKER['over'] = step, over


def _cpu_worker(args):
    n_steps, budget_s = args
    step, over = _WORKER['step'], _WORKER['over']
    t0 = time.perf_counter()
    done_steps = 0
    while done_steps < n_steps and (budget_s is None or time.perf_counter() - t0 < budget_s):
        if step():
            over()
        done_steps += 1
    return done_steps, time.perf_counter() - t0


def cpu_pool(grid, cores=None):
    """(pool, cores, kind).  The unmodified reference when baseline/_ref is installed, else the port."""
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    if reference_available():
        try:
            folder = _reference_folder(grid)
            pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(grid, ctx.Value('i', 0), folder))
            pool.map(_cpu_worker, [(1, None)] * cores)
            return pool, cores, 'reference'
        except Exception as exc:                          # e.g. a box without PyYAML: fall back to the port, loudly
            sys.stderr.write('reference CPU arm unavailable (%r), timing the port instead\n' % (exc,))
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(grid, ctx.Value('i', 0), None))
    pool.map(_cpu_worker, [(1, None)] * cores)          # make sure every process is up before anything is timed
    return pool, cores, 'port'


def cpu_baseline(pool, cores, steps_per_proc, budget_s=None):
    """All processes step their env concurrently; throughput = total env-steps / the slowest process' time."""
    res = pool.map(_cpu_worker, [(steps_per_proc, budget_s)] * cores, chunksize=1)
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, total, busy


# ------------------------------------------------------------------------------------------------------ clock sampler
class ClockSampler(object):
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 5.0:      # up and sampling before anything is timed
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------------- main
def run_reference(args):
    """CPU arm: the unmodified reference package (baseline/_ref) -- or, without it, oracle/flat.py -- on every host core,
    do-nothing agent, a bounded sample of the workload per bench step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pool, cores, kind = cpu_pool(args.grid)
    per_step = 20 if kind == 'reference' else 100        # env-steps per process per bench step (bounded sample)
    for _ in range(args.warmup):
        cpu_baseline(pool, cores, 2 if kind == 'reference' else 10)
    total = 0
    busy = 0.0
    for _ in range(args.steps):
        v, n, b = cpu_baseline(pool, cores, per_step)
        total += n
        busy += b
    pool.close()
    value = total / busy
    what = 'the UNMODIFIED reference package (baseline/_ref) on oracle/shims' if kind == 'reference' else \
        'oracle/flat.py, the numpy restatement of the reference path'
    sample = '%d processes (one env each) x %d do-nothing env-steps of %s per bench step, synthetic chronics; %s' % (
        cores, per_step, args.grid, what)
    cfg = workload_config(args, None)
    cfg['workload'] = 'CPU arm: %s AC, do-nothing agent, restart on game over; BOUNDED SAMPLE of the 4096-env workload: ' \
                      '%d single-env processes x %d env-steps per bench step (%s on the host cores)' % (
                          GRIDS[args.grid], cores, per_step, what)
    line = {'impl': 'reference', 'metric': 'env steps/sec (batched grids)', 'value': value, 'unit': 'env-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * busy / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': cfg,
            'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_name(grid, envs, agent, cascade):
    return '%s AC%s, %d batched envs per GPU, %s agent, auto-restart on game over' % (
        GRIDS[grid], ' with synthetic thermal limits (cascading-failure loop fires)' if cascade else '', envs,
        'do-nothing' if agent == 'nothing' else 'random node-split + line-switch')


def workload_config(args, extra):
    cfg = {'workload': workload_name(args.grid, args.envs, args.agent, args.cascade) +
           ' (BASELINE.json configs[1] shape)',
           'grid': args.grid, 'envs_per_gpu': args.envs, 'chronics': '%d synthetic x %d rows' % (N_CHRONICS, N_ROWS),
           'env_starts': 'local env k: chronic k mod 12, first row spread evenly over the chronic (+97 rows per rank)',
           'solver': 'fast-decoupled XB, tol 1e-6, <=25 it (the reference\'s PF_ALG=2)',
           'l2': 'flushed between timed steps (256 MiB write)', 'parallelism': 'env-sharded, dp%d' % args.gpus}
    if extra:
        cfg.update(extra)
    return cfg


def hbm_peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def traffic_of(grid, envs, agent):
    """DRAM bytes per launch of the step kernel from the committed ncu captures (profiles/traffic.json), or None."""
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(tpath):
        return None
    with open(tpath) as f:
        t = json.load(f)
    return t.get('%s_%d%s' % (grid, envs, '' if agent == 'nothing' else '_' + agent))


class Ctx(object):
    """torch / torch.distributed handles of this process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_list(self, values):
        """[world][len(values)] float64 on every rank."""
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [t.cpu().tolist()]
        out = self.torch.zeros((self.world, len(values)), dtype=self.torch.float64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().tolist()


def shard_starts(n_local, rank, world, sharding_mode='spread'):
    """(start chronic, first row) of the envs of a rank.  'spread' (default): every GPU's batch samples the whole data
    set (sharding.env_starts_spread) -- the per-GPU workload is statistically the same at 1, 2, 4 and 8 GPUs.  For
    analysis: 'blocks' = rank r owns the contiguous block [r B, (r+1) B) of a global batch laid out as env e -> chronic
    e mod 12, row (e // 12) mod 719 (one window of ~341 consecutive rows per GPU); 'strided' = env e -> GPU e mod n of the
    same global batch.  With those two the slowest env of a step -- which is what a step lasts -- depends on the window a
    GPU happens to hold (profiles/r2e_sharding_workload_effect.txt)."""
    from pypownet_b200 import sharding
    if sharding_mode == 'spread':
        return sharding.env_starts_spread(N_CHRONICS, N_ROWS, n_local, rank)
    ids = rank + world * np.arange(int(n_local)) if sharding_mode == 'strided' else rank * int(n_local) + np.arange(int(n_local))
    return sharding.env_starts_of(N_CHRONICS, N_ROWS, ids)


def measure(ctx, grid, envs, agent, cascade, steps, warmup, with_e2e=True, sampler=None, sharding_mode='spread',
            emulate=None):
    """One workload on this process' GPU (and, under torchrun, on every rank at once: weak scaling).  Returns a dict."""
    torch = ctx.torch
    from pypownet_b200.vec_env import VecRunEnv
    from pypownet_b200 import sharding
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    case, cfg, chronics, imaps = build_workload(grid, cascade=cascade)
    B = envs
    # weak scaling: B envs per GPU; `emulate` = (rank, world) plays that shard's envs on this GPU alone (analysis)
    sc, sr = shard_starts(B, emulate[0], emulate[1], sharding_mode) if emulate else shard_starts(B, rank, world, sharding_mode)
    env = VecRunEnv(case, cfg, chronics, B, device=ctx.local, reward_constant=float(case.n_sub), thermal_limits=imaps,
                    start_chronics=sc, start_rows=sr)
    actions = torch.zeros((B, case.action_length), dtype=torch.uint8, device=dev)      # do-nothing agent
    action_bank = None
    if agent == 'random':
        action_bank = torch.from_numpy(random_action_bank(case, B, seed=1234 + rank)).to(dev)
    # Sharded runs: every rank's step kernel stores its packed (reward[5], done, flag) rows straight into rank 0's
    # GPU memory over NVLink (sharding.PeerGather); rank 0 copies the rows of step t-1 to the host on a side stream
    # while step t computes.  No collective between two steps: ranks never run in lock-step.
    pg = sharding.PeerGather(env, rank, world, ring=16) if world > 1 else None   # 16 slots: a rank may run 16 steps ahead
    tstep = [0]
    if os.environ.get('PPN_BENCH_PG', '') == 'none':   # analysis only: no result gather at all
        pg = None

    def one_step():
        t = tstep[0]
        tstep[0] += 1
        a = actions if action_bank is None else action_bank[t % 16]
        if pg is not None:
            pg.before_step(t)
        obs, reward, done, flag = env.step(a, auto_reset=True)
        if pg is not None:
            pg.after_step(t)
            if t >= 1:
                pg.collect(t - 1)
        return done

    def drain():
        """the rows of the last step reach the host inside the timed region too"""
        if pg is not None and pg.is_root:
            pg.collect(tstep[0] - 1)
            torch.cuda.current_stream().wait_stream(pg.side)

    if sampler:
        # before the warm-up steps: launching nvidia-smi (fork + NVML start-up) stalls this process and the GPU's driver for
        # milliseconds -- inside the timed loop that was one 1-5 ms step on rank 0 (profiles/r2_scaling.txt, visit r2p)
        sampler.start()
    warm = max(warmup, 3)
    for _ in range(warm):
        one_step()
        ctx.flush.fill_(1)
    if pg is not None and pg.is_root:
        pg.collect(tstep[0] - 1)
        pg.wait_all()
    torch.cuda.synchronize()
    c0 = env.counters()
    ctx.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    n_done = torch.zeros((), dtype=torch.int64, device=dev)
    wall0 = time.perf_counter()
    if pg is not None:
        # the root's collect() calls of the timed loop start at the first timed step
        pass
    for k in range(steps):
        ev[k][0].record()
        d = one_step()
        if k == steps - 1:
            drain()
        ev[k][1].record()
        n_done += d.sum()
        ctx.flush.fill_(k & 1)                           # evict state/observation/chronics from L2 (126 MB)
    torch.cuda.synchronize()
    my_wall = time.perf_counter() - wall0
    ctx.barrier()
    wall = time.perf_counter() - wall0
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms = sum(per_step)
    ms_max = ctx.max_over_ranks(ms)
    c1 = env.counters()
    launches = c1['kernel_launches'] - c0['kernel_launches']
    value = world * B * steps / (ms_max * 1e-3)
    per_rank = ctx.gather_list([ms / steps, max(per_step), my_wall * 1e3 / steps])

    # ---- warm-L2 back-to-back figure (the natural RL loop: state stays in L2 between steps)
    ctx.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(steps):
        one_step()
    drain()
    a1.record()
    torch.cuda.synchronize()
    warm_value = world * B * steps / (ctx.max_over_ranks(a0.elapsed_time(a1)) * 1e-3)
    if pg is not None:
        pg.wait_all()

    out = {'grid': grid, 'envs_per_gpu': B, 'agent': agent, 'cascade': cascade, 'value': value,
           'ms_per_step': ms_max / steps, 'warm_l2_value': warm_value, 'launches': launches, 'wall': wall,
           'per_rank_ms_per_step': [round(r[0], 4) for r in per_rank],
           'per_rank_slowest_step_ms': [round(r[1], 4) for r in per_rank],
           'per_rank_host_ms_per_step': [round(r[2], 4) for r in per_rank]}
    steps_done = c1['env_steps'] - c0['env_steps']
    out['counters'] = {
        'loadflows_per_env_step': (c1['loadflows'] - c0['loadflows']) / max(steps_done, 1),
        'fd_iterations_per_loadflow': (c1['fd_iterations'] - c0['fd_iterations']) / max(c1['loadflows'] - c0['loadflows'], 1),
        'game_over_rate': float(n_done.item()) / (B * steps),
        'max_cascade_depth': c1['max_cascade_depth'],
        'cascade_depth_histogram': c1['cascade_depth_histogram'],
        'threads_per_env': c1['threads_per_env'], 'smem_bytes_per_env': c1['smem_bytes_per_env']}
    A = algorithmic_bytes(case)
    peak, peak_src = hbm_peak()
    kernel_ms = ms_max / steps
    achieved = A * B / (kernel_ms * 1e-3) / 1e9
    out['algorithmic_bytes_per_env_step'] = A
    out['roofline'] = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                       'traffic': traffic_of(grid, B, agent), 'peak_source': peak_src,
                       'kernel': 'ppn_step_kernel (step + restart launches of one env-step)', 'kernel_ms': kernel_ms}

    # ---- end to end through the public API with host buffers (pinned), copies inside the timed region
    if with_e2e:
        act_pinned = torch.zeros((B, case.action_length), dtype=torch.uint8).pin_memory()
        host_bank = [action_bank[k].cpu().pin_memory() for k in range(16)] if action_bank is not None else None
        for _ in range(3):
            env.step_pinned(act_pinned)
        ctx.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            po, pr, pd, pf = env.step_pinned(host_bank[k % 16] if host_bank is not None else act_pinned)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_all = ctx.gather_list([e2e_s])
        e2e_value = world * B * steps / max(r[0] for r in e2e_all)
        h2d = B * case.action_length
        d2h = B * (case.obs_dynamic_length * 8 + 5 * 8 + 1 + 4)
        out['e2e'] = {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h}
        out['per_rank_e2e_ms_per_step'] = [round(1e3 * r[0] / steps, 4) for r in e2e_all]
        out['per_rank_e2e_d2h_gbs'] = [round(d2h * steps / r[0] / 1e9, 2) for r in e2e_all]
        # the same loop with float32 observation rows (ppn_step_host_f32: an extension, NOT the headline -- the
        # reference's observation is float64)
        for _ in range(3):
            env.step_pinned(act_pinned, obs_dtype=torch.float32)
        ctx.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            env.step_pinned(host_bank[k % 16] if host_bank is not None else act_pinned, obs_dtype=torch.float32)
        torch.cuda.synchronize()
        f32_all = ctx.gather_list([time.perf_counter() - t0])
        out['e2e_f32'] = {'value': world * B * steps / max(r[0] for r in f32_all), 'unit': 'env-steps/s',
                          'd2h_bytes_per_step': B * (case.obs_dynamic_length * 4 + 5 * 8 + 1 + 4)}
    if pg is not None:
        pg.close()
    env.close()
    del env
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import __graft_entry__ as graft
    ctx = Ctx()
    if ctx.rank == 0:
        graft.build()
    ctx.barrier()
    world, rank = ctx.world, ctx.rank
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    emulate = tuple(int(x) for x in args.emulate_shard.split('/')) if args.emulate_shard else None
    m = measure(ctx, args.grid, args.envs, args.agent, args.cascade, args.steps, args.warmup, sampler=sampler,
                sharding_mode=args.sharding, emulate=emulate)
    # ---- the other BASELINE configurations, measured in the same run (fewer steps): configs[2] IEEE-30 with the
    # cascading-failure loop firing, configs[3] IEEE-118 do-nothing, configs[4] IEEE-118 with the random agent.  Under
    # `--gpus 8` configs[3] is 65536 envs and configs[4] 32768 envs over the 8 GPUs.
    secondary = []
    if not args.no_secondary:
        k = max(10, min(args.steps // 4, 40))
        for name, grid, envs, agent, cascade in (
                ('configs[2] default30 AC + cascade, 8192 envs per GPU', 'case30', 8192, 'nothing', True),
                ('configs[3] default118 AC, 8192 envs per GPU (65536 over 8 GPUs)', 'case118', 8192, 'nothing', False),
                ('configs[4] default118 AC, random node-split + line-switch, 4096 envs per GPU (32768 over 8 GPUs)',
                 'case118', 4096, 'random', False)):
            r = measure(ctx, grid, envs, agent, cascade, k, 3, with_e2e=True, sharding_mode=args.sharding)
            secondary.append({'workload': name, 'grid': grid, 'envs_per_gpu': envs, 'n_gpus': world, 'steps': k,
                              'value': r['value'], 'unit': 'env-steps/s', 'ms_per_step': r['ms_per_step'],
                              'e2e': r['e2e']['value'], 'e2e_float32_observations': r['e2e_f32']['value'],
                              'roofline_frac': r['roofline']['frac'],
                              'algorithmic_bytes_per_env_step': r['algorithmic_bytes_per_env_step'],
                              'gpu_launches': r['launches'], 'per_rank_ms_per_step': r['per_rank_ms_per_step'],
                              'counters': r['counters']})
    # one GPU only: the same kernel on round 1's env starts (contiguous row windows), for round-to-round comparison
    r1w = None
    if world == 1 and not args.no_secondary and args.sharding == 'spread' and not args.emulate_shard:
        r = measure(ctx, args.grid, args.envs, args.agent, args.cascade, max(10, min(args.steps, 50)), 3, with_e2e=False,
                    sharding_mode='blocks')
        r1w = {'value': r['value'], 'ms_per_step': r['ms_per_step'],
               'note': 'env e -> chronic e mod 12, row e // 12 (the windows BENCH_r01 was measured on)'}
    clocks = sampler.stop() if sampler else None          # sampled over every timed loop of this run
    if args.profile_ranks and rank == 0:
        with open(args.profile_ranks, 'a') as f:
            f.write('# n_gpus=%d grid=%s envs_per_gpu=%d agent=%s: per rank  kernel ms/step | slowest step ms | host ms/step'
                    ' | e2e ms/step | e2e D2H GB/s\n' % (world, args.grid, args.envs, args.agent))
            for r in range(world):
                f.write('rank %d  %.4f  %.4f  %.4f  %.4f  %.2f\n' % (
                    r, m['per_rank_ms_per_step'][r], m['per_rank_slowest_step_ms'][r], m['per_rank_host_ms_per_step'][r],
                    m['per_rank_e2e_ms_per_step'][r], m['per_rank_e2e_d2h_gbs'][r]))
    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return
    extra = dict(m['counters'])
    extra.update({'algorithmic_bytes_per_env_step': m['algorithmic_bytes_per_env_step'],
                  'warm_l2_value': m['warm_l2_value'], 'wall_s_timed_region': m['wall'],
                  'sharding': 'one GPU' if world == 1 else
                  {'spread': 'every GPU samples the whole data set (rows spread over the chronics, rank r shifted by 97 r rows)',
                   'blocks': 'contiguous blocks of the global batch', 'strided': 'round-robin (env e -> GPU e mod n)'}[args.sharding],
                  'result_gather': 'none (one GPU)' if world == 1 else
                  'step kernels store their reward/done/flag rows into rank 0\'s GPU memory over NVLink (peer mapping); '
                  'rank 0 copies them to the host one step behind on a side stream; NCCL only for set-up and timing',
                  'per_rank_ms_per_step': m['per_rank_ms_per_step'],
                  'per_rank_e2e_d2h_gbs': m['per_rank_e2e_d2h_gbs'],
                  'e2e_float32_observations': m['e2e_f32']})
    if r1w:
        extra['same_kernel_on_round1_env_starts'] = r1w
    line = {'metric': 'env steps/sec (batched grids)', 'value': m['value'], 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': m['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, extra),
            'clocks': clocks,
            'e2e': m['e2e'],
            'gpu_launches': m['launches'],
            'roofline': m['roofline']}
    if secondary:
        line['secondary'] = secondary
    if world == 1 and not args.no_cpu:
        pool, cores, kind = cpu_pool(args.grid)
        v, n, busy = cpu_baseline(pool, cores, 10 ** 9, budget_s=args.cpu_seconds)
        pool.close()
        line['cpu_baseline'] = {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': kind,
                                'sample': '%d processes x %.0f s of do-nothing env-steps of %s (%s), %d env-steps in total'
                                          % (cores, args.cpu_seconds, args.grid,
                                             'the unmodified reference package in baseline/_ref on oracle/shims'
                                             if kind == 'reference' else
                                             'oracle/flat.py, the CPU restatement of the reference path', n)}
    print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--grid', default='case14', choices=sorted(GRIDS))
    ap.add_argument('--envs', type=int, default=4096, help='envs per GPU')
    ap.add_argument('--agent', default='nothing', choices=['nothing', 'random'],
                    help="'random': one random node-splitting + one line switch per env and step (BASELINE configs[4])")
    ap.add_argument('--cascade', action='store_true',
                    help='synthetic thermal limits that make the cascading-failure loop fire (BASELINE configs[2])')
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the other BASELINE configurations')
    ap.add_argument('--sharding', default='spread', choices=['spread', 'blocks', 'strided'],
                    help='which chronic rows the envs of a GPU start on (see shard_starts)')
    ap.add_argument('--emulate-shard', default=None, metavar='R/W',
                    help='one GPU plays the envs rank R of a W-GPU run would play (analysis of the workload effect)')
    ap.add_argument('--profile-ranks', default=None, metavar='FILE',
                    help='append the per-rank timing table (kernel, slowest step, host, end-to-end, PCIe rate) to FILE')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
