"""Benchmark of the batched pypownet step path on B200 (BASELINE.json metric: env steps/s on batched grids).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--grid case14] [--envs 4096]

N=1 workload = BASELINE.json configs[1]: default14 AC, 4096 batched envs, do-nothing agent, one B200; envs start on
different chronics/rows, games that end are restarted in the same step (Runner semantics).  The reference's chronics
do not travel with the repo: chronics are synthetic with the shipped ones' statistics (pypownet_b200/synthetic.py).
A "step" is one env-step of every env of the batch (one fused kernel launch).  One JSON line on stdout (rank 0).
`--impl reference` times the CPU restatement of the reference's path (oracle/flat.py, the reference package itself
is not present on the GPU box) on all host cores for the same workload, on a bounded sample per step.
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


# ------------------------------------------------------------------------------------------------ CPU baseline (port)
_WORKER = {}


def _cpu_init(grid, counter):
    """Pool initializer: every process builds the workload and ONE env once (as a reference process would)."""
    sys.path.insert(0, ROOT)
    from oracle.flat import FlatEnv, Config
    with counter.get_lock():
        env_id = counter.value
        counter.value += 1
    case, cfg, chronics, imaps = build_workload(grid)
    c, r = env_starts(1, 97 * env_id)
    env = FlatEnv(case, Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub), chronics,
                  start_id=int(c[0]), thermal_limits=imaps, start_row=int(r[0]))
    a = np.zeros(case.action_length, dtype=np.uint8)
    for _ in range(5):
        if env.step(a)[2]:
            env.process_game_over()
    _WORKER['env'], _WORKER['action'] = env, a


def _cpu_worker(args):
    n_steps, budget_s = args
    env, a = _WORKER['env'], _WORKER['action']
    t0 = time.perf_counter()
    done_steps = 0
    while done_steps < n_steps and (budget_s is None or time.perf_counter() - t0 < budget_s):
        if env.step(a)[2]:
            env.process_game_over()
        done_steps += 1
    return done_steps, time.perf_counter() - t0


def cpu_pool(grid, cores=None):
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(grid, ctx.Value('i', 0)))
    pool.map(_cpu_worker, [(1, None)] * cores)          # make sure every process is up before anything is timed
    return pool, cores


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
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
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
    """CPU arm: oracle/flat.py (port of the reference's step path) on every host core, do-nothing agent."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = 100                                       # env-steps per process per bench step (bounded sample)
    pool, cores = cpu_pool(args.grid)
    for _ in range(args.warmup):
        cpu_baseline(pool, cores, 10)
    total = 0
    busy = 0.0
    for _ in range(args.steps):
        v, n, b = cpu_baseline(pool, cores, per_step)
        total += n
        busy += b
    pool.close()
    value = total / busy
    sample = '%d processes x %d do-nothing env-steps of %s per bench step, synthetic chronics' % (cores, per_step,
                                                                                                  args.grid)
    line = {'impl': 'reference', 'metric': 'env steps/sec (batched grids)', 'value': value, 'unit': 'env-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * busy / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, None),
            'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(args, extra):
    cfg = {'workload': '%s AC, %d batched envs per GPU, %s agent, auto-restart on game over '
                       '(BASELINE.json configs[1] shape)' % (GRIDS[args.grid], args.envs,
                                                            'do-nothing' if args.agent == 'nothing' else
                                                            'random node-split + line-switch'),
           'grid': args.grid, 'envs_per_gpu': args.envs, 'chronics': '%d synthetic x %d rows' % (N_CHRONICS, N_ROWS),
           'solver': 'fast-decoupled XB, tol 1e-6, <=25 it (the reference\'s PF_ALG=2)',
           'l2': 'flushed between timed steps (256 MiB write)', 'parallelism': 'env-sharded, dp%d' % args.gpus}
    if extra:
        cfg.update(extra)
    return cfg


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()
    from pypownet_b200.vec_env import VecRunEnv
    from pypownet_b200 import sharding
    dev = torch.device('cuda', local)
    case, cfg, chronics, imaps = build_workload(args.grid)
    B = args.envs
    # weak scaling: the global batch of world * B envs is dealt round-robin, rank r owns envs r, r + world, ...
    sc, sr = sharding.env_starts_of(N_CHRONICS, N_ROWS, sharding.strided_env_ids(B, rank, world))
    env = VecRunEnv(case, cfg, chronics, B, device=local, reward_constant=float(case.n_sub), thermal_limits=imaps,
                    start_chronics=sc, start_rows=sr)
    actions = torch.zeros((B, case.action_length), dtype=torch.uint8, device=dev)      # do-nothing agent
    action_bank = None
    if args.agent == 'random':
        action_bank = torch.from_numpy(random_action_bank(case, B, seed=1234 + rank)).to(dev)
    step_counter = [0, 0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # Sharded runs: the step kernel writes the packed (reward[5], done, flag) rows itself and one NCCL all-gather per
    # step brings every shard's rows to every rank.  The gather of step t runs on NCCL's stream while step t+1 is
    # already computing (two pack / result buffers); it is waited for inside the timed interval of step t+1, the last
    # one before the closing synchronisation, so every collective completes inside the timed region.
    packs = [torch.zeros((B, sharding.PACK_WIDTH), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    gathered = [torch.zeros((world * B, sharding.PACK_WIDTH), dtype=torch.float64, device=dev) for _ in range(2)] \
        if world > 1 else None
    works = [None, None]

    def one_step():
        a = actions
        if action_bank is not None:
            a = action_bank[step_counter[0] % 16]
            step_counter[0] += 1
        if world > 1:
            buf = step_counter[1] & 1
            step_counter[1] += 1
            env.enable_result_pack(packs[buf])
        obs, reward, done, flag = env.step(a, auto_reset=True)
        if world > 1:      # rewards / dones / flags of every shard on every rank (NCCL over NVLink)
            if works[buf ^ 1] is not None:
                works[buf ^ 1].wait()          # the previous step's gather (stream wait, not a host block)
            works[buf] = dist.all_gather_into_tensor(gathered[buf], packs[buf], async_op=True)
        return done

    def drain():
        for w in works:
            if w is not None:
                w.wait()

    for _ in range(max(args.warmup, 3)):
        one_step()
        flush.fill_(1)
    torch.cuda.synchronize()
    c0 = env.counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n_done = torch.zeros((), dtype=torch.int64, device=dev)
    wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        d = one_step()
        if k == args.steps - 1:
            drain()                                      # the last gather also completes inside the timed region
        ev[k][1].record()
        n_done += d.sum()
        flush.fill_(k & 1)                               # evict state/observation/chronics from L2 (126 MB)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    c1 = env.counters()
    launches = c1['kernel_launches'] - c0['kernel_launches']
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- warm-L2 back-to-back figure (the natural RL loop: state stays in L2 between steps)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
        one_step()
    drain()
    a1.record()
    torch.cuda.synchronize()
    warm_value = world * B * args.steps / (a0.elapsed_time(a1) * 1e-3)

    # ---- end to end through the public API with host buffers (pinned), copies inside the timed region
    act_pinned = torch.zeros((B, case.action_length), dtype=torch.uint8).pin_memory()
    host_bank = [action_bank[k].cpu().pin_memory() for k in range(16)] if action_bank is not None else None
    for _ in range(3):
        env.step_pinned(act_pinned)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        po, pr, pd, pf = env.step_pinned(host_bank[k % 16] if host_bank is not None else act_pinned)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    clocks = sampler.stop() if sampler else None      # sampled over the timed, back-to-back and end-to-end loops
    h2d = B * case.action_length
    d2h = B * (case.obs_dynamic_length * 8 + 5 * 8 + 1 + 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    A = algorithmic_bytes(case)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    kernel_ms = ms_max / args.steps
    achieved = A * B / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get('%s_%d' % (args.grid, B))
    steps_done = c1['env_steps'] - c0['env_steps']
    extra = {'loadflows_per_env_step': (c1['loadflows'] - c0['loadflows']) / max(steps_done, 1),
             'fd_iterations_per_loadflow': (c1['fd_iterations'] - c0['fd_iterations']) /
             max(c1['loadflows'] - c0['loadflows'], 1),
             'game_over_rate': float(n_done.item()) / (B * args.steps),
             'max_cascade_depth': c1['max_cascade_depth'], 'threads_per_env': c1['threads_per_env'],
             'smem_bytes_per_env': c1['smem_bytes_per_env'], 'algorithmic_bytes_per_env_step': A,
             'warm_l2_value': warm_value, 'wall_s_timed_region': wall}
    line = {'metric': 'env steps/sec (batched grids)', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': kernel_ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, extra),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src, 'kernel': 'ppn_step_kernel',
                         'kernel_ms': kernel_ms}}
    if world == 1 and not args.no_cpu:
        pool, cores = cpu_pool(args.grid)
        v, n, busy = cpu_baseline(pool, cores, 10 ** 9, budget_s=args.cpu_seconds)
        pool.close()
        line['cpu_baseline'] = {'value': v, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d processes x %.0f s of do-nothing env-steps of %s (oracle/flat.py, the '
                                          'CPU restatement of the reference path), %d env-steps in total'
                                          % (cores, args.cpu_seconds, args.grid, n)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
