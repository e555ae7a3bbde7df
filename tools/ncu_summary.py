"""Summarises gpurun_out/prof_<tag>.ncu-rep and launches_<tag>.csv into profiles/ (text + traffic.json).
    python tools/ncu_summary.py <tag> <grid> <envs> [agent]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
tag, grid, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
agent = sys.argv[4] if len(sys.argv) > 4 else 'nothing'
rep = os.path.join(ROOT, 'gpurun_out', 'prof_%s.ncu-rep' % tag)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__grid_size', 'launch__block_size', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.sum', 'launch__shared_mem_per_block_dynamic']
out = ['ncu --set full --clock-control none, kernel ppn_step_kernel, %s x %d envs (tag %s)' % (grid, envs, tag), '']
traffic = []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    out.append('launch: ' + name[:110])
    for w in want:
        if w in hdr:
            out.append('  %-70s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    for i, h in enumerate(hdr):   # local memory (register spills / stack): instructions and bytes
        if '_local' in h and h.endswith('.sum'):
            try:
                if float(r[i].replace(',', '')) > 0:
                    out.append('  %-70s %s %s' % (h, r[i], units[i]))
            except ValueError:
                pass
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                if float(r[i]) > 3:
                    out.append('  %-70s %s %%' % (h, r[i]))
            except ValueError:
                pass

    def byt(metric):
        v, u = float(r[hdr.index(metric)].replace(',', '')), units[hdr.index(metric)]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    traffic.append(byt('dram__bytes_read.sum') + byt('dram__bytes_write.sum'))
    out.append('')
lpath = os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % tag)
if os.path.exists(lpath):
    lr = [r for r in csv.reader(open(lpath)) if len(r) > 10]
    h2 = lr[0]
    k, v, u = h2.index('Kernel Name'), h2.index('Metric Value'), h2.index('Metric Unit')
    agg = {}
    for r in lr[1:]:
        val = float(r[v].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[u], 1)
        a = agg.setdefault(r[k][:60], [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(a[1] for a in agg.values())
    out.append('launch list of `bench.py --steps 20 --warmup 3 --no-cpu` (gpu__time_duration.sum, serialised, cold cache):')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('  %-62s n=%4d total %10.1f us  share %5.1f%%  avg %8.1f us' % (n, c, t, 100 * t / tot, t / c))
    import shutil
    shutil.copy(lpath, os.path.join(ROOT, 'profiles', '%s_launches.csv' % tag))
with open(os.path.join(ROOT, 'profiles', '%s_ncu_summary.txt' % tag), 'w') as f:
    f.write('\n'.join(out) + '\n')
tp = os.path.join(ROOT, 'profiles', 'traffic.json')
t = json.load(open(tp)) if os.path.exists(tp) else {}
t['%s_%d%s' % (grid, envs, '' if agent == 'nothing' else '_' + agent)] = sum(traffic) / len(traffic)
t.setdefault('_source', {})
if not isinstance(t['_source'], dict):
    t['_source'] = {}
t['_source']['%s_%d%s' % (grid, envs, '' if agent == 'nothing' else '_' + agent)] = 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, tag %s' % tag
json.dump(t, open(tp, 'w'), indent=1)
print('\n'.join(out))
