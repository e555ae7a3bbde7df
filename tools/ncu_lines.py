"""Per-function / per-line breakdown (stall samples, instructions) of a profile: python tools/ncu_lines.py <rep>"""
import csv, re, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; hdr = None; cur = None
for r in rows:
    if r and r[0] == 'File Path': cur = r[1]
    if r and r[0] == 'Line No': hdr = r; blocks.append([cur, []]); continue
    if hdr is None or not r: continue
    if r[0] != '':
        try: blocks[-1][1].append((int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0)))
        except Exception: pass
f, b = max((x for x in blocks if x[0] and x[0].endswith('ppn_kernels.cu')), key=lambda x: len(x[1]), default=blocks[0])
ts = sum(d[2] for d in b) or 1; ti = sum(d[3] for d in b) or 1
agg = {}
for l, s, sm, i in b:
    a = agg.setdefault(l, [s, 0, 0]); a[1] += sm; a[2] += i
src = open('/root/repo/pypownet_b200/csrc/ppn_kernels.cu').read().split('\n')
def func_of(line):
    for k in range(line - 1, 0, -1):
        t = src[k - 1]
        if t.startswith('__device__') or t.startswith('ppn_step_kernel') or (t.startswith('template') and '__device__' in t):
            return t[:90]
    return '?'
fa = {}
for l, (s, sm, i) in agg.items():
    fn = func_of(l); a = fa.setdefault(fn, [0, 0]); a[0] += sm; a[1] += i
print('samples', ts, 'instructions', ti)
for fn, (sm, i) in sorted(fa.items(), key=lambda kv: -kv[1][0])[:12]: print('%5.1f%% smp %5.1f%% inst  %s' % (100 * sm / ts, 100 * i / ti, fn))
print('--- top lines by samples')
for l, (s, sm, i) in sorted(sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]): print('%4d %5.1f%% smp %5.1f%% inst | %s' % (l, 100 * sm / ts, 100 * i / ti, s[:100]))
