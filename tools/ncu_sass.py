"""Hottest SASS instructions of a profile with their stall reasons: python tools/ncu_sass.py <rep> [top_n] [only-executed-by-few-warps]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
tot_inst = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print('samples', tot, 'instructions', tot_inst)
order = sorted(range(len(data)), key=lambda k: -int(data[k][ix['# Samples']] or 0))[:top]
for k in sorted(order):
    r = data[k]; n = int(r[ix['# Samples']] or 0)
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print('%6d %5.1f%% inst %9s | %-60s | %s' % (k, 100.0 * n / tot, r[ix['Instructions Executed']], r[ix['Source']].strip()[:60], ' '.join('%s=%d' % (b, a) for a, b in st if a)))
