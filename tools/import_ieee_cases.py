"""Writes pypownet_b200/data/case{14,30,118}.json from the reference's shipped grids (public IEEE test cases in
the 2S-bus layout of parameters/make_reference_grid.py:25-57) and the per-line thermal limits of each
environment's first chronic.  Run once in the build container (the reference tree does not travel to the GPU box):
    python tools/import_ieee_cases.py /root/reference
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from pypownet_b200.case import read_case_file  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
out = os.path.join(os.path.dirname(__file__), '..', 'pypownet_b200', 'data')
for env, name in (('default14', 'case14'), ('default30', 'case30'), ('default118', 'case118')):
    lvl = os.path.join(ref, 'parameters', env, 'level0')
    ppc = read_case_file(os.path.join(lvl, 'reference_grid.py'))
    imaps = np.genfromtxt(os.path.join(lvl, 'chronics', 'a', '_N_imaps.csv'), dtype=np.float32, delimiter=';',
                          skip_header=True)
    d = {'baseMVA': ppc['baseMVA'], 'bus': ppc['bus'].tolist(), 'gen': ppc['gen'].tolist(),
         'branch': ppc['branch'].tolist(), 'imaps': [float(v) for v in imaps]}
    with open(os.path.join(out, name + '.json'), 'w') as f:
        json.dump(d, f, separators=(',', ':'))
    print(name, len(ppc['bus']) // 2, 'substations', len(ppc['gen']), 'gens', len(ppc['branch']), 'lines')
