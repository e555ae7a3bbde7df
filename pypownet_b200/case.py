"""Grid case loading and the static tables every backend (CUDA library, oracle) is built from.

Reads the reference's `reference_grid.py` format (a python file whose single function returns
{'baseMVA','bus','gen','branch'} built with a bare `array(...)`, /root/reference/parameters/default14/level0/
reference_grid.py:1-78; PYPOWER `loadcase` exec's it) and derives the fixed-shape model implied by
pypownet/grid.py:74-93, 428-494 and SURVEY.md Appendix C:

  * S substations -> 2S buses: bus s (node 0) and its artificial sister '666'+id (node 1), row s+S;
  * <=1 generator and <=1 load per substation; loads are the buses with non-zero Pd or Qd (grid.py:77);
  * an element never changes substation, only its node bit, so topology is (G+L+2N) bits + N line status.
"""
import json
import os

import numpy as np

ARTIFICIAL_NODE_STARTING_STRING = '666'          # pypownet/__init__.py:10

# MATPOWER column indices used here
BUS_I, BUS_TYPE, PD, QD, GS, BS, _AREA, VM, VA, BASE_KV = range(10)
GEN_BUS, PG, QG, QMAX, QMIN, VG, MBASE, GEN_STATUS = range(8)
F_BUS, T_BUS, BR_R, BR_X, BR_B, RATE_A, RATE_B, RATE_C, TAP, SHIFT, BR_STATUS = range(11)


def read_case_file(path):
    """exec a `reference_grid.py`-style file and return its dict with float arrays."""
    scope = {'array': np.array, 'np': np}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), scope)
    funcs = [v for v in scope.values() if hasattr(v, '__code__') and v.__code__.co_filename == path]
    if not funcs:
        raise ValueError('%s defines no case function' % path)
    ppc = funcs[-1]()
    return {'baseMVA': float(ppc['baseMVA']), 'bus': np.array(ppc['bus'], dtype=float),
            'gen': np.array(ppc['gen'], dtype=float), 'branch': np.array(ppc['branch'], dtype=float)}


def write_case_file(path, ppc):
    """Write a case dict in the reference's `reference_grid.py` layout (bare `array(`)."""
    def rows(a):
        return ',\n'.join('        [%s]' % ', '.join(repr(float(v)) if float(v) != int(v) else str(int(v))
                                                     for v in r) for r in a)
    with open(path, 'w') as f:
        f.write('def reference_grid():\n    ppc = {"version": "2"}\n    ppc["baseMVA"] = %r\n' % ppc['baseMVA'])
        for k in ('bus', 'gen', 'branch'):
            f.write('    ppc["%s"] = array([\n%s,\n    ])\n' % (k, rows(ppc[k])))
        f.write('    return ppc\n')


class Case(object):
    """Static description of one grid family, in struct-of-arrays form (the `ppn_case` of include/pypownet_b200.h)."""

    def __init__(self, ppc):
        bus, gen, br = ppc['bus'], ppc['gen'], ppc['branch']
        if len(bus) % 2:
            raise ValueError('expected 2S bus rows (real + artificial buses)')
        S = len(bus) // 2
        ids = bus[:S, BUS_I].astype(np.int64)
        for s in range(S):
            if int(bus[s + S, BUS_I]) != int(ARTIFICIAL_NODE_STARTING_STRING + str(ids[s])):
                raise ValueError('bus row %d is not the artificial sister of bus %d' % (s + S, ids[s]))
        sub_of_id = {int(b): s for s, b in enumerate(ids)}
        self.base_mva = float(ppc['baseMVA'])
        self.n_sub, self.n_gen, self.n_line = S, len(gen), len(br)
        self.sub_ids = ids.astype(np.int32)
        if np.any(br[:, SHIFT] != 0):
            raise ValueError('phase shifters are not supported (all shipped grids have SHIFT = 0)')

        def sub_and_node(bus_id):
            bus_id = int(bus_id)
            if bus_id in sub_of_id:
                return sub_of_id[bus_id], 0
            txt = str(bus_id)
            if txt.startswith(ARTIFICIAL_NODE_STARTING_STRING) and int(txt[3:]) in sub_of_id:
                return sub_of_id[int(txt[3:])], 1
            raise ValueError('unknown bus id %d' % bus_id)

        gs = [sub_and_node(b) for b in gen[:, GEN_BUS]]
        self.gen_sub = np.array([g[0] for g in gs], dtype=np.int32)
        self.gen_node0 = np.array([g[1] for g in gs], dtype=np.uint8)
        if len(set(self.gen_sub.tolist())) != len(gs):
            raise ValueError('at most one generator per substation is supported (grid.py:474-475)')
        are_loads = (bus[:, PD] != 0) | (bus[:, QD] != 0)                      # grid.py:77
        load_rows = np.flatnonzero(are_loads)
        self.load_sub = (load_rows % S).astype(np.int32)
        self.load_node0 = (load_rows // S).astype(np.uint8)
        if len(set(self.load_sub.tolist())) != len(load_rows):
            raise ValueError('at most one load per substation is supported (grid.py:477-478)')
        if np.any(np.diff(self.load_sub) <= 0) or np.any(self.load_node0):
            raise ValueError('loads must sit on the real buses, ascending')
        self.n_load = len(load_rows)
        ors = [sub_and_node(b) for b in br[:, F_BUS]]
        exs = [sub_and_node(b) for b in br[:, T_BUS]]
        self.line_or_sub = np.array([o[0] for o in ors], dtype=np.int32)
        self.line_ex_sub = np.array([e[0] for e in exs], dtype=np.int32)
        if any(o[1] for o in ors) or any(e[1] for e in exs) or np.any(self.gen_node0):
            raise ValueError('the reference grid must start with every element on node 0')
        self.line_r, self.line_x, self.line_b = br[:, BR_R].copy(), br[:, BR_X].copy(), br[:, BR_B].copy()
        self.line_tap = np.where(br[:, TAP] != 0, br[:, TAP], 1.0)
        self.line_status0 = (br[:, BR_STATUS] != 0).astype(np.uint8)
        self.bus_gs, self.bus_bs = bus[:, GS].copy(), bus[:, BS].copy()
        self.bus_basekv = bus[:, BASE_KV].copy()
        self.bus_vm0, self.bus_va0 = bus[:, VM].copy(), bus[:, VA].copy()      # VA in degrees, as stored
        self.bus_pd0, self.bus_qd0 = bus[:, PD].copy(), bus[:, QD].copy()
        self.gen_qmax, self.gen_qmin = gen[:, QMAX].copy(), gen[:, QMIN].copy()
        self.gen_pg0, self.gen_qg0, self.gen_vg0 = gen[:, PG].copy(), gen[:, QG].copy(), gen[:, VG].copy()
        ref_rows = np.flatnonzero(bus[:, BUS_TYPE] == 3)
        if len(ref_rows) == 0 or ref_rows[0] >= S:
            raise ValueError('the case needs a type-3 (slack) real bus')
        self.slack_sub = int(ref_rows[0])                                        # grid.py:74
        self.ppc = ppc
        # element -> substation map of the action/topology vector, prods|loads|lines or|lines ex (grid.py:428-494)
        self.elem_sub = np.concatenate([self.gen_sub, self.load_sub, self.line_or_sub, self.line_ex_sub]) \
            .astype(np.int32)
        self.n_elements_per_sub = np.bincount(self.elem_sub, minlength=S).astype(np.int32)

    # sizes --------------------------------------------------------------------------------------------------
    @property
    def action_length(self):
        return self.n_gen + self.n_load + 3 * self.n_line

    @property
    def obs_length(self):
        G, L, N, S = self.n_gen, self.n_load, self.n_line, self.n_sub
        return 9 * L + 9 * G + 18 * N + 2 * S + 6

    @property
    def obs_dynamic_length(self):
        G, L, N, S = self.n_gen, self.n_load, self.n_line, self.n_sub
        return 7 * L + 7 * G + 13 * N + S + 6

    @classmethod
    def from_file(cls, path):
        return cls(read_case_file(path))

    @classmethod
    def builtin(cls, name):
        """One of the packaged IEEE grids ('case14', 'case30', 'case118'), stored as JSON under data/."""
        path = os.path.join(os.path.dirname(__file__), 'data', name + '.json')
        with open(path) as f:
            d = json.load(f)
        return cls({'baseMVA': d['baseMVA'], 'bus': np.array(d['bus'], dtype=float),
                    'gen': np.array(d['gen'], dtype=float), 'branch': np.array(d['branch'], dtype=float)})
