"""Chronic (time-series) ingestion: the HBM-resident float32 read stream of the step kernels.

Follows pypownet/chronic.py:124-246: 13 ';'-separated CSVs per chronic with one ignored header row, values
parsed as float32 (:175), `planned[t] := planned[t+1]` (:202-205), rows zipped to the shortest file (:225-232),
`timesteps_before_planned_maintenance` = argmax over the next `horizon` rows of maintenance != 0 (:239-246).
"""
import os
from datetime import datetime

import numpy as np

_FILES = {
    'loads_p': '_N_loads_p.csv', 'loads_q': '_N_loads_q.csv', 'prods_p': '_N_prods_p.csv',
    'prods_v': '_N_prods_v.csv', 'loads_p_planned': '_N_loads_p_planned.csv',
    'loads_q_planned': '_N_loads_q_planned.csv', 'prods_p_planned': '_N_prods_p_planned.csv',
    'prods_v_planned': '_N_prods_v_planned.csv', 'ids': '_N_simu_ids.csv', 'imaps': '_N_imaps.csv',
    'maintenance': 'maintenance.csv', 'hazards': 'hazards.csv',
}
_DATETIMES = '_N_datetimes.csv'


def _read_csv_f32(path):
    """float32 matrix of a ';' CSV, header skipped.  Parsed through float64 then narrowed, like np.genfromtxt with
    dtype=float32 does (chronic.py:174-175)."""
    with open(path) as f:
        lines = f.read().splitlines()[1:]
    rows = [ln.split(';') for ln in lines if ln.strip()]
    return np.array(rows, dtype=np.float64).astype(np.float32)


class Chronic(object):
    """One chronic as dense float32 tables with T rows (T = shortest of the files)."""

    def __init__(self, name, prods_p, prods_v, loads_p, loads_q, prods_p_planned, prods_v_planned,
                 loads_p_planned, loads_q_planned, maintenance, hazards, ids, datetimes, imaps):
        tabs = [np.atleast_2d(np.asarray(a, dtype=np.float32)) for a in
                (prods_p, prods_v, loads_p, loads_q, prods_p_planned, prods_v_planned, loads_p_planned,
                 loads_q_planned, maintenance, hazards)]
        ids = np.asarray(ids).astype(np.int32).ravel()
        T = min([len(a) for a in tabs] + [len(ids), len(datetimes)])
        # planned rows are shifted BEFORE truncation (chronic.py:202-205 then :225-232)
        for k in (4, 5, 6, 7):
            a = tabs[k].copy()
            a[:-1] = a[1:]
            tabs[k] = a
        (self.prods_p, self.prods_v, self.loads_p, self.loads_q, self.prods_p_planned, self.prods_v_planned,
         self.loads_p_planned, self.loads_q_planned, self.maintenance, self.hazards) = \
            [np.ascontiguousarray(a[:T]) for a in tabs]
        self.name = name
        self.ids = ids[:T].copy()
        if len(np.unique(ids)) != len(ids):
            raise ValueError('There are timesteps with the same id')
        self.datetimes = np.asarray(datetimes[:T], dtype=np.int32).reshape(T, 6)
        self.imaps = np.asarray(imaps, dtype=np.float32).astype(np.float64).ravel()
        self.n_rows = T
        for nm in ('maintenance', 'hazards'):
            a = getattr(self, nm)
            if np.any(a != np.round(a)) or np.any(a < 0):
                raise ValueError('%s durations must be non-negative integers' % nm)
        pos0 = np.flatnonzero(self.ids == 0)
        # row played first after a chronic change: get_next_chronic sets the id to 0, then "next id" (game.py:399, 492)
        self.row_after_switch = min(int(pos0[0]) + 1, T - 1) if len(pos0) else -1

    @classmethod
    def from_folder(cls, folder):
        if not os.path.exists(folder):
            raise ValueError('Source folder %s does not exist' % folder)
        present = set(os.listdir(folder))
        for fn in list(_FILES.values()) + [_DATETIMES]:
            if fn not in present:
                raise FileExistsError('File %s does not exist but is mandatory' % fn)
        d = {k: _read_csv_f32(os.path.join(folder, fn)) for k, fn in _FILES.items()}
        with open(os.path.join(folder, _DATETIMES)) as f:
            stamps = f.read().splitlines()[1:]
        dts = []
        for s in stamps:
            t = datetime.strptime(s.lower(), '%Y-%b-%d;%H:%M')                 # chronic.py:31
            dts.append((t.year, t.month, t.day, t.hour, t.minute, t.second))
        return cls(os.path.basename(os.path.normpath(folder)), d['prods_p'], d['prods_v'], d['loads_p'],
                   d['loads_q'], d['prods_p_planned'], d['prods_v_planned'], d['loads_p_planned'],
                   d['loads_q_planned'], d['maintenance'], d['hazards'], d['ids'], dts, d['imaps'])

    def planned_maintenance_table(self, horizon):
        """[T, N] int32: for each row t, index of the first row in [t, t+horizon) with maintenance != 0 (0 if none)."""
        T, N = self.maintenance.shape
        out = np.zeros((T, N), dtype=np.int32)
        nz = self.maintenance != 0
        for t in range(T):
            out[t] = nz[t:t + horizon].argmax(axis=0)
        return out


class ChronicSet(object):
    """All chronics of a level, alphabetical (ChronicLooper, chronic.py:260-295)."""

    def __init__(self, chronics):
        self.chronics = list(chronics)
        if not self.chronics:
            raise ValueError('no chronic')

    @classmethod
    def from_folder(cls, chronics_folder):
        chronics_folder = os.path.abspath(chronics_folder)
        if not os.path.exists(chronics_folder):
            raise FileNotFoundError('Chronic folder %s does not exist' % chronics_folder)
        names = sorted(d for d in os.listdir(chronics_folder)
                       if not os.path.isfile(os.path.join(chronics_folder, d)))
        return cls([Chronic.from_folder(os.path.join(chronics_folder, n)) for n in names])

    def __len__(self):
        return len(self.chronics)

    def __getitem__(self, i):
        return self.chronics[i]
