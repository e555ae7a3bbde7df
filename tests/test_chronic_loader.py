"""pypownet_b200.chronic against the reference's own parser on the SHIPPED chronics (chronic.py:124-246): same float32
tables, planned rows shifted by one, rows zipped to the shortest file, simu ids with gaps (default118: ids 0..168 over
167 rows), datetimes, imaps, planned-maintenance horizon.  Needs /root/reference (build container); skipped elsewhere."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = os.environ.get('PYPOWNET_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'parameters')),
                                reason='the reference tree is not present on this machine')


@pytest.fixture(scope='module')
def ref_chronic_cls():
    for p in (os.path.join(ROOT, 'oracle', 'shims'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging
    logging.disable(logging.CRITICAL)
    from pypownet.chronic import Chronic
    return Chronic


@pytest.mark.parametrize('env,name', [('default14', 'a'), ('default14', 'h'), ('default30', 'c'), ('default118', 'a'),
                                      ('default118', 'k')])
def test_loader_equals_the_reference_parser(ref_chronic_cls, env, name):
    from pypownet_b200.chronic import Chronic
    folder = os.path.join(REF, 'parameters', env, 'level0', 'chronics', name)
    ours = Chronic.from_folder(folder)
    ref = ref_chronic_cls(folder)
    entries = ref.timesteps_entries
    assert ours.n_rows == len(entries)
    ids = np.array([e.get_id() for e in entries])
    assert np.array_equal(ours.ids, ids)
    for ours_t, getter in (('prods_p', 'get_prods_p'), ('prods_v', 'get_prods_v'), ('loads_p', 'get_loads_p'),
                           ('loads_q', 'get_loads_q'), ('prods_p_planned', 'get_planned_prods_p'),
                           ('prods_v_planned', 'get_planned_prods_v'), ('loads_p_planned', 'get_planned_loads_p'),
                           ('loads_q_planned', 'get_planned_loads_q'), ('maintenance', 'get_maintenance'),
                           ('hazards', 'get_hazards')):
        expect = np.array([getattr(e, getter)() for e in entries])
        got = getattr(ours, ours_t)
        assert got.dtype == np.float32 and expect.dtype == np.float32, ours_t      # chronic.py:175
        assert np.array_equal(got, expect), ours_t
    dts = np.array([[d.year, d.month, d.day, d.hour, d.minute, d.second] for d in (e.get_datetime() for e in entries)])
    assert np.array_equal(ours.datetimes, dts)
    assert np.array_equal(ours.imaps, np.asarray(ref.get_imaps(), dtype=np.float64))
    # timesteps before planned maintenance (chronic.py:239-246), horizon 20 as in the shipped configurations
    pm = ours.planned_maintenance_table(20)
    for t in (0, 1, len(entries) // 2, len(entries) - 1):
        assert np.array_equal(pm[t], np.asarray(ref.get_planned_maintenance(ids[t], 20))), t


def test_default118_has_id_gaps_and_chronic_set_is_alphabetical():
    from pypownet_b200.chronic import ChronicSet
    cs = ChronicSet.from_folder(os.path.join(REF, 'parameters', 'default118', 'level0', 'chronics'))
    assert [c.name for c in cs.chronics] == sorted(c.name for c in cs.chronics) and len(cs) == 12
    a = cs[0]
    assert a.ids[-1] > a.n_rows - 1                      # ids run past the number of rows: there are gaps
    assert a.row_after_switch == 1                       # id 0 is the first row: play resumes at the second one
