"""Host-side logic of bench.py and of the env-start helpers (CPU only: no kernel is launched here)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pypownet_b200 import sharding  # noqa: E402
from pypownet_b200.case import Case  # noqa: E402


@pytest.mark.parametrize('grid,expected', [('case14', 5348), ('case30', 10423), ('case118', 48013)])
def test_algorithmic_bytes_are_the_figures_design_md_quotes(grid, expected):
    """SURVEY.md 8(d) / DESIGN.md 3.3: the bytes one env-step must move, the numerator of roofline.achieved."""
    assert bench.algorithmic_bytes(Case.builtin(grid)) == expected


@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_spread_starts_are_valid_and_cover_the_data_set_on_every_rank(world):
    """Every rank's batch samples all chronics and rows spread over the whole chronic (weak scaling assumes statistically
    identical per-GPU work); rows stay inside the chronic; rank 0 / env 0 starts where the reference starts."""
    B = 4096
    for rank in range(world):
        c, r = bench.shard_starts(B, rank, world)
        assert c.dtype == np.int32 and r.dtype == np.int32 and len(c) == B == len(r)
        assert c.min() == 0 and c.max() == bench.N_CHRONICS - 1
        assert np.all(np.bincount(c, minlength=bench.N_CHRONICS) >= B // bench.N_CHRONICS)
        assert r.min() >= 0 and r.max() < bench.N_ROWS - 1
        for k in range(bench.N_CHRONICS):                       # rows of one chronic: evenly spread, span > 90 % of it
            rk = np.sort(r[c == k])
            assert rk[-1] - rk[0] > 0.9 * (bench.N_ROWS - 1)
            assert np.max(np.diff(rk)) <= 6                     # (the rank offset wraps around the end of the chronic)
    c0, r0 = bench.shard_starts(B, 0, world)
    assert c0[0] == 0 and r0[0] == 0
    if world > 1:                                               # different ranks, different rows of the same chronics
        c1, r1 = bench.shard_starts(B, 1, world)
        assert np.array_equal(c0, c1) and not np.array_equal(r0, r1)


def test_block_and_strided_shards_partition_one_global_batch():
    """The two analysis modes cut ONE global batch (env e -> chronic e mod 12, row e // 12): together the ranks hold
    every env of it exactly once."""
    B, world = 512, 4
    want_c, want_r = sharding.env_starts_of(bench.N_CHRONICS, bench.N_ROWS, np.arange(B * world))
    want = set(zip(want_c.tolist(), want_r.tolist(), range(B * world)))
    for mode in ('blocks', 'strided'):
        got = []
        for rank in range(world):
            c, r = bench.shard_starts(B, rank, world, mode)
            ids = rank * B + np.arange(B) if mode == 'blocks' else rank + world * np.arange(B)
            got += list(zip(c.tolist(), r.tolist(), ids.tolist()))
        assert set(got) == want and len(got) == B * world


def test_random_action_bank_is_one_substation_and_one_line_per_env():
    """agent.py:78-158 (RandomNodeSplitting + RandomLineSwitch): node bits only inside ONE substation, exactly one line
    switch; different batches differ."""
    case = Case.builtin('case118')
    bank = bench.random_action_bank(case, 64, n_batches=3, seed=7)
    assert bank.shape == (3, 64, case.action_length) and bank.dtype == np.uint8
    elem_sub = np.asarray(case.elem_sub)
    nt = len(elem_sub)
    for k in range(3):
        for e in range(64):
            subs = np.unique(elem_sub[bank[k, e, :nt] != 0])
            assert len(subs) <= 1
            assert bank[k, e, nt:].sum() == 1
    assert not np.array_equal(bank[0], bank[1])


def test_clock_sampler_without_nvidia_smi_reports_it(monkeypatch):
    """On a box without nvidia-smi the clocks block says so instead of failing the bench."""
    monkeypatch.setenv('PATH', '/nonexistent')
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert out['sm_mhz'] is None and out['reasons'] == ['nvidia-smi unavailable']


def test_reference_arm_describes_what_it_ran():
    """cpu_pool picks the unmodified reference when baseline/_ref is installed and the numpy port otherwise; the label
    in the JSON line (`kind`) follows."""
    assert isinstance(bench.reference_available(), bool)
    if not bench.reference_available():
        pytest.skip('baseline/_ref not installed here')
    assert os.path.isdir(os.path.join(bench.REF_DIR, 'pypownet'))
