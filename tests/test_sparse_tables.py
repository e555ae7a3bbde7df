"""Host-side check of the sparse LDL^T tables behind the CUDA solvers (no GPU needed): ppn_sparse_selfcheck builds the
tables of a grid exactly as ppn_create does and replays the kernels' factorisation and solves on the host against a
dense Gaussian elimination, on random positive definite matrices over random topologies (lines off, inactive buses,
split buses)."""
import ctypes as C

import numpy as np
import pytest

from pypownet_b200.case import Case


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from pypownet_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize('grid', ['case14', 'case30', 'case118'])
@pytest.mark.parametrize('full', [0, 1])
def test_tables_reproduce_a_dense_solve(lib, grid, full):
    case = Case.builtin(grid)
    lor = np.ascontiguousarray(case.line_or_sub, dtype=np.int32)
    lex = np.ascontiguousarray(case.line_ex_sub, dtype=np.int32)
    info = np.zeros(5, dtype=np.int32)
    worst = 0.0
    for seed in range(12):
        err = C.c_double(-1.0)
        rc = lib.ppn_sparse_selfcheck(case.n_sub, case.n_line, lor.ctypes.data_as(C.c_void_p), lex.ctypes.data_as(C.c_void_p),
                                      full, seed, C.byref(err), info.ctypes.data_as(C.c_void_p))
        assert rc == 0, lib.ppn_last_error(None)
        worst = max(worst, err.value)
    n, nnz, n_lev, cut_lev, nt = info.tolist()
    assert n == case.n_sub * (2 if full else 1)
    assert 0 < nt <= 40 and 0 <= cut_lev < n_lev and nnz < 6 * n      # narrow dense block, little fill
    assert worst < 1e-9, worst


def test_ieee118_structure_is_the_documented_one(lib):
    """265 off-diagonal entries of L and 16 levels for the un-split IEEE-118 grid."""
    case = Case.builtin('case118')
    lor = np.ascontiguousarray(case.line_or_sub, dtype=np.int32)
    lex = np.ascontiguousarray(case.line_ex_sub, dtype=np.int32)
    info = np.zeros(5, dtype=np.int32)
    err = C.c_double()
    assert lib.ppn_sparse_selfcheck(118, case.n_line, lor.ctypes.data_as(C.c_void_p), lex.ctypes.data_as(C.c_void_p), 0, 1,
                                    C.byref(err), info.ctypes.data_as(C.c_void_p)) == 0
    assert info.tolist()[:3] == [118, 265, 16]


def test_tables_of_random_grids(lib):
    """Any grid, not only the IEEE families: random connected line graphs (a spanning tree plus extra and parallel
    lines), 5 to 120 substations, both structures."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        S = int(rng.integers(5, 121))
        lor = [int(rng.integers(0, k)) for k in range(1, S)]          # spanning tree: k attaches to an earlier node
        lex = list(range(1, S))
        for _ in range(int(rng.integers(0, S))):                      # extra lines, some parallel to existing ones
            a, b = rng.integers(0, S, size=2)
            if a != b:
                lor.append(int(a)); lex.append(int(b))
        lor = np.ascontiguousarray(lor, dtype=np.int32)
        lex = np.ascontiguousarray(lex, dtype=np.int32)
        info = np.zeros(5, dtype=np.int32)
        for full in (0, 1):
            err = C.c_double(-1.0)
            rc = lib.ppn_sparse_selfcheck(S, len(lor), lor.ctypes.data_as(C.c_void_p), lex.ctypes.data_as(C.c_void_p),
                                          full, trial, C.byref(err), info.ctypes.data_as(C.c_void_p))
            assert rc == 0, (trial, S, lib.ppn_last_error(None))
            assert err.value < 1e-8, (trial, S, full, err.value, info.tolist())
            assert 0 < info[4] <= 40 or info[2] == 1, info.tolist()
