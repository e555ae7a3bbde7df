"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/pypownet_b200.h declares; the product package has no CPU path and never touches oracle/."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from pypownet_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from pypownet_b200 import _lib
    with open(os.path.join(ROOT, 'include', 'pypownet_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(ppn_[a-z_0-9]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SYMBOLS), 'binding table and header disagree: %s' % (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None
    assert b'sm_100a' in lib.ppn_build_info()


def test_struct_layouts_match_header(lib):
    """Field order of the ctypes structures follows the header's struct declarations."""
    from pypownet_b200 import _lib
    with open(os.path.join(ROOT, 'include', 'pypownet_b200.h')) as f:
        header = f.read()
    for cname, cls in (('ppn_case', _lib.PpnCase), ('ppn_config', _lib.PpnConfig), ('ppn_chronic', _lib.PpnChronic)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), header, re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                names.append(re.findall(r'[A-Za-z_0-9]+', part)[-1])
        assert names == [f[0] for f in cls._fields_], cname


def test_create_rejects_bad_input_without_gpu(lib):
    import ctypes as C
    from pypownet_b200 import _lib
    out = C.c_void_p()
    assert lib.ppn_create(None, None, 4, 0, C.byref(out)) == -1
    assert b'null' in lib.ppn_last_error(None)


def test_create_validates_the_grid_before_touching_the_gpu(lib):
    """Malformed grids are refused with a message (environment.py raises ValueError on malformed input, never inside a
    step): checked before the first CUDA call, so it runs without a GPU."""
    import copy
    import ctypes as C
    import numpy as np
    from golden_util import Fixture
    from pypownet_b200 import _lib
    fx = Fixture('d14_tests_basic')
    cf = _lib.config_struct(fx.config, 'soft', False, 'natural', float(fx.case.n_sub), 0, 0)
    out = C.c_void_p()
    for attr, index, value, needle in (('line_ex_sub', 3, None, b'same substation'), ('line_x', 5, 0.0, b'zero reactance'),
                                       ('line_or_sub', 0, 99, b'out of range')):
        case = copy.copy(fx.case)
        a = np.array(getattr(case, attr)).copy()
        a[index] = np.array(case.line_or_sub)[index] if value is None else value
        setattr(case, attr, a)
        cs, keep = _lib.case_struct(case, fx.thermal_limits)
        assert lib.ppn_create(C.byref(cs), C.byref(cf), 4, 0, C.byref(out)) == -1
        assert needle in lib.ppn_last_error(None), lib.ppn_last_error(None)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'pypownet_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn
                assert 'oracle/' not in src.replace('the oracle/', ''), fn


def test_vec_env_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from pypownet_b200.vec_env import VecRunEnv, PpnError
    from golden_util import Fixture
    fx = Fixture('d14_tests_basic')
    with pytest.raises(PpnError):
        VecRunEnv(fx.case, fx.config, fx.chronics, 2)
