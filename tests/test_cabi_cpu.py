"""CPU-side checks of the product boundary: the C-ABI library loads, exports every symbol
include/nmpc_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "nmpc_b200.h")).read()
    declared = set(re.findall(r"\b(nmpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.SYMBOLS)
    L = ctypes.CDLL(pkg.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s


def test_default_opts_match_the_oracle_defaults(pkg):
    from nmpc_b200._cabi import Opts, default_opts
    from oracle.oracle_lib import Opts as OOpts, lib as olib
    a, b = default_opts(), OOpts()
    olib().orc_default_opts(ctypes.byref(b))
    assert [f[0] for f in Opts._fields_] == [f[0] for f in OOpts._fields_]
    for name, _ in Opts._fields_:
        assert getattr(a, name) == getattr(b, name), name
    assert a.max_iter == 2000 and a.acceptable_tol == 1e-8 and a.acceptable_obj_change_tol == 1e-6   # six...py:345


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.NmpcError):
        pkg.Problem(2, 5, 0.1)
    with pytest.raises(pkg.NmpcError):
        pkg.Problem(1, 5, 0.1, obstacles=[[0.5, 0.5, 0.3]])      # static-obstacle family
    with pytest.raises(pkg.NmpcError):
        pkg.Problem(16, 5, 0.1)                                   # dense-block path
    with pytest.raises(pkg.NmpcError):
        pkg.SmallOcp("van_der_pol")                               # small-OCP family


def test_casadi_helpers_are_column_major(pkg):
    a = np.arange(6.0).reshape(2, 3)
    r = pkg.reshape(a, 3, 2).full()
    np.testing.assert_array_equal(r, a.reshape(3, 2, order="F"))
    d = pkg.DM(np.arange(6.0))
    np.testing.assert_array_equal(d[2:4].full(), [[2.0], [3.0]])
    assert pkg.repmat(np.array([[1.0, 2.0]]), 3, 1).shape == (3, 2)
    assert pkg.vertcat(np.zeros(3), np.ones(2)).shape == (5, 1)
    assert pkg.horzcat(np.zeros((1, 3)), np.ones((1, 2))).shape == (1, 5)
