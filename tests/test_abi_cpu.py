"""
CPU-only checks of the C-ABI library: it loads, exports every symbol include/differt2d_b200.h
declares, validates arguments, and its HOST candidate enumerator matches the oracle.  No compute
kernel is launched here (no GPU in this tier).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from differt2d_b200 import _lib as L
from differt2d_b200 import functional as F
from oracle import c_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "differt2d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d2d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/differt2d_b200.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.d2d_abi_version() == 3


def test_struct_layout_matches_header_order():
    src = open(os.path.join(ROOT, "include", "differt2d_b200.h")).read()
    body = src[src.index("typedef struct D2DProblem {"):src.index("} D2DProblem;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip().split("{")[-1].strip()
        if not decl:
            continue
        names = [re.sub(r"[\*\s]", "", x.split()[-1]) for x in decl.split(",")]
        fields += names
    assert fields == [f[0] for f in L.D2DProblem._fields_]


def test_defaults_are_the_references():
    p = L.new_problem()
    assert (p.min_order, p.max_order) == (0, 1)           # scene.py:1816-1817
    assert p.alpha == 100.0 and p.patch == 0.0             # defaults.py:3,7
    assert p.r_coef == 0.5 and p.height == 0.1             # defaults.py:12,15
    assert abs(p.tol - 1e-2) < 1e-9 and p.steps == 100 and abs(p.lr - 0.1) < 1e-8
    assert p.mode == L.MODE_HARD and p.method == L.METHOD_IMAGE and p.fun == L.FUN_RECEIVED_POWER


@pytest.mark.parametrize("n,k,f", [(8, 0, ()), (8, 1, ()), (8, 2, ()), (28, 2, ()), (6, 2, (0, 1, 2, 4, 5)),
                                    (5, 3, (2,)), (7, 4, ()), (3, 1, (0, 1, 2)), (2, 3, ()), (1, 2, ()), (0, 1, ()),
                                    (60, 3, (3, 9, 59))])
def test_host_candidates_match_oracle(n, k, f):
    got = F.candidates(n, k, f)
    want = CO.candidates(n, k, list(f))
    assert got.dtype == np.int32 and got.shape == want.shape and np.array_equal(got, want)
    assert L.lib().d2d_candidates_count(n, k, None, 0) == len(CO.candidates(n, k))


def test_candidate_counts_closed_form():
    for n in (1, 2, 8, 28, 500):
        for k in range(0, 4):
            want = 1 if k == 0 else n * (n - 1) ** (k - 1)
            assert L.lib().d2d_candidates_count(n, k, None, 0) == want
    assert L.lib().d2d_candidates_count(500, 3, None, 0) == 124_500_500
    assert L.lib().d2d_candidates_count(-1, 1, None, 0) == -1
    assert L.lib().d2d_candidates_count(L.MAX_OBJECTS + 1, 1, None, 0) == -1


def test_argument_validation_without_gpu():
    lib = L.lib()
    assert lib.d2d_power_fwd(None, None, None, None) == 1
    assert b"NULL" in lib.d2d_last_error()
    p = L.new_problem()
    p.max_order = L.MAX_ORDER + 1
    assert lib.d2d_problem_num_candidates(C.byref(p)) == -1
    p = L.new_problem()
    p.min_order, p.max_order = 2, 1
    assert lib.d2d_power_fwd(C.byref(p), None, None, None) == 1
    p = L.new_problem()
    p.mode, p.alpha = L.MODE_SIGMOID, -1.0
    assert lib.d2d_power_fwd(C.byref(p), None, None, None) == 1
    assert b"alpha" in lib.d2d_last_error()
    p = L.new_problem()
    p.n_objects = 28
    xys = np.zeros((28, 2, 2), np.float32)
    p.objects_xys = xys.ctypes.data
    p.max_order = 2
    assert lib.d2d_problem_num_candidates(C.byref(p)) == 785
    p.method = L.METHOD_FERMAT
    # Fermat/MinPath reverse mode needs the x0 table, like the forward
    assert lib.d2d_power_bwd(C.byref(p), None, None, None, None, None, None, None, None) == 1
    assert b"x0" in lib.d2d_last_error()
    # gradient semantics: clean (default) or nan_parity; the latter covers ImagePath only
    p.grad_mode = L.GRAD_NAN_PARITY
    assert lib.d2d_problem_num_candidates(C.byref(p)) == -1 and b"ImagePath" in lib.d2d_last_error()
    p.method = L.METHOD_IMAGE
    assert lib.d2d_problem_num_candidates(C.byref(p)) == 785
    p.grad_mode = 7
    assert lib.d2d_problem_num_candidates(C.byref(p)) == -1 and b"grad_mode" in lib.d2d_last_error()


def test_no_cpu_fallback():
    """The product must fail loudly off-GPU instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import differt2d_b200 as d

    sc = d.Scene.square_scene()
    X, Y = sc.grid(4)
    with pytest.raises(d.D2DError):
        list(sc.accumulate_on_receivers_grid_over_paths(X, Y, approx=False))
    with pytest.raises(d.D2DError):
        F.power_fwd(F.TraceConfig(), np.zeros((1, 2, 2), np.float32), np.zeros((1, 2), np.float32),
                    np.zeros((4, 2), np.float32), device="cpu")
    with pytest.raises(d.D2DError):  # the host-buffer entry stages to a CUDA device: no device, no result
        F.power_host(F.TraceConfig(), np.zeros((1, 2, 2), np.float32), np.zeros((1, 2), np.float32),
                     np.zeros((4, 2), np.float32))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "differt2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("oracle-free", ""), f"{f} mentions the oracle"
