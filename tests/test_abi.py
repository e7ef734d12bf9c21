"""CPU (not gpu): the C-ABI shared library loads without a GPU and exports every symbol include/mantapress.h declares;
host-side logic that needs no device (slab rule, defaults, error strings); the product path fails loudly without CUDA."""
import ctypes as C
import os

import numpy as np
import pytest

import mantaflow_b200 as mf
from mantaflow_b200 import _lib, sharded


def test_library_exports_every_declared_symbol():
    lib = mf.load()
    names = mf.declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mp_version() == 100


def test_defaults_match_reference_signature():
    p = mf.PressureParams()
    mf.load().mp_pressure_params_default(C.byref(p))
    # pressure.cpp:481-494
    assert (p.cgAccuracy, p.gfClamp, p.cgMaxIterFac, p.precondition, p.preconditioner) == (1e-3, 1e-4, 1.5, 1, mf.PcMIC)
    assert (p.enforceCompatibility, p.useL2Norm, p.zeroPressureFixing, p.surfTens) == (0, 0, 0, 0.0)
    assert (mf.PcNone, mf.PcMIC, mf.PcMGDynamic, mf.PcMGStatic) == (0, 1, 2, 3)        # python/defines.py:46-50
    import inspect
    sig = inspect.signature(mf.solvePressure)
    assert list(sig.parameters) == ["vel", "pressure", "flags", "cgAccuracy", "phi", "perCellCorr", "fractions", "obvel", "gfClamp",
                                    "cgMaxIterFac", "precondition", "preconditioner", "enforceCompatibility", "useL2Norm",
                                    "zeroPressureFixing", "curv", "surfTens", "retRhs"]          # pressure.cpp:480-495
    assert list(inspect.signature(mf.correctVelocity).parameters)[:4] == ["vel", "pressure", "flags", "cgAccuracy"]
    assert list(inspect.signature(mf.computePressureRhs).parameters)[:4] == ["rhs", "vel", "pressure", "flags"]


def test_status_strings_and_slab_rule_without_gpu():
    lib = mf.load()
    assert lib.mp_status_string(0) == b"MP_OK" and lib.mp_status_string(3) == b"MP_ERR_DIVERGED"
    for sz, world in ((512, 8), (41, 2), (100, 3), (7, 2)):
        covered = []
        for r in range(world):
            k0, k1 = C.c_int(0), C.c_int(0)
            assert lib.mp_dist_slab(sz, r, world, C.byref(k0), C.byref(k1)) == 0
            assert (k0.value, k1.value) == sharded.slab(sz, r, world)
            covered += list(range(k0.value, k1.value))
        assert covered == list(range(sz))
    assert lib.mp_dist_slab(10, 3, 2, C.byref(C.c_int()), C.byref(C.c_int())) == _lib.MP_ERR_INVALID


def test_slab_scatter_gather_roundtrip():
    a = np.arange(11 * 3 * 2, dtype=np.float32).reshape(11, 3, 2)
    for world in (1, 2, 3):
        parts = [sharded.local_slab(a, r, world) for r in range(world)]
        assert np.array_equal(sharded.assemble(sharded.owned(p) for p in parts), a)
        for r, p in enumerate(parts):
            k0, k1 = sharded.slab(11, r, world)
            assert p.shape[0] == k1 - k0 + 2
            assert np.array_equal(p[0], a[k0 - 1] if k0 > 0 else 0 * a[0]) and np.array_equal(p[-1], a[k1] if k1 < 11 else 0 * a[0])


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (MP_ERR_CUDA), never compute on the CPU."""
    if mf.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mf.MantaError, match="no CUDA device"):
        mf.Solver(gridSize=(8, 8, 8), dim=3)


def test_product_package_never_imports_the_oracle():
    import re
    root = os.path.dirname(os.path.abspath(mf.__file__))
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|oracle_api|libmf_oracle|libmanta_ref|mf_oracle\.c", re.M)
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_python_mirror_passes_as_many_arguments_as_the_header_declares():
    """ctypes does not check argument counts; the C-ABI calls of the Python mirror (and of the test adapter) are compared with the declarations in
    include/mantapress.h.  Calls with a starred argument are counted with the three values a vector constant expands to."""
    import ast
    import glob
    import re
    from mantaflow_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "mantapress.h")).read(), flags=re.S)
    declared = {}
    for m in re.finditer(r"\b(mp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        declared[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    assert set(declared) >= set(_lib.declared_symbols())
    checked, bad = 0, []
    for f in glob.glob(os.path.join(root, "mantaflow_b200", "*.py")) + [os.path.join(root, "tests", "cuda_impl.py")]:
        for node in ast.walk(ast.parse(open(f).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in declared:
                n = sum(3 if isinstance(a, ast.Starred) and "xyz" in ast.dump(a) or isinstance(a, ast.Starred) and "_vec3" in ast.dump(a) else (None if isinstance(a, ast.Starred) else 1)
                        for a in node.args) if not any(isinstance(a, ast.Starred) and "xyz" not in ast.dump(a) and "_vec3" not in ast.dump(a) for a in node.args) else None
                if n is None:
                    continue
                checked += 1
                if n != declared[node.func.attr]:
                    bad.append((os.path.basename(f), node.lineno, node.func.attr, n, declared[node.func.attr]))
    assert not bad, bad
    assert checked > 80
