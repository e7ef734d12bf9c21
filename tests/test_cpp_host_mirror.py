"""CPU (not gpu): the C++ host mirror include/mantapress.hpp (the reference's FluidSolver / Grid<T> / MACGrid / FlagGrid classes and
plugin signatures over the C-ABI) compiles warning-free as C++14 in both precisions, links against libmantapress.so with every plugin
it declares, and without a CUDA device fails loudly (Manta::Error, no CPU fallback).  The parity run of the same program is
tests/test_gpu_zz_cpp_host_mirror.py."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cpp", "host_mirror_test.cpp")
LIBDIR = os.path.join(ROOT, "mantaflow_b200")


def build_host_mirror_test(out, defines=()):
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-Werror"] + ["-D" + d for d in defines] + \
          ["-o", out, SRC, "-L" + LIBDIR, "-lmantapress", "-ldl", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return out


def test_cpp_mirror_builds_and_fails_loudly_without_a_device(tmp_path):
    import mantaflow_b200 as mf
    exe = build_host_mirror_test(str(tmp_path / "host_mirror_test"))
    if mf.device_count() > 0:
        pytest.skip("a CUDA device is present: the parity run covers this machine")
    r = subprocess.run([exe, "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout and "plugins link" in r.stdout


def test_cpp_mirror_compiles_in_double_precision(tmp_path):
    build_host_mirror_test(str(tmp_path / "host_mirror_test_f64"), defines=["DOUBLEPRECISION=1", "MIRROR_COMPILE_ONLY=1"])


def test_cpp_mirror_keeps_the_reference_signatures():
    """parameter names, order and defaults of the C++ plugins are the reference's (same fixture as tests/test_api_signatures.py)"""
    import json
    sigs = json.load(open(os.path.join(HERE, "golden", "plugin_signatures.json")))
    txt = open(os.path.join(ROOT, "include", "mantapress.hpp")).read()
    txt = re.sub(r"//[^\n]*", "", txt)
    checked = 0
    for name, params in sigs.items():
        m = re.search(r"inline void %s\s*\(" % re.escape(name), txt)
        if not m:
            continue                      # PD_fluid_guiding is reachable through the C-ABI / Python mirror only
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        body = " ".join(txt[m.end():i - 1].split())
        got = []
        for p in [q.strip() for q in body.split(",")]:
            default = None
            if "=" in p:
                p, default = (t.strip() for t in p.split("=", 1))
            got.append([re.findall(r"[A-Za-z_][A-Za-z_0-9]*", p)[-1], default])
        norm = lambda d: None if d is None else {"NULL": "0", "nullptr": "0"}.get(d, d).rstrip("f")
        assert [g[0] for g in got] == [p[0] for p in params], (name, got, params)
        assert [norm(g[1]) for g in got] == [norm(p[1]) for p in params], (name, got, params)
        checked += 1
    assert checked >= 14
