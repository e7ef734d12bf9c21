"""Builds libmantapress.so (sm_100a only) in-tree with nvcc.  Called by __graft_entry__.build().
Every .cu is compiled to its own object (in parallel, only when it or a header changed), then linked."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
OUT = os.path.join(HERE, "libmantapress.so")
SOURCES = ["mp_api.cu", "mp_assembly.cu", "mp_cg.cu", "mp_mic.cu", "mp_micrb.cu", "mp_ic.cu", "mp_mg.cu", "mp_plugin.cu", "mp_step.cu", "mp_liquid.cu", "mp_particles.cu", "mp_guiding.cu", "mp_dist.cu"]
# -fmad=false: the reference build has no FMA; assembly and MIC kernels must be bit-exact.
# IEEE div/sqrt and no flush-to-zero are nvcc's defaults without --use_fast_math.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))] + \
           [os.path.join(HERE, "..", "include", "mantapress.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(OUT, [os.path.join(CSRC, s) for s in SOURCES] + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        list(ex.map(run, jobs))
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", OUT] +
        [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES] + ["-ldl"])
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
