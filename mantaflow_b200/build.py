"""Builds libmantapress.so (sm_100a only) in-tree with nvcc.  Called by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmantapress.so")
SOURCES = ["mp_api.cu", "mp_assembly.cu", "mp_cg.cu", "mp_mic.cu", "mp_ic.cu", "mp_mg.cu", "mp_plugin.cu", "mp_step.cu", "mp_liquid.cu", "mp_particles.cu", "mp_guiding.cu", "mp_dist.cu"]
# -fmad=false: the reference build has no FMA; assembly and MIC kernels must be bit-exact.
# IEEE div/sqrt and no flush-to-zero are nvcc's defaults without --use_fast_math.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mantapress.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
