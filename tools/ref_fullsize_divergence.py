"""What the UNMODIFIED reference (oracle/_ref) leaves behind at BASELINE's full single-GPU size: the projection of the bench workload
(512^3 float smoke plume, PcNone, cgAccuracy 1e-4) run on the host cores, then the divergence of the projected field.  The recursive CG
residual meets the tolerance; the TRUE divergence b - A x drifts away from it over ~1600 float iterations, so it is the reference's own
post-projection divergence -- not 2 x cgAccuracy -- that bounds what tests/test_gpu_step_properties_fullsize.py may ask of the CUDA path.
Writes tests/golden/fullsize_divergence.json.   usage: python tools/ref_fullsize_divergence.py [res] [kind] [pcs, e.g. 0,3]     (~16 min at 512 for PcNone)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from mantaflow_b200 import scenes  # noqa: E402
from oracle.oracle_api import Oracle  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kind = sys.argv[2] if len(sys.argv) > 2 else "reference"
acc = 1e-4
O = Oracle(kind, 4)
flags, vel = scenes.smoke_plume(res, 4, random_vel=False)
fluid = (flags & 1) != 0
import hashlib


def stats(pc, p, it, rn, v, t0):
    div, _, _ = O.compute_rhs(flags, v)
    d = np.abs(div[fluid].astype(np.float64))
    return dict(res=res, kind=kind, preconditioner=pc, cgAccuracy=acc, iterations=int(it), resNorm=float(rn), max_div=float(d.max()),
                cells_over_2acc=int((d > 2 * acc).sum()), cells_over_acc=int((d > acc).sum()), fluid_cells=int(fluid.sum()),
                pressure_sha1=hashlib.sha1(np.ascontiguousarray(p).tobytes()).hexdigest(),
                pressure_sha1_poszero=hashlib.sha1(np.ascontiguousarray(p + 0.0).tobytes()).hexdigest(),      # -0 -> +0
                pressure_l2=float(np.linalg.norm(p.astype(np.float64))), seconds=time.time() - t0)


path = os.path.join(ROOT, "tests", "golden", "fullsize_divergence.json")
pcs = [int(q) for q in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 3]
for pc in pcs:
    t0 = time.time()
    v = vel.copy()
    if pc == 0:      # the reference plugin asserts on PcNone in 3-D: rhs + matrix + GridCg + correctVelocity driven directly
        rhs, _, _ = O.compute_rhs(flags, vel)
        A = O.make_matrix(flags)
        p, it, rn = O.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=int(np.float32(99) * res))
        O.correct_velocity(flags, v, p)
        del A, rhs
    else:
        p, it, rn = O.solve_pressure(flags, v, cgAccuracy=acc, cgMaxIterFac=99, preconditioner=pc, zeroPressureFixing=(pc >= 2))
    out = stats(pc, p, it, rn, v, t0)
    print(json.dumps(out), flush=True)
    prev = json.load(open(path)) if os.path.exists(path) else {}
    prev["%s_%d%s" % (kind, res, "" if pc == 0 else "_pc%d" % pc)] = out
    json.dump(prev, open(path, "w"), indent=1, sort_keys=True)
