"""What the UNMODIFIED reference (oracle/_ref) leaves behind at BASELINE's full single-GPU size: the projection of the bench workload
(512^3 float smoke plume, PcNone, cgAccuracy 1e-4) run on the host cores, then the divergence of the projected field.  The recursive CG
residual meets the tolerance; the TRUE divergence b - A x drifts away from it over ~1600 float iterations, so it is the reference's own
post-projection divergence -- not 2 x cgAccuracy -- that bounds what tests/test_gpu_step_properties_fullsize.py may ask of the CUDA path.
Writes tests/golden/fullsize_divergence.json.   usage: python tools/ref_fullsize_divergence.py [res] [kind]     (~10-20 min at 512)"""
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
t0 = time.time()
rhs, _, _ = O.compute_rhs(flags, vel)
A = O.make_matrix(flags)
p, it, rn = O.cg_solve(flags, rhs, *A, pc=0, accuracy=acc, maxIter=int(np.float32(99) * res))
v = vel.copy()
O.correct_velocity(flags, v, p)
div, _, _ = O.compute_rhs(flags, v)
d = np.abs(div[fluid].astype(np.float64))
out = dict(res=res, kind=kind, preconditioner=0, cgAccuracy=acc, iterations=int(it), resNorm=float(rn), max_div=float(d.max()),
           cells_over_2acc=int((d > 2 * acc).sum()), cells_over_acc=int((d > acc).sum()), fluid_cells=int(fluid.sum()),
           pressure_sha1=__import__("hashlib").sha1(np.ascontiguousarray(p).tobytes()).hexdigest(),
           pressure_l2=float(np.linalg.norm(p.astype(np.float64))), seconds=time.time() - t0)
print(json.dumps(out))
path = os.path.join(ROOT, "tests", "golden", "fullsize_divergence.json")
prev = json.load(open(path)) if os.path.exists(path) else {}
prev["%s_%d" % (kind, res)] = out
json.dump(prev, open(path, "w"), indent=1, sort_keys=True)
