# Run INSIDE the bound reference binary (tools/build_bound_manta.sh):   SCENE=<scene.py> OUT=<prefix> oracle/_ref/bound/manta tools/bound_scene_driver.py
# Executes the scene file UNMODIFIED (exec of its text in this namespace), then writes what the run left behind -- the grids named
# density / vel / pressure / phi, if the scene has them -- to <prefix>.npz and the wall time to <prefix>.json, so that a run with
# solvePressure on the GPU (default) and one with the reference's own CPU body (MANTA_CPU_PRESSURE=1) can be compared value for value.
import json
import os
import time

import numpy as np
from manta import *   # noqa: F401,F403

_real = np.float64 if globals().get("DOUBLEPRECISION", False) else np.float32      # python/defines.py
_scene, _out = os.environ["SCENE"], os.environ.get("OUT", "bound_run")
setDebugLevel(int(os.environ.get("MANTA_DEBUG", "1")))
_t0 = time.time()
exec(compile(open(_scene).read(), _scene, "exec"))
_wall = time.time() - _t0
_dump = {}
for _name in ("density", "pressure", "phi"):
    if _name in globals() and hasattr(globals()[_name], "getSize"):
        _g = globals()[_name]
        _sz = _g.getSize()
        _a = np.zeros((int(_sz.z), int(_sz.y), int(_sz.x)), _real)
        copyGridToArrayReal(_g, _a)
        _dump[_name] = _a
if "vel" in globals() and hasattr(globals()["vel"], "getSize"):
    _sz = vel.getSize()
    _a = np.zeros((int(_sz.z), int(_sz.y), int(_sz.x), 3), _real)
    copyGridToArrayMAC(vel, _a)
    _dump["vel"] = _a
np.savez(_out + ".npz", **_dump)
json.dump({"scene": _scene, "wall_s": _wall, "cpu_pressure": os.environ.get("MANTA_CPU_PRESSURE", "0"), "grids": sorted(_dump)}, open(_out + ".json", "w"))
print("bound_scene_driver: %s finished in %.2f s, wrote %s.npz" % (_scene, _wall, _out))
