"""ctypes binding of the C-ABI (include/mantapress.h).  The product path: no CPU fallback, no oracle.

Loading the library needs no GPU (cudart is linked statically and only touches the driver on the first
call), so `python -m pytest -m "not gpu"` can check that every declared symbol is exported."""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmantapress.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "mantapress.h")

MP_OK = 0
MP_ERR_INVALID, MP_ERR_CUDA, MP_ERR_DIVERGED, MP_ERR_NOT_SET, MP_ERR_UNSUPPORTED, MP_ERR_COMM = 1, 2, 3, 4, 5, 6
MP_GRID_REAL, MP_GRID_FLAGS, MP_GRID_MAC = 1, 2, 8
PcNone, PcMIC, PcMGDynamic, PcMGStatic = 0, 1, 2, 3          # python/defines.py:46-50
MP_CG_PC_NONE, MP_CG_PC_ICP, MP_CG_PC_MICP, MP_CG_PC_MGP = 0, 1, 2, 3


class MantaError(RuntimeError):
    """What the reference raises as Manta::Error -> RuntimeError (general.h:42-57, pclass.cpp:57-61)."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


class PressureParams(C.Structure):
    _fields_ = [("cgAccuracy", C.c_double), ("gfClamp", C.c_double), ("cgMaxIterFac", C.c_double),
                ("precondition", C.c_int), ("preconditioner", C.c_int), ("enforceCompatibility", C.c_int),
                ("useL2Norm", C.c_int), ("zeroPressureFixing", C.c_int), ("surfTens", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("resNorm", C.c_double), ("maxIter", C.c_int), ("fixedCell", C.c_longlong),
                ("mgLevels", C.c_int), ("msRhs", C.c_float), ("msMatrix", C.c_float), ("msSolve", C.c_float),
                ("msCorrect", C.c_float), ("msTotal", C.c_float), ("msH2D", C.c_float), ("msD2H", C.c_float),
                ("msMatvecAvg", C.c_float), ("msAxpyAvg", C.c_float), ("msUpdateAvg", C.c_float), ("msPrecondAvg", C.c_float),
                ("profSamples", C.c_int), ("matvecKernel", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def declared_symbols():
    """Every function name include/mantapress.h declares."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z0-9_]+)\s*\(", txt)))


def load():
    """Load libmantapress.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("mantaflow_b200: %s is missing -- build it with `python -m mantaflow_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.mp_last_error.restype = C.c_char_p
    lib.mp_status_string.restype = C.c_char_p
    lib.mp_context_stream.restype = C.c_void_p
    lib.mp_grid_device_ptr.restype = C.c_void_p
    lib.mp_pressure_params_default.restype = None
    _lib = lib
    return lib


def check(status):
    if status != MP_OK:
        raise MantaError(status, load().mp_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int(0)
    load().mp_device_count(C.byref(n))
    return n.value
