"""Host-side mirror of the reference's FluidSolver / Grid<Real> / MACGrid / FlagGrid types for the pressure path
(grid.h:92-365, fluidsolver.h:27-91) with a device-resident storage mirror: every grid owns a numpy array in the
reference layout ([Z,Y,X] / [Z,Y,X,3], the convention of plugin/numpyconvert.cpp:145-183) AND an mp_grid in HBM.
Two dirty bits keep them coherent lazily, so a sequence of pressure plugins never leaves the device.

Only what the pressure path and its callers need is here; scene construction helpers (initDomain, fillGrid, setConst) are
the simple host loops of grid.cpp:732-861; updateFromLevelset and setBound, which liquid scenes call every step, run on the device."""
import ctypes as C
import weakref
import numpy as np

from . import _lib
from ._lib import MP_GRID_FLAGS, MP_GRID_MAC, MP_GRID_REAL, check

# FlagGrid::CellType grid.h:292-304 / python/defines.py
FlagFluid, FlagObstacle, FlagEmpty, FlagInflow, FlagOutflow, FlagOpen, FlagStick = 1, 2, 4, 8, 16, 32, 64


class Solver:
    """FluidSolver (fluidsolver.h:27): grid size, dimension, precision, and the CUDA context that plays the
    `parent` key of gMapMG (pressure.cpp:250).  `prec` 4 = the reference's float build, 8 = -DDOUBLEPRECISION."""

    def __init__(self, gridSize, dim=3, prec=4, device=0, name="main"):
        gs = tuple(int(v) for v in gridSize)
        if dim not in (2, 3):
            raise _lib.MantaError(1, "Only 2D and 3D solvers allowed.")
        if dim == 2 and gs[2] != 1:
            raise _lib.MantaError(1, "Trying to create 2D solver with size.z != 1")
        self.gridSize, self.dim, self.prec, self.name = gs, dim, prec, name
        self.real = np.float32 if prec == 4 else np.float64
        self.timestep = 1.0
        # time stepping state, fluidsolver.cpp:107-110 (Python names of fluidsolver.h:58-69)
        self.timeTotal, self.frame, self.cfl, self.timestepMin, self.timestepMax, self.frameLength, self.timePerFrame = 0.0, 0, 1000.0, 1.0, 1.0, 1.0, 0.0
        self._lockDt = False
        self.lib = _lib.load()
        self._ctx = C.c_void_p()
        check(self.lib.mp_context_create(C.c_int(device), C.byref(self._ctx)))
        # everything that holds device memory or a handle of this context (grids, particle arrays, GridCg, GridMg): close() releases them
        # BEFORE the context goes, so no handle outlives the mp_context it points into and no device block is leaked
        self._children = weakref.WeakSet()

    def _adopt(self, child):
        self._children.add(child)

    # solver.create(RealGrid) like the reference's Python API
    def create(self, cls, **kw):
        return cls(self, **kw)

    def getGridSize(self):
        return self.gridSize

    def is3D(self):
        return self.dim == 3

    def step(self):
        """FluidSolver::step fluidsolver.cpp:142-158 (the time bookkeeping; there is no GUI to update)"""
        R = self.real
        eps = 1e-6 if self.prec == 4 else 1e-10
        self.timePerFrame = float(R(self.timePerFrame) + R(self.timestep))
        self.timeTotal = float(R(self.timeTotal) + R(self.timestep))
        if self.timePerFrame + eps > self.frameLength:
            self.frame += 1
            self.timeTotal = float(R(float(self.frame) * float(R(self.frameLength))))
            self.timePerFrame = 0.0
            self._lockDt = False

    def adaptTimestep(self, maxVel):
        """FluidSolver::adaptTimestep fluidsolver.cpp:176-196: CFL-limited step, clamped to [timestepMin, timestepMax] and fitted to the frame"""
        R = self.real
        dt, tpf, fl = R(self.timestep), R(self.timePerFrame), R(self.frameLength)
        mvt = R(maxVel) * dt
        if not self._lockDt:
            dt = max(min(dt * R(float(R(self.cfl)) / (float(mvt) + 1e-05)), R(self.timestepMax)), R(self.timestepMin))
            if float(tpf) + float(dt) * 1.05 > float(fl):
                dt = R(float(fl - tpf) + 1e-04)
            elif float(tpf + dt + R(self.timestepMin)) > float(fl) or float(tpf) + float(dt) * 1.25 > float(fl):
                dt = R((float(fl - tpf) + 1e-04) * 0.5)
                self._lockDt = True
        if not float(dt) > self.timestepMin / 2.0:
            raise _lib.MantaError(1, "Invalid dt encountered! Shouldnt happen...")
        self.timestep = float(dt)

    def setMicOrdering(self, mode=0, tileY=0, tileZ=0):
        """the ordering MIC(0) (PcMIC, GridCg PC_mICP, Init/ApplyPreconditionModifiedIncompCholesky2) is formed in: 0 the reference's
        lexicographic one (default, bit-identical to the reference), 1 block red-black with tiles of tileY x tileZ rows -- the reformulated,
        bandwidth-bound preconditioner; more iterations, reported by lastSolveInfo() (mp_set_mic_ordering)"""
        check(self.lib.mp_set_mic_ordering(self._ctx, int(mode), int(tileY), int(tileZ)))

    def micOrdering(self):
        """(mode, tileY, tileZ) of the last MIC(0) factorisation: (1, TY, TZ) in block red-black ordering, (0, 0, 0) in the reference's"""
        m, ty, tz = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.lib.mp_get_mic_ordering(self._ctx, C.byref(m), C.byref(ty), C.byref(tz)))
        return m.value, ty.value, tz.value

    def trim(self):
        """return the context's pooled (unused) device blocks to the driver (mp_context_trim)"""
        check(self.lib.mp_context_trim(self._ctx))

    def synchronize(self):
        check(self.lib.mp_context_synchronize(self._ctx))

    def kernelLaunches(self):
        n = C.c_longlong(0)
        check(self.lib.mp_context_kernel_launches(self._ctx, C.byref(n)))
        return n.value

    def setProfiling(self, period):
        check(self.lib.mp_context_set_profiling(self._ctx, C.c_int(period)))

    def stream(self):
        return self.lib.mp_context_stream(self._ctx)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            kids = list(getattr(self, "_children", ()))
            # solvers before their operands: a GridCg / GridMg refers to grids
            for c in sorted(kids, key=lambda c: 0 if type(c).__name__ in ("GridCg", "GridMg") else 1):
                try:
                    c.close()
                except Exception:
                    pass
            self.lib.mp_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _GridBase:
    KIND = MP_GRID_REAL

    def __init__(self, parent, data=None):
        self.parent = parent
        sx, sy, sz = parent.gridSize
        self.size = (sx, sy, sz)
        shape = (sz, sy, sx) + ((3,) if self.KIND == MP_GRID_MAC else ())
        dtype = np.int32 if self.KIND == MP_GRID_FLAGS else parent.real
        if data is None:
            self._host = np.zeros(shape, dtype)
        else:
            self._host = np.array(data, dtype=dtype, order="C", copy=True).reshape(shape)   # never alias the caller's array
        self._dev = C.c_void_p()
        check(parent.lib.mp_grid_create(parent._ctx, C.c_int(self.KIND), C.c_int(parent.prec), C.c_int(sx), C.c_int(sy), C.c_int(sz), C.byref(self._dev)))
        self._hostDirty = data is not None      # host holds newer data than the device
        parent._adopt(self)
        self._devDirty = False                  # device holds newer data than the host

    # ---- coherence ----
    def dev(self):
        """mp_grid handle with the device copy made current (uploads only if the host copy is newer)."""
        if self._hostDirty:
            check(self.parent.lib.mp_grid_upload(self._dev, self._host.ctypes.data_as(C.c_void_p)))
            self._hostDirty = False
        return self._dev

    def markDeviceWritten(self):
        self._devDirty, self._hostDirty = True, False

    def numpy(self, writable=False):
        """numpy view of the host copy made current (downloads only if the device copy is newer)."""
        if self._devDirty:
            check(self.parent.lib.mp_grid_download(self._dev, self._host.ctypes.data_as(C.c_void_p)))
            self._devDirty = False
        if writable:
            self._hostDirty = True
        return self._host

    def copyFromArray(self, arr):
        self._host[...] = np.asarray(arr, dtype=self._host.dtype).reshape(self._host.shape)
        self._hostDirty, self._devDirty = True, False

    # ---- reference API subset ----
    def getSizeX(self): return self.size[0]
    def getSizeY(self): return self.size[1]
    def getSizeZ(self): return self.size[2]
    def getSize(self): return self.size
    def is3D(self): return self.size[2] > 1

    def clear(self):
        """Grid<T>::clear grid.cpp:93-96: both copies are zeroed, nothing crosses the bus"""
        check(self.parent.lib.mp_grid_clear(self._dev))
        self._host[...] = 0
        self._hostDirty, self._devDirty = False, False

    # ---- element-wise arithmetic on the device, grid.cpp:258-284 (constants: a scalar, or a 3-tuple for Vec3 grids) ----
    _OPS = {"setConst": 0, "addConst": 1, "multConst": 2, "add": 3, "sub": 4, "mult": 5, "addScaled": 6, "clamp": 7, "stomp": 8, "safeDivide": 9}

    def _arith(self, op, other=None, x=0.0, y=0.0, z=0.0):
        check(self.parent.lib.mp_grid_arith(self.parent._ctx, self.dev(), C.c_int(self._OPS[op]), None if other is None else other.dev(),
                                            C.c_double(x), C.c_double(y), C.c_double(z)))
        self.markDeviceWritten()

    @staticmethod
    def _xyz(v):
        try:
            x, y, z = (float(c) for c in v)
        except TypeError:
            x = y = z = float(v)
        return x, y, z

    def setConst(self, value): self._arith("setConst", None, *self._xyz(value))
    def addConst(self, value): self._arith("addConst", None, *self._xyz(value))
    def multConst(self, value): self._arith("multConst", None, *self._xyz(value))
    def add(self, a): self._arith("add", a)
    def sub(self, a): self._arith("sub", a)
    def mult(self, a): self._arith("mult", a)
    def addScaled(self, a, factor): self._arith("addScaled", a, *self._xyz(factor))
    def clamp(self, min, max): self._arith("clamp", None, float(min), float(max), 0.0)
    def stomp(self, threshold): self._arith("stomp", None, *self._xyz(threshold))
    def safeDivide(self, a): self._arith("safeDivide", a)

    def copyFrom(self, other):
        """Grid<T>::copyFrom grid.cpp:205-210: a device-to-device copy when both grids live in the same context"""
        if other.parent is self.parent and other.KIND == self.KIND and other.size == self.size:
            check(self.parent.lib.mp_grid_copy_from(self._dev, other.dev()))
            self.markDeviceWritten()
        else:
            self.copyFromArray(other.numpy())

    def save(self, name):
        """Grid<T>::save grid.cpp:134-156 (.uni, .raw, .npz) from the device-resident grid"""
        from . import fileio
        return fileio.save(self, name)

    def load(self, name):
        """Grid<T>::load grid.cpp:112-132"""
        from . import fileio
        return fileio.load(self, name)

    def setBound(self, value, boundaryWidth=1):
        """Grid<T>::setBound grid.cpp:585-593, on the device (Vec3 grids take a 3-tuple or one value for every component)"""
        try:
            vx, vy, vz = (float(c) for c in value)
        except TypeError:
            vx = vy = vz = float(value)
        check(self.parent.lib.mp_grid_set_bound(self.parent._ctx, self.dev(), C.c_double(vx), C.c_double(vy), C.c_double(vz), C.c_int(int(boundaryWidth))))
        self.markDeviceWritten()

    def close(self):
        if getattr(self, "_dev", None) is not None and self._dev and self.parent._ctx:
            self.parent.lib.mp_grid_destroy(self._dev)
        self._dev = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RealGrid(_GridBase):
    """Grid<Real> (grid.h:92-235)"""
    KIND = MP_GRID_REAL

    def getMaxAbs(self):
        out = C.c_double(0)
        check(self.parent.lib.mp_grid_max_abs(self.parent._ctx, self.dev(), C.byref(out)))
        return out.value


class LevelsetGrid(RealGrid):
    pass


class MACGrid(_GridBase):
    """MACGrid (grid.h:243-281), AoS Vec3"""
    KIND = MP_GRID_MAC

    def getMaxAbs(self):
        """Grid<Vec3>::getMaxAbs grid.cpp:330-332: the largest |v| (what the liquid scenes hand to adaptTimestep)"""
        out = C.c_double(0)
        check(self.parent.lib.mp_grid_max_abs(self.parent._ctx, self.dev(), C.byref(out)))
        return out.value


class VecGrid(MACGrid):
    """Grid<Vec3> (cell-centred vectors): the same AoS storage as a MACGrid"""


class FlagGrid(_GridBase):
    """FlagGrid (grid.h:284-365)"""
    KIND = MP_GRID_FLAGS

    def initDomain(self, boundaryWidth=0, wall="xXyYzZ", open="      ", inflow="      ", outflow="      "):
        """grid.cpp:732-842: everything Empty, then the six boundary slabs typed wall/open/inflow/outflow."""
        wall, open_, inflow, outflow = (s + "      " for s in (wall, open, inflow, outflow))
        types = [0] * 6
        for d, ch in enumerate("xXyYzZ"):
            for i in range(6):
                if types[d]:
                    break
                if open_[i] == ch: types[d] = FlagOpen
                elif inflow[i] == ch: types[d] = FlagInflow
                elif outflow[i] == ch: types[d] = FlagOutflow
                elif wall[i] == ch: types[d] = FlagObstacle
        f = self._host
        f[...] = FlagEmpty
        w = boundaryWidth
        sx, sy, sz = self.size
        # the reference overwrites in the order x, X, y, Y, z, Z per cell (initBoundaries grid.cpp:826-842)
        f[:, :, :w + 1] = types[0]
        f[:, :, sx - 1 - w:] = types[1]
        f[:, :w + 1, :] = types[2]
        f[:, sy - 1 - w:, :] = types[3]
        if self.is3D():
            f[:w + 1, :, :] = types[4]
            f[sz - 1 - w:, :, :] = types[5]
        self._hostDirty, self._devDirty = True, False

    def fillGrid(self, type=FlagFluid):
        """grid.cpp:856-861"""
        f = self.numpy(writable=True)
        m = (f & (FlagObstacle | FlagInflow | FlagOutflow | FlagOpen)) == 0
        f[m] = (f[m] & ~(FlagEmpty | FlagFluid)) | type

    def updateFromLevelset(self, levelset):
        """grid.cpp:844-854 (invalidTimeValue = -1000, fastmarch.h:134), on the device: neither grid returns to the host"""
        check(self.parent.lib.mp_flags_update_from_levelset(self.parent._ctx, self.dev(), levelset.dev()))
        self.markDeviceWritten()
