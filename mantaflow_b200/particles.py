"""Host-side mirror of the reference's particle types and FLIP particle <-> grid plugins (second slice), device resident:

    BasicParticleSystem   particle.h:182-274   (pos + flag per particle; PDELETE = 1 << 10 marks a deleted one)
    ParticleIndexSystem   particle.h:276-300   (sourceIndex per slot)
    PdataVec3 / PdataInt / PdataReal   ParticleDataImpl<T> particle.h:392-470
    IntGrid               Grid<int>

    markFluidCells         plugin/flip.cpp:158-177      gridParticleIndex     plugin/flip.cpp:260-306
    unionParticleLevelset  plugin/flip.cpp:340-350      mapPartsToMAC         plugin/flip.cpp:573-595
    mapMACToParts          plugin/flip.cpp:651-656      flipVelocityUpdate    plugin/flip.cpp:669-677
    pushOutofObs           plugin/flip.cpp:542-545      BasicParticleSystem.advectInGrid / projectOutOfBnd  particle.h:154-158

Same names, argument order and defaults as the reference.  Every particle array owns a numpy array AND a device array (an mp_grid of
size (capacity, 1, 1)); two dirty bits keep them coherent lazily, as for the grids (grid.py), so the plugins of a FLIP step
(scenes/benchmark_dam.py:100-125) never leave the device.  There is no CPU fallback."""
import ctypes as C
import numpy as np

from ._lib import MP_GRID_FLAGS, MP_GRID_MAC, MP_GRID_REAL, MantaError, check
from .grid import _GridBase

IntEuler, IntRK2, IntRK4 = 0, 1, 2      # util/integrator.h:23
PNONE, PNEW, PSPRAY, PBUBBLE, PFOAM, PTRACER, PDELETE, PINVALID = 0, 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 10, 1 << 30      # ParticleBase::ParticleStatus particle.h:34-43


class IntGrid(_GridBase):
    """Grid<int> (the `index` argument of gridParticleIndex): the storage of a FlagGrid"""
    KIND = MP_GRID_FLAGS


class _DevArray:
    """a per-particle array: numpy [n] / [n,3] on the host, mp_grid (capacity,1,1) on the device, coherent lazily"""

    def __init__(self, solver, kind, n=0):
        self.solver, self.kind = solver, kind
        self.dtype = np.int32 if kind == MP_GRID_FLAGS else solver.real
        self._tail = (3,) if kind == MP_GRID_MAC else ()
        self._host = np.zeros((n,) + self._tail, self.dtype)
        self._dev, self._cap = C.c_void_p(), 0
        self._hostDirty, self._devDirty = n > 0, False
        solver._adopt(self)

    def __len__(self):
        return len(self._host)

    def _reserve(self, n):
        if n <= self._cap:
            return
        lib = self.solver.lib
        if self._dev:
            check(lib.mp_grid_destroy(self._dev))
        self._dev, self._cap = C.c_void_p(), max(n, 1)
        check(lib.mp_grid_create(self.solver._ctx, C.c_int(self.kind), C.c_int(self.solver.prec), C.c_int(self._cap), C.c_int(1), C.c_int(1), C.byref(self._dev)))

    def resize(self, n):
        """keeps the first min(old, n) entries (of whichever copy is current), new entries are zero"""
        if n == len(self._host):
            return
        old = self.numpy()
        new = np.zeros((n,) + self._tail, self.dtype)
        m = min(n, len(old))
        new[:m] = old[:m]
        self._host, self._hostDirty, self._devDirty = new, True, False

    def set(self, arr):
        arr = np.asarray(arr, dtype=self.dtype)
        self._host = np.array(arr.reshape((-1,) + self._tail), order="C", copy=True)
        self._hostDirty, self._devDirty = True, False

    def dev(self):
        """mp_grid handle with the device copy current (uploads only if the host copy is newer); None for an empty array"""
        n = len(self._host)
        if n == 0:
            return None
        if self._hostDirty or n > self._cap:
            self._reserve(n)
            buf = self._host
            if self._cap != n:          # the upload moves `capacity` entries
                buf = np.zeros((self._cap,) + self._tail, self.dtype)
                buf[:n] = self._host
            check(self.solver.lib.mp_grid_upload(self._dev, buf.ctypes.data_as(C.c_void_p)))
            self._hostDirty = False
        return self._dev

    def markDeviceWritten(self):
        self._devDirty, self._hostDirty = True, False

    def numpy(self, writable=False):
        if self._devDirty and self._dev:
            buf = np.zeros((self._cap,) + self._tail, self.dtype)
            check(self.solver.lib.mp_grid_download(self._dev, buf.ctypes.data_as(C.c_void_p)))
            self._host = np.ascontiguousarray(buf[:len(self._host)])
            self._devDirty = False
        if writable:
            self._hostDirty = True
        return self._host

    def close(self):
        if getattr(self, "_dev", None) is not None and self._dev and self.solver._ctx:
            self.solver.lib.mp_grid_destroy(self._dev)
        self._dev, self._cap = C.c_void_p(), 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Pdata:
    """ParticleDataImpl<T> particle.h:392-470: follows the size of the particle system it was created in"""
    KIND = MP_GRID_REAL

    def __init__(self, parent, parts=None):
        self.parent, self.parts = parent, parts
        self._a = _DevArray(parent, self.KIND, parts.size() if parts is not None else 0)

    def size(self): return len(self._a)
    def numpy(self, writable=False): return self._a.numpy(writable)
    def copyFromArray(self, arr): self._a.set(arr)
    def dev(self): return self._a.dev()
    def markDeviceWritten(self): self._a.markDeviceWritten()

    def setConst(self, value):
        self._a.numpy(writable=True)[...] = value

    def clear(self):
        self.setConst(0)


class PdataReal(_Pdata):
    KIND = MP_GRID_REAL


class PdataInt(_Pdata):
    KIND = MP_GRID_FLAGS


class PdataVec3(_Pdata):
    KIND = MP_GRID_MAC


class BasicParticleSystem:
    """BasicParticleSystem particle.h:182-274 (positions + status flags), with the data fields created through it"""

    def __init__(self, parent):
        self.parent = parent
        self._pos = _DevArray(parent, MP_GRID_MAC)
        self._flag = _DevArray(parent, MP_GRID_FLAGS)
        self._pdata = []

    def create(self, cls, **kw):
        """ParticleBase::create particle.cpp:60-70: a data field that is resized with the system"""
        pd = cls(self.parent, parts=self, **kw)
        self._pdata.append(pd)
        return pd

    def size(self): return len(self._pos)
    pySize = size

    def resize(self, n):
        for a in [self._pos, self._flag] + [pd._a for pd in self._pdata]:
            a.resize(n)

    def setParticles(self, pos, flag=None):
        """positions [n,3] (and status flags [n]); the data fields are resized to n"""
        pos = np.asarray(pos).reshape(-1, 3)
        self._pos.set(pos)
        self._flag.set(np.zeros(len(pos), np.int32) if flag is None else flag)
        if len(self._flag) != len(pos):
            raise MantaError(1, "BasicParticleSystem.setParticles: pos and flag differ in length")
        for pd in self._pdata:
            pd._a.resize(len(pos))

    def advectInGrid(self, flags, vel, integrationMode, deleteInObstacle=True, stopInObstacle=True, skipNew=False, ptype=None, exclude=0):
        """ParticleSystem::advectInGrid particle.h:154,:512-536 (IntEuler / IntRK2 / IntRK4), on the device: positions and flags stay there"""
        s = self.parent
        _pdcheck(self, ptype, "ptype")
        if self.size() == 0:
            return
        check(s.lib.mp_parts_advect_in_grid(s._ctx, flags.dev(), vel.dev(), C.c_longlong(self.size()), self._pos.dev(), self._flag.dev(), C.c_double(s.timestep),
                                            C.c_int(int(integrationMode)), C.c_int(int(deleteInObstacle)), C.c_int(int(stopInObstacle)), C.c_int(int(skipNew)),
                                            None if ptype is None else ptype.dev(), C.c_int(exclude)))
        self._pos.markDeviceWritten()
        self._flag.markDeviceWritten()

    def projectOutOfBnd(self, flags, bnd, plane="xXyYzZ", ptype=None, exclude=0):
        """ParticleSystem::projectOutOfBnd particle.h:158,:578-590, on the device"""
        s = self.parent
        _pdcheck(self, ptype, "ptype")
        if self.size() == 0:
            return
        check(s.lib.mp_parts_project_out_of_bnd(s._ctx, flags.dev(), C.c_longlong(self.size()), self._pos.dev(), self._flag.dev(), C.c_double(bnd), str(plane).encode(),
                                                None if ptype is None else ptype.dev(), C.c_int(exclude)))
        self._pos.markDeviceWritten()

    def getPosPdata(self, target):
        """ParticleSystem::getPosPdata particle.h:135,:422-427: a device-to-device copy of the positions"""
        s = self.parent
        _pdcheck(self, target, "target")
        if self.size() == 0:
            return
        target._a._reserve(self.size())
        target._a._hostDirty = False       # every entry in use is written
        check(s.lib.mp_parts_get_pos_pdata(s._ctx, C.c_longlong(self.size()), self._pos.dev(), target._a._dev))
        target.markDeviceWritten()

    def positions(self, writable=False): return self._pos.numpy(writable)
    def flags(self, writable=False): return self._flag.numpy(writable)

    def clear(self):
        self.resize(0)


class ParticleIndexSystem:
    """ParticleIndexSystem particle.h:276-300: sourceIndex of every slot, filled by gridParticleIndex"""

    def __init__(self, parent):
        self.parent = parent
        self._a = _DevArray(parent, MP_GRID_FLAGS)
        self._count = 0

    def size(self): return self._count
    def numpy(self): return self._a.numpy()[:self._count]


def _d(g):
    return None if g is None else g.dev()


def _ps(parts):
    return C.c_longlong(parts.size()), parts._pos.dev(), parts._flag.dev()


def _pdcheck(parts, pd, name):
    if pd is not None and pd.size() != parts.size():
        raise MantaError(1, "%s holds %d entries, the particle system %d" % (name, pd.size(), parts.size()))


def markFluidCells(parts, flags, phiObs=None, ptype=None, exclude=0):
    s = flags.parent
    _pdcheck(parts, ptype, "ptype")
    n, pos, pflag = _ps(parts)
    check(s.lib.mp_mark_fluid_cells(s._ctx, n, pos, pflag, flags.dev(), _d(phiObs), _d(ptype), C.c_int(exclude)))
    flags.markDeviceWritten()


def gridParticleIndex(parts, indexSys, flags, index, counter=None):
    """`counter` (a scratch Grid<int> in the reference) is not needed: the slots of a cell are filled by a stable sort"""
    s = index.parent
    n, pos, pflag = _ps(parts)
    indexSys._a.resize(parts.size())
    indexSys._a._hostDirty = False      # every slot in use is written on the device
    if parts.size():
        indexSys._a._reserve(parts.size())
    count = C.c_longlong(0)
    check(s.lib.mp_grid_particle_index(s._ctx, n, pos, pflag, indexSys._a._dev if parts.size() else None, _d(flags), index.dev(), C.byref(count)))
    indexSys._count = count.value
    if parts.size():
        indexSys._a.markDeviceWritten()
    index.markDeviceWritten()


def unionParticleLevelset(parts, indexSys, flags, index, phi, radiusFactor=1., ptype=None, exclude=0):
    s = phi.parent
    _pdcheck(parts, ptype, "ptype")
    n, pos, _ = _ps(parts)
    isys = indexSys._a.dev() if indexSys._count else None
    check(s.lib.mp_union_particle_levelset(s._ctx, n, pos, isys, C.c_longlong(indexSys._count), _d(flags), index.dev(), phi.dev(), C.c_double(radiusFactor),
                                           _d(ptype), C.c_int(exclude)))
    phi.markDeviceWritten()


def pushOutofObs(parts, flags, phiObs, shift=0, thresh=0, ptype=None, exclude=0):
    s = phiObs.parent
    _pdcheck(parts, ptype, "ptype")
    if parts.size() == 0:
        return
    n, pos, pflag = _ps(parts)
    check(s.lib.mp_push_out_of_obs(s._ctx, n, pos, pflag, _d(flags), phiObs.dev(), C.c_double(shift), C.c_double(thresh), _d(ptype), C.c_int(exclude)))
    parts._pos.markDeviceWritten()


def addForcePvel(vel, a, dt, ptype, exclude):
    """plugin/ptsplugins.cpp:26 (no defaults in the reference either)"""
    s = vel.parent
    if vel.size() == 0:
        return
    ax, ay, az = (float(c) for c in a)
    check(s.lib.mp_add_force_pvel(s._ctx, C.c_longlong(vel.size()), vel.dev(), C.c_double(ax), C.c_double(ay), C.c_double(az), C.c_double(dt), _d(ptype), C.c_int(exclude)))
    vel.markDeviceWritten()


def updateVelocityFromDeltaPos(parts, vel, x_prev, dt, ptype, exclude):
    """plugin/ptsplugins.cpp:38"""
    s = vel.parent
    _pdcheck(parts, vel, "vel"); _pdcheck(parts, x_prev, "x_prev"); _pdcheck(parts, ptype, "ptype")
    if parts.size() == 0:
        return
    check(s.lib.mp_update_velocity_from_delta_pos(s._ctx, C.c_longlong(parts.size()), parts._pos.dev(), vel.dev(), x_prev.dev(), C.c_double(dt), _d(ptype), C.c_int(exclude)))
    vel.markDeviceWritten()


def eulerStep(parts, vel, ptype, exclude):
    """plugin/ptsplugins.cpp:50 (dt = the solver's timestep)"""
    s = parts.parent
    _pdcheck(parts, vel, "vel"); _pdcheck(parts, ptype, "ptype")
    if parts.size() == 0:
        return
    check(s.lib.mp_euler_step(s._ctx, C.c_longlong(parts.size()), parts._pos.dev(), vel.dev(), C.c_double(s.timestep), _d(ptype), C.c_int(exclude)))
    parts._pos.markDeviceWritten()


def setPartType(parts, ptype, mark, stype, flags, cflag):
    """plugin/ptsplugins.cpp:62: particles of type `stype` in `cflag` cells become `mark`"""
    s = flags.parent
    _pdcheck(parts, ptype, "ptype")
    if parts.size() == 0:
        return
    check(s.lib.mp_set_part_type(s._ctx, C.c_longlong(parts.size()), parts._pos.dev(), ptype.dev(), C.c_int(mark), C.c_int(stype), flags.dev(), C.c_int(cflag)))
    ptype.markDeviceWritten()


def markIsolatedFluidCell(flags, mark):
    """grid.cpp:885-890"""
    s = flags.parent
    check(s.lib.mp_mark_isolated_fluid_cell(s._ctx, flags.dev(), C.c_int(mark)))
    flags.markDeviceWritten()


def mapPartsToMAC(flags, vel, velOld, parts, partVel, weight=None, ptype=None, exclude=0):
    s = vel.parent
    _pdcheck(parts, partVel, "partVel"); _pdcheck(parts, ptype, "ptype")
    n, pos, pflag = _ps(parts)
    check(s.lib.mp_map_parts_to_mac(s._ctx, _d(flags), vel.dev(), velOld.dev(), n, pos, pflag, partVel.dev(), _d(weight), _d(ptype), C.c_int(exclude)))
    vel.markDeviceWritten(); velOld.markDeviceWritten()
    if weight is not None:
        weight.markDeviceWritten()


def mapMACToParts(flags, vel, parts, partVel, ptype=None, exclude=0):
    s = vel.parent
    _pdcheck(parts, partVel, "partVel"); _pdcheck(parts, ptype, "ptype")
    n, pos, pflag = _ps(parts)
    check(s.lib.mp_map_mac_to_parts(s._ctx, _d(flags), vel.dev(), n, pos, pflag, partVel.dev(), _d(ptype), C.c_int(exclude)))
    if parts.size():
        partVel.markDeviceWritten()


def flipVelocityUpdate(flags, vel, velOld, parts, partVel, flipRatio, ptype=None, exclude=0):
    s = vel.parent
    _pdcheck(parts, partVel, "partVel"); _pdcheck(parts, ptype, "ptype")
    n, pos, pflag = _ps(parts)
    check(s.lib.mp_flip_velocity_update(s._ctx, _d(flags), vel.dev(), velOld.dev(), n, pos, pflag, partVel.dev(), C.c_double(flipRatio), _d(ptype), C.c_int(exclude)))
    if parts.size():
        partVel.markDeviceWritten()
