"""Grid<T>::save / load (grid.cpp:113-156) for the device-resident grids: the reference's .uni, .raw and .npz grid files
(fileio/iogrids.cpp), read and written from the host mirror of a grid -- `save` downloads the device copy only if it is the newer one,
`load` marks the host copy as the newer one, so a loaded field reaches HBM with the next plugin that uses it.

.uni  (iogrids.cpp:34-44, :386-436, :440-520): gzip stream of  "MNT3" | UniHeader (288 bytes) | raw elements, x fastest.
        UniHeader = int dimX, dimY, dimZ, gridType, elementType, bytesPerElement; char info[252]; int dimT; unsigned long long timestamp.
        Values are always stored in single precision (the double build converts on the way, :66-84, gridReadConvert); flags as int32.
        The legacy "MNT2" header (256-byte info, no dimT / timestamp... :474-488) is read as well.
.raw  (iogrids.cpp:255-300): gzip stream of the raw elements in the grid's own precision, no header.
.npz  (iogrids.cpp:836-884, vendored cnpy): uncompressed zip with one array "arr_0" of shape [Z, Y, X, 1 or 3]; float build only in the
        reference (:841-844), float32 here for both precisions.
"""
import gzip
import struct
import time
import zipfile

import numpy as np

from ._lib import MP_GRID_FLAGS, MP_GRID_MAC, MantaError

# GridBase::GridType grid.h:29
TypeNone, TypeReal, TypeInt, TypeVec3, TypeMAC, TypeLevelset, TypeFlags = 0, 1, 2, 4, 8, 16, 32
_HEAD = struct.Struct("<6i252siQ")          # 288 bytes, the layout gcc gives UniHeader on x86-64 (no padding needed: 280 % 8 == 0)
_HEAD_MNT2 = struct.Struct("<6i256sQ")      # UniLegacyHeader3 iogrids.cpp (288 bytes as well)
assert _HEAD.size == 288 and _HEAD_MNT2.size == 288
INFO = b"mantaflow_b200 device-resident grid"


def grid_type_of(grid):
    """mType as the reference's constructors set it (grid.h:246,:287, levelset.cpp:94, grid.cpp:47-59)"""
    from . import grid as G
    if isinstance(grid, G.FlagGrid):
        return TypeFlags | TypeInt
    if isinstance(grid, G.VecGrid):
        return TypeVec3
    if isinstance(grid, G.MACGrid):
        return TypeMAC | TypeVec3
    if isinstance(grid, G.LevelsetGrid):
        return TypeLevelset | TypeReal
    return TypeReal


def _unify(t):
    """unifyGridType iogrids.cpp:213-221: real <-> levelset, vec3 <-> mac"""
    if t & TypeReal: t |= TypeLevelset
    if t & TypeLevelset: t |= TypeReal
    if t & TypeVec3: t |= TypeMAC
    if t & TypeMAC: t |= TypeVec3
    return t


def _ext(name):
    if "." not in name.rsplit("/", 1)[-1]:
        raise MantaError(1, "file '%s' does not have an extension" % name)
    return name[name.rfind("."):]


def write_uni(name, host, grid_type):
    """host: [Z,Y,X] int32 / real, or [Z,Y,X,3] real"""
    sz, sy, sx = host.shape[:3]
    if grid_type & TypeInt:
        et, data = 0, np.ascontiguousarray(host, dtype=np.int32)
    elif grid_type & TypeReal:
        et, data = 1, np.ascontiguousarray(host, dtype=np.float32)
    elif grid_type & TypeVec3:
        et, data = 2, np.ascontiguousarray(host, dtype=np.float32)
    else:
        raise MantaError(1, "writeGridUni: unknown element type")
    bpe = 12 if et == 2 else 4
    head = _HEAD.pack(sx, sy, sz, grid_type, et, bpe, INFO, 0, int(time.time() * 1000))
    with gzip.open(name, "wb", compresslevel=1) as f:      # "wb1" iogrids.cpp:413
        f.write(b"MNT3"); f.write(head); f.write(data.tobytes())
    return 1


def read_uni(name, shape, grid_type, dtype):
    """returns the array of `shape` ([Z,Y,X] or [Z,Y,X,3]) in `dtype`; checks size and type like readGridUni (iogrids.cpp:489-506)"""
    try:
        f = gzip.open(name, "rb")
    except OSError:
        raise MantaError(1, "readGridUni: can't open file " + name)
    with f:
        magic = f.read(4)
        if magic == b"MNT3":
            raw = f.read(_HEAD.size)
            if len(raw) != _HEAD.size:
                raise MantaError(1, "can't read file, no header present")
            dx, dy, dz, gt, et, bpe, _info, _dimT, _stamp = _HEAD.unpack(raw)
        elif magic == b"MNT2":
            raw = f.read(_HEAD_MNT2.size)
            if len(raw) != _HEAD_MNT2.size:
                raise MantaError(1, "can't read file, no header present")
            dx, dy, dz, gt, et, bpe, _info, _stamp = _HEAD_MNT2.unpack(raw)
        else:
            raise MantaError(1, "readGridUni: Unknown header '%s' " % magic.decode(errors="replace"))
        sz, sy, sx = shape[:3]
        if (dx, dy, dz) != (sx, sy, sz):
            raise MantaError(1, "grid dim doesn't match, [%d,%d,%d] vs [%d,%d,%d]" % (dx, dy, dz, sx, sy, sz))
        if _unify(gt) != _unify(grid_type):
            raise MantaError(1, "grid type doesn't match %d vs %d" % (gt, grid_type))
        n = sx * sy * sz * (3 if len(shape) == 4 else 1)
        per = bpe // (3 if len(shape) == 4 else 1)
        if per not in (4, 8) or bpe != per * (3 if len(shape) == 4 else 1):
            raise MantaError(1, "grid element size doesn't match %d" % bpe)
        src = np.int32 if grid_type & TypeInt else (np.float32 if per == 4 else np.float64)
        buf = f.read(n * per)
        if len(buf) != n * per:
            raise MantaError(1, "readGridUni: file %s is truncated" % name)
        return np.frombuffer(buf, dtype=src).reshape(shape).astype(dtype)


def write_raw(name, host):
    with gzip.open(name, "wb", compresslevel=1) as f:
        f.write(np.ascontiguousarray(host).tobytes())
    return 1


def read_raw(name, shape, dtype):
    try:
        f = gzip.open(name, "rb")
    except OSError:
        raise MantaError(1, "readGridRaw: can't open file " + name)
    with f:
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = f.read(n)
        if len(buf) != n:
            raise MantaError(1, "readGridRaw: can't read raw file, stream length does not match " + name)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


def write_npz(name, host, is_int):
    a = np.ascontiguousarray(host, dtype=np.int32 if is_int else np.float32)
    if a.ndim == 3:
        a = a[..., None]                     # shape [Z, Y, X, 1] iogrids.cpp:868
    with zipfile.ZipFile(name, "w", zipfile.ZIP_STORED) as z:       # "generates a zip file without compression" :871
        with z.open("arr_0.npy", "w") as f:
            np.lib.format.write_array(f, a, version=(1, 0))
    return 1


def read_npz(name, shape, dtype):
    try:
        with np.load(name) as z:
            a = z["arr_0"]
    except (OSError, KeyError):
        raise MantaError(1, "readGridNumpy: can't read arr_0 from " + name)
    want = tuple(shape) if len(shape) == 4 else tuple(shape) + (1,)
    if tuple(a.shape) != want:
        raise MantaError(1, "grid dim doesn't match, %s vs %s" % (a.shape, want))
    return a.reshape(shape).astype(dtype)


def save(grid, name):
    """Grid<T>::save grid.cpp:134-156"""
    ext = _ext(name)
    host = grid.numpy()                      # downloads only if the device copy is the newer one
    if ext == ".uni":
        return write_uni(name, host, grid_type_of(grid))
    if ext == ".raw":
        return write_raw(name, host)
    if ext == ".npz":
        return write_npz(name, host, grid.KIND == MP_GRID_FLAGS)
    raise MantaError(1, "file '%s' filetype not supported" % name)


def load(grid, name):
    """Grid<T>::load grid.cpp:112-132"""
    ext = _ext(name)
    shape = grid._host.shape
    dtype = grid._host.dtype
    if ext == ".uni":
        a = read_uni(name, shape, grid_type_of(grid), dtype)
    elif ext == ".raw":
        a = read_raw(name, shape, dtype)
    elif ext == ".npz":
        a = read_npz(name, shape, dtype)
    else:
        raise MantaError(1, "file '%s' filetype not supported" % name)
    grid.copyFromArray(a)
    return 1
