"""CPU (not gpu): the grid files of mantaflow_b200/fileio.py (.uni, .raw, .npz -- Grid<T>::save / load, grid.cpp:113-156,
fileio/iogrids.cpp) are the reference's: files written by the unmodified reference (committed under tests/golden/io/, generator
tests/golden/make_golden.py --only-io) are read back exactly, files written here are read by the reference (when oracle/_ref is
built), headers agree field by field."""
import gzip
import os

import numpy as np
import pytest

from mantaflow_b200 import MantaError, fileio

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "golden", "io")
KINDS = {   # kind -> (GridType value the reference's constructors set, components)
    "real": (fileio.TypeReal, 1), "levelset": (fileio.TypeLevelset | fileio.TypeReal, 1), "mac": (fileio.TypeMAC | fileio.TypeVec3, 3),
    "vec3": (fileio.TypeVec3, 3), "flags": (fileio.TypeFlags | fileio.TypeInt, 1),
}
SHAPES = {"3d": (5, 6, 7), "2d": (1, 9, 8)}


def sample(kind, tag, prec):
    shape = SHAPES[tag] + ((3,) if KINDS[kind][1] == 3 else ())
    rng = np.random.default_rng(len(kind) + shape[1])
    if kind == "flags":
        return rng.integers(0, 128, shape).astype(np.int32)
    a = (rng.random(shape) * 8 - 4).astype(np.float32)          # float32-representable, so that the double build round-trips exactly too
    return a.astype(np.float32 if prec == 4 else np.float64)


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("tag", list(SHAPES))
@pytest.mark.parametrize("kind", list(KINDS))
def test_reads_uni_files_written_by_the_reference(kind, tag, prec):
    a = sample(kind, tag, prec)
    path = os.path.join(IO, "ref_%s_%s_f%d.uni" % (kind, tag, prec * 8))
    assert np.array_equal(fileio.read_uni(path, a.shape, KINDS[kind][0], a.dtype), a)
    # real <-> levelset and vec3 <-> mac are interchangeable (unifyGridType iogrids.cpp:213-221), anything else is an error
    other = {"real": "levelset", "levelset": "real", "mac": "vec3", "vec3": "mac"}.get(kind)
    if other:
        assert np.array_equal(fileio.read_uni(path, a.shape, KINDS[other][0], a.dtype), a)
    with pytest.raises(MantaError):
        fileio.read_uni(path, a.shape, KINDS["flags" if kind != "flags" else "real"][0], a.dtype)
    with pytest.raises(MantaError):
        fileio.read_uni(path, (a.shape[0], a.shape[1] + 1) + a.shape[2:], KINDS[kind][0], a.dtype)


@pytest.mark.parametrize("kind", list(KINDS))
def test_uni_header_matches_the_reference_field_by_field(kind, tmp_path):
    a = sample(kind, "3d", 4)
    mine = str(tmp_path / "mine.uni")
    fileio.write_uni(mine, a, KINDS[kind][0])
    ref = os.path.join(IO, "ref_%s_3d_f32.uni" % kind)
    hm, hr = gzip.open(mine).read(), gzip.open(ref).read()
    assert hm[:4] == hr[:4] == b"MNT3" and len(hm) == len(hr) == 4 + 288 + a.nbytes
    assert hm[4:28] == hr[4:28], "dimX, dimY, dimZ, gridType, elementType, bytesPerElement"
    assert hm[280:284] == hr[280:284], "dimT"
    assert hm[292:] == hr[292:], "payload"


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("ext", [".uni", ".raw", ".npz"])
@pytest.mark.parametrize("kind", list(KINDS))
def test_round_trip_with_the_reference(kind, ext, prec, tmp_path, ref32, ref64):
    """written here -> loaded by the reference, written by the reference -> loaded here (the reference writes no .npz in the double build)"""
    R = ref32 if prec == 4 else ref64
    a = sample(kind, "3d", prec)
    gt = KINDS[kind][0]
    mine, theirs = str(tmp_path / ("mine" + ext)), str(tmp_path / ("theirs" + ext))
    write = {".uni": lambda n: fileio.write_uni(n, a, gt), ".raw": lambda n: fileio.write_raw(n, a), ".npz": lambda n: fileio.write_npz(n, a, kind == "flags")}[ext]
    read = {".uni": lambda n: fileio.read_uni(n, a.shape, gt, a.dtype), ".raw": lambda n: fileio.read_raw(n, a.shape, a.dtype),
            ".npz": lambda n: fileio.read_npz(n, a.shape, a.dtype)}[ext]
    write(mine)
    assert np.array_equal(read(mine), a)
    if ext == ".npz" and prec == 8:
        return
    got = np.zeros_like(a)
    R.grid_file(mine, got, kind, load=True)
    assert np.array_equal(got, a), "the reference does not read back what was written here"
    R.grid_file(theirs, a.copy(), kind, load=False)
    assert np.array_equal(read(theirs), a), "what the reference wrote is not read back here"


def test_errors(tmp_path):
    a = sample("real", "3d", 4)
    with pytest.raises(MantaError):
        fileio.read_uni(str(tmp_path / "missing.uni"), a.shape, fileio.TypeReal, a.dtype)
    bad = str(tmp_path / "bad.uni")
    with gzip.open(bad, "wb") as f:
        f.write(b"XXXX" + b"\0" * 300)
    with pytest.raises(MantaError):
        fileio.read_uni(bad, a.shape, fileio.TypeReal, a.dtype)
    with pytest.raises(MantaError):
        fileio._ext("noextension")
