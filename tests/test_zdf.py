"""ZDF output: files written by the product must be byte-compatible with the reference writer and
readable by our reader (and by the reference's own python/lib/zdf.py when it is around)."""
import ctypes as C
import filecmp
import os
import sys

import numpy as np
import pytest

from tests import helpers as H
from zpic_b200 import abi_em2d as A
from zpic_b200 import zdf


def _report_set(deck):
    """the diagnostics of the shipped Weibel deck that need no device (reference input/weibel.c:44-57)"""
    L = deck.lib
    L.emf_report(C.byref(deck.sim.emf), bytes([A.BFLD]), 0)
    L.emf_report(C.byref(deck.sim.emf), bytes([A.EFLD]), 2)
    L.current_report(C.byref(deck.sim.current), 2)
    L.spec_report(deck.species[0], 0x3000, None, None)          # PARTICLES
    nx = (C.c_int * 2)(64, 32)
    rng = ((C.c_float * 2) * 2)((C.c_float * 2)(0.0, 3.2), (C.c_float * 2)(-1.0, 1.0))
    L.spec_report(deck.species[1], 0x2000 + 1 + 16 * 6, nx, rng)  # PHASESPACE(X1, U3)


def test_iteration0_files_identical_to_reference(ours, ref, tmp_path):
    cwd = os.getcwd()
    dirs = {}
    try:
        for name, lib in (("ours", ours), ("ref", ref)):
            d = tmp_path / name
            d.mkdir()
            os.chdir(d)
            deck = H.weibel(lib, n=32, ppc=(2, 2))
            _report_set(deck)
            dirs[name] = d
    finally:
        os.chdir(cwd)
    files = sorted(p.relative_to(dirs["ref"]) for p in dirs["ref"].rglob("*.zdf"))
    assert len(files) == 5
    for f in files:
        assert (dirs["ours"] / f).exists(), f
        assert filecmp.cmp(dirs["ours"] / f, dirs["ref"] / f, shallow=False), f


def test_reader_round_trip(ours, tmp_path):
    cwd = os.getcwd()
    try:
        os.chdir(tmp_path)
        deck = H.weibel(ours, n=32, ppc=(2, 2))
        _report_set(deck)
    finally:
        os.chdir(cwd)
    data, info = zdf.read(str(tmp_path / "PARTICLES" / "electrons" / "particles-electrons-000000.zdf"))
    assert info.type == "particles" and info.particles.nparts == 32 * 32 * 4
    p = deck.parts(0)
    dx = deck.sim.emf.dx[0]
    assert np.array_equal(data["ux"], p["ux"])
    assert np.array_equal(data["x"], ((p["ix"] + p["x"]) * np.float32(dx)).astype(np.float32))
    data, info = zdf.read(str(tmp_path / "PHASESPACE" / "positrons" / "positrons-x1u3-000000.zdf"))
    assert info.type == "grid" and data.shape == (32, 64)
    assert info.grid.axis[1].min == -1.0 and info.iteration.n == 0
    assert 0.9 * 1024 < data.sum() <= 1024.0          # total charge, minus the weight falling off the grid edges
    ref_reader = "/root/reference/python/lib"
    if os.path.isdir(ref_reader):
        sys.path.insert(0, ref_reader)
        import zdf as refzdf          # the reference's reader must parse our files too
        d2, i2 = refzdf.read(str(tmp_path / "PHASESPACE" / "positrons" / "positrons-x1u3-000000.zdf"))
        sys.path.remove(ref_reader)
        assert np.array_equal(d2, data) and i2.grid.nx[0] == 64


# ---------------------------------------------------------------------------------------------------
# the rest of the zdf.h API (tracks metadata, chunked datasets, updating a file): same calls on both libraries

class _ZFile(C.Structure):       # t_zdf_file (include/em2d/zdf.h)
    _fields_ = [("fp", C.c_void_p), ("mode", C.c_int), ("ndatasets", C.c_uint32)]


class _ZDataset(C.Structure):    # t_zdf_dataset
    _fields_ = [("name", C.c_char_p), ("data_type", C.c_int), ("ndims", C.c_uint32), ("count", C.c_uint64 * 3),
                ("data", C.c_void_p), ("id", C.c_uint64), ("offset", C.c_uint64)]


class _ZChunk(C.Structure):      # t_zdf_chunk
    _fields_ = [("count", C.c_uint64 * 3), ("start", C.c_uint64 * 3), ("stride", C.c_uint64 * 3), ("data", C.c_void_p)]


class _ZTracks(C.Structure):     # t_zdf_track_info
    _fields_ = [("name", C.c_char_p), ("label", C.c_char_p), ("ntracks", C.c_uint32), ("ndump", C.c_uint32),
                ("niter", C.c_uint32), ("nquants", C.c_uint32), ("quants", C.POINTER(C.c_char_p)),
                ("qlabels", C.POINTER(C.c_char_p)), ("qunits", C.POINTER(C.c_char_p))]


ZDF_CREATE, ZDF_READ, ZDF_UPDATE = 0, 1, 2
T_UINT8, T_INT16, T_FLOAT32, T_FLOAT64 = 2, 3, 9, 10        # enum zdf_data_type


def _declare_zdf(lib):
    P = C.POINTER
    for name, res, args in (
            ("zdf_open_file", C.c_int, [P(_ZFile), C.c_char_p, C.c_int]), ("zdf_close_file", C.c_int, [P(_ZFile)]),
            ("zdf_add_dataset", C.c_size_t, [P(_ZFile), P(_ZDataset)]),
            ("zdf_add_track_info", C.c_size_t, [P(_ZFile), P(_ZTracks)]),
            ("zdf_vector_write", C.c_size_t, [P(_ZFile), C.c_void_p, C.c_int, C.c_size_t]),
            ("zdf_start_cdset", C.c_size_t, [P(_ZFile), P(_ZDataset)]),
            ("size_zdf_chunk_header", C.c_size_t, [P(_ZDataset)]),
            ("zdf_write_chunk_header", C.c_size_t, [P(_ZFile), P(_ZDataset), P(_ZChunk)]),
            ("zdf_write_cdset", C.c_size_t, [P(_ZFile), P(_ZDataset), P(_ZChunk)]),
            ("zdf_end_cdset", C.c_size_t, [P(_ZFile), P(_ZDataset)]),
            ("zdf_open_dataset", C.c_size_t, [P(_ZFile), P(_ZDataset)]),
            ("zdf_extend_dataset", C.c_int, [P(_ZFile), P(_ZDataset), P(C.c_uint64)])):
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args


def _tracks_and_chunks(lib, path):
    """one call sequence over the whole low-level API; returns every return value and the dataset bookkeeping"""
    _declare_zdf(lib)
    out = []
    f = _ZFile()
    assert lib.zdf_open_file(C.byref(f), path, ZDF_CREATE) == 1
    names = [(C.c_char_p * 3)(*v) for v in ((b"t", b"x1", b"ene"), (b"t", b"x_1", b"Energy"), (b"1/w_p", b"c/w_p", b"m_e c^2"))]
    tr = _ZTracks(b"electrons", b"test tracks", 7, 100, 5, 3, names[0], names[1], names[2])
    out.append(lib.zdf_add_track_info(C.byref(f), C.byref(tr)))
    # plain datasets of the narrow types: 8-bit data is padded to 4 bytes, 16-bit data is not
    # (8 bytes: the reference does not count the padding of an 8-bit vector in the record length, so a length
    # that is not a multiple of 4 would derail its own record scan in zdf_open_dataset - and ours, which follows it)
    a8 = np.arange(8, dtype=np.uint8)
    a16 = np.arange(5, dtype=np.int16) - 2
    for name, arr, t in ((b"bytes", a8, T_UINT8), (b"shorts", a16, T_INT16)):
        ds = _ZDataset(name, t, 1, (C.c_uint64 * 3)(arr.size, 0, 0), arr.ctypes.data, 0, 0)
        out += [lib.zdf_add_dataset(C.byref(f), C.byref(ds)), ds.id, ds.offset]
    empty = _ZDataset(b"nothing", T_FLOAT32, 1, (C.c_uint64 * 3)(0, 0, 0), a8.ctypes.data, 0, 0)
    out.append(lib.zdf_add_dataset(C.byref(f), C.byref(empty)))       # reported like a failed write by both
    # a 10 x 6 float32 dataset written in two chunks, then closed
    full = np.arange(60, dtype=np.float32).reshape(6, 10)
    ds = _ZDataset(b"tracks", T_FLOAT32, 2, (C.c_uint64 * 3)(10, 6, 0), None, 0, 0)
    out += [lib.zdf_start_cdset(C.byref(f), C.byref(ds)), ds.id, ds.offset, lib.size_zdf_chunk_header(C.byref(ds))]
    for j0, nj in ((0, 4), (4, 2)):
        part = np.ascontiguousarray(full[j0:j0 + nj])
        ch = _ZChunk((C.c_uint64 * 3)(10, nj, 0), (C.c_uint64 * 3)(0, j0, 0), (C.c_uint64 * 3)(1, 1, 0), part.ctypes.data)
        out.append(lib.zdf_write_cdset(C.byref(f), C.byref(ds), C.byref(ch)))
    raw = np.linspace(0, 1, 3)
    out.append(lib.zdf_vector_write(C.byref(f), raw.ctypes.data, T_FLOAT64, 3))
    out.append(lib.zdf_vector_write(C.byref(f), a8.ctypes.data, T_UINT8, 5))
    assert lib.zdf_close_file(C.byref(f)) == 1
    return out


def _grow(lib, path):
    """re-open, find the chunked dataset, extend it by two rows, append the chunk, close the dataset"""
    f = _ZFile()
    assert lib.zdf_open_file(C.byref(f), path, ZDF_UPDATE) == 1
    out = []
    missing = _ZDataset(b"no such dataset", 0, 0, (C.c_uint64 * 3)(0, 0, 0), None, 0, 0)
    out.append(lib.zdf_open_dataset(C.byref(f), C.byref(missing)))
    ds = _ZDataset(b"tracks", 0, 0, (C.c_uint64 * 3)(0, 0, 0), None, 0, 0)
    out += [lib.zdf_open_dataset(C.byref(f), C.byref(ds)), ds.id, ds.offset, ds.data_type, ds.ndims, list(ds.count)[:2]]
    out.append(lib.zdf_extend_dataset(C.byref(f), C.byref(ds), (C.c_uint64 * 3)(10, 5, 0)))     # would shrink: refused
    out.append(lib.zdf_extend_dataset(C.byref(f), C.byref(ds), (C.c_uint64 * 3)(10, 8, 0)))
    extra = np.arange(20, dtype=np.float32) + 100
    ch = _ZChunk((C.c_uint64 * 3)(10, 2, 0), (C.c_uint64 * 3)(0, 6, 0), (C.c_uint64 * 3)(1, 1, 0), extra.ctypes.data)
    out += [lib.zdf_write_chunk_header(C.byref(f), C.byref(ds), C.byref(ch)),
            lib.zdf_vector_write(C.byref(f), extra.ctypes.data, T_FLOAT32, 20),
            lib.zdf_end_cdset(C.byref(f), C.byref(ds))]
    assert lib.zdf_close_file(C.byref(f)) == 1
    return out


def test_tracks_chunked_datasets_and_updates_identical_to_reference(ours, ref, tmp_path):
    paths = {k: str(tmp_path / (k + ".zdf")).encode() for k in ("ours", "ref")}
    a, b = _tracks_and_chunks(ours, paths["ours"]), _tracks_and_chunks(ref, paths["ref"])
    assert a == b
    assert filecmp.cmp(paths["ours"], paths["ref"], shallow=False)
    a, b = _grow(ours, paths["ours"]), _grow(ref, paths["ref"])
    assert a == b
    assert a[1] == 1 and a[6] == [10, 6] and a[7] == -1 and a[8] == 1
    assert filecmp.cmp(paths["ours"], paths["ref"], shallow=False)
    # the file is not a ZDF file / does not exist: both refuse
    junk = tmp_path / "junk.zdf"
    junk.write_bytes(b"not a zdf file")
    for lib in (ours, ref):
        f = _ZFile()
        assert lib.zdf_open_file(C.byref(f), str(junk).encode(), ZDF_READ) == 0
        assert lib.zdf_open_file(C.byref(f), str(tmp_path / "absent.zdf").encode(), ZDF_UPDATE) == 0


def test_reader_assembles_chunked_datasets(ours, tmp_path):
    _declare_zdf(ours)
    path = str(tmp_path / "tracks.zdf")
    f = _ZFile()
    assert ours.zdf_open_file(C.byref(f), path.encode(), ZDF_CREATE) == 1
    names = [(C.c_char_p * 2)(*v) for v in ((b"t", b"x1"), (b"t", b"x_1"), (b"1/w_p", b"c/w_p"))]
    tr = _ZTracks(b"electrons", b"two tracks", 2, 10, 1, 2, names[0], names[1], names[2])
    assert ours.zdf_add_track_info(C.byref(f), C.byref(tr)) > 0
    full = np.arange(60, dtype=np.float32).reshape(6, 10)
    ds = _ZDataset(b"data", T_FLOAT32, 2, (C.c_uint64 * 3)(10, 6, 0), None, 0, 0)
    assert ours.zdf_start_cdset(C.byref(f), C.byref(ds)) > 0
    # rows 0, 2, 4 as one strided chunk, then rows 1, 3, 5
    for j0 in (0, 1):
        part = np.ascontiguousarray(full[j0::2])
        ch = _ZChunk((C.c_uint64 * 3)(10, 3, 0), (C.c_uint64 * 3)(0, j0, 0), (C.c_uint64 * 3)(1, 2, 0), part.ctypes.data)
        assert ours.zdf_write_cdset(C.byref(f), C.byref(ds), C.byref(ch)) > 0
    flags = np.arange(6, dtype=np.uint8)
    bs = _ZDataset(b"flags", T_UINT8, 1, (C.c_uint64 * 3)(6, 0, 0), None, 0, 0)
    assert ours.zdf_start_cdset(C.byref(f), C.byref(bs)) > 0
    for k0 in (0, 3):       # 3-byte chunks: padded to 4 on disk
        part = np.ascontiguousarray(flags[k0:k0 + 3])
        ch = _ZChunk((C.c_uint64 * 3)(3, 0, 0), (C.c_uint64 * 3)(k0, 0, 0), (C.c_uint64 * 3)(1, 0, 0), part.ctypes.data)
        assert ours.zdf_write_cdset(C.byref(f), C.byref(bs), C.byref(ch)) > 0
    assert ours.zdf_end_cdset(C.byref(f), C.byref(ds)) > 0 and ours.zdf_end_cdset(C.byref(f), C.byref(bs)) > 0
    assert ours.zdf_close_file(C.byref(f)) == 1
    data, info = zdf.read(path)
    assert info.tracks.ntracks == 2 and info.tracks.quants == ["t", "x1"] and info.tracks.qunits == ["1/w_p", "c/w_p"]
    assert np.array_equal(data["data"], full)
    assert np.array_equal(data["flags"], flags)
    assert [k for k, _, _ in zdf.list_records(path)] == ["track_info", "cdset_start", "cdset_chunk", "cdset_chunk",
                                                        "cdset_start", "cdset_chunk", "cdset_chunk", "cdset_end", "cdset_end"]
