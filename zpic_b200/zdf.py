"""Reader for ZDF files (the container every ZPIC diagnostic is written in; format: SURVEY.md App. C,
reference em2d/zdf.c:78-90, 735-1409).  Own implementation with the calling convention of the
reference's python/lib/zdf.py: `read(path) -> (data, info)`.

Supported records: int32, double, string, iteration, grid_info, part_info, dataset - everything the
em1d / em2d codes write - plus track_info and chunked datasets (start / chunk / end records, assembled
into one array) of the rest of the zdf.h API."""
import struct
from types import SimpleNamespace

import numpy as np

_REC = {0x0001: "int32", 0x0002: "double", 0x0003: "string", 0x0010: "dataset",
        0x0011: "cdset_start", 0x0012: "cdset_chunk", 0x0013: "cdset_end",
        0x0020: "iteration", 0x0021: "grid_info", 0x0022: "part_info", 0x0023: "track_info"}
_DTYPES = {1: "i1", 2: "u1", 3: "<i2", 4: "<u2", 5: "<i4", 6: "<u4", 7: "<i8", 8: "<u8", 9: "<f4", 10: "<f8"}


class _Cursor:
    def __init__(self, buf):
        self.b, self.p = buf, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def string(self):
        n = self.take("I")
        s = self.b[self.p:self.p + n].decode("utf-8", "replace")
        self.p += (n + 3) & ~3
        return s


def _records(buf):
    if buf[:4] != b"ZDF1":
        raise ValueError("not a ZDF file")
    c = _Cursor(buf)
    c.p = 4
    wide = {}                                 # chunked dataset id -> element size, for the padding rule below
    while c.p < len(buf):
        rid = c.take("I")
        name = c.string()
        length = c.take("Q")
        start = c.p
        kind = _REC.get(rid >> 16, "unknown")
        yield kind, rid & 0xffff, name, c, length
        # 8-bit vectors are padded to 4 bytes on disk and the padding is NOT part of the record length
        # (reference zdf.c:709-725); nothing else is padded
        end = start + length
        if kind in ("dataset", "cdset_start"):
            did, code, ndims = struct.unpack_from("<IiI", buf, start)
            if kind == "cdset_start":
                wide[did] = np.dtype(_DTYPES[code]).itemsize
            elif np.dtype(_DTYPES[code]).itemsize == 1:
                end = start + ((length + 3) & ~3)             # the dataset header is a multiple of 4 bytes
        elif kind == "cdset_chunk" and wide.get(struct.unpack_from("<I", buf, start)[0]) == 1:
            end = start + ((length + 3) & ~3)
        c.p = end


def read(path):
    """-> (data, info): grid files give an ndarray shaped (ny, nx) [x fastest]; particle files a dict
    quantity -> 1-D array.  info has .type, .grid / .particles, .iteration like the reference reader."""
    with open(path, "rb") as f:
        buf = f.read()
    info = SimpleNamespace(type=None, grid=None, particles=None, tracks=None, iteration=None, extra={})
    datasets = {}
    chunked = {}                              # dataset id -> (name, dtype, ndims) of the open chunked datasets
    for kind, version, name, c, length in _records(buf):
        if kind == "string":
            v = c.string()
            if name == "TYPE":
                info.type = v
            else:
                info.extra[name] = v
        elif kind == "int32":
            info.extra[name] = c.take("i")
        elif kind == "double":
            info.extra[name] = c.take("d")
        elif kind == "iteration":
            info.iteration = SimpleNamespace(name=name, n=c.take("i"), t=c.take("d"), tunits=c.string())
        elif kind == "grid_info":
            ndims = c.take("I")
            nx = [c.take("Q") for _ in range(ndims)]
            g = SimpleNamespace(name=name, ndims=ndims, nx=nx, label=c.string(), units=c.string(), axis=[])
            g.has_axis = c.take("i")
            if g.has_axis:
                for _ in range(ndims):
                    ax = SimpleNamespace(name=c.string(), type=c.take("i"), min=c.take("d"), max=c.take("d"))
                    ax.label, ax.units = c.string(), c.string()
                    g.axis.append(ax)
            info.grid = g
        elif kind == "part_info":
            p = SimpleNamespace(name=name, label=c.string(), nparts=c.take("Q"), nquants=c.take("I"))
            p.quants = [c.string() for _ in range(p.nquants)]
            p.qlabels = [c.string() for _ in range(p.nquants)]
            p.qunits = [c.string() for _ in range(p.nquants)]
            info.particles = p
        elif kind == "track_info":
            t = SimpleNamespace(name=name, label=c.string(), ntracks=c.take("I"), ndump=c.take("I"),
                                niter=c.take("I"), nquants=c.take("I"))
            t.quants = [c.string() for _ in range(t.nquants)]
            t.qlabels = [c.string() for _ in range(t.nquants)]
            t.qunits = [c.string() for _ in range(t.nquants)]
            info.tracks = t
        elif kind == "cdset_start":
            did = c.take("I")
            dt = _DTYPES[c.take("i")]
            ndims = c.take("I")
            count = [c.take("Q") for _ in range(ndims)]
            datasets[name] = np.zeros(count[::-1], dtype=dt)
            chunked[did] = (name, dt, ndims)
        elif kind == "cdset_chunk":
            did = c.take("I")
            cname, dt, ndims = chunked[did]
            count = [c.take("Q") for _ in range(ndims)]
            first = [c.take("Q") for _ in range(ndims)]
            stride = [c.take("Q") for _ in range(ndims)]
            n = int(np.prod(count))
            piece = np.frombuffer(buf, dtype=dt, count=n, offset=c.p).reshape(count[::-1])
            where = tuple(slice(s0, s0 + k * st, st) for s0, k, st in zip(first[::-1], count[::-1], stride[::-1]))
            datasets[cname][where] = piece
        elif kind == "cdset_end":
            pass
        elif kind == "dataset":
            c.take("I")                       # id
            dt = _DTYPES[c.take("i")]
            ndims = c.take("I")
            count = [c.take("Q") for _ in range(ndims)]
            n = int(np.prod(count)) if count else 0
            arr = np.frombuffer(buf, dtype=dt, count=n, offset=c.p).copy()
            datasets[name] = arr.reshape(count[::-1]) if ndims > 1 else arr
    if info.type == "grid":
        return datasets.get(info.grid.name), info
    if info.type == "particles":
        return datasets, info
    return datasets, info


def list_records(path):
    with open(path, "rb") as f:
        buf = f.read()
    return [(kind, name, length) for kind, version, name, c, length in _records(buf)]
