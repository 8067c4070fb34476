"""Reader for ZDF files (the container every ZPIC diagnostic is written in; format: SURVEY.md App. C,
reference em2d/zdf.c:78-90, 769-1268).  Own implementation with the calling convention of the
reference's python/lib/zdf.py: `read(path) -> (data, info)`.

Supported records: int32, double, string, iteration, grid_info, part_info, dataset - everything the
em1d / em2d codes write."""
import struct
from types import SimpleNamespace

import numpy as np

_REC = {0x0001: "int32", 0x0002: "double", 0x0003: "string", 0x0010: "dataset",
        0x0020: "iteration", 0x0021: "grid_info", 0x0022: "part_info"}
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
    while c.p < len(buf):
        rid = c.take("I")
        name = c.string()
        length = c.take("Q")
        start = c.p
        yield _REC.get(rid >> 16, "unknown"), rid & 0xffff, name, c, length
        # datasets pad their payload to 4 bytes
        c.p = start + ((length + 3) & ~3)


def read(path):
    """-> (data, info): grid files give an ndarray shaped (ny, nx) [x fastest]; particle files a dict
    quantity -> 1-D array.  info has .type, .grid / .particles, .iteration like the reference reader."""
    with open(path, "rb") as f:
        buf = f.read()
    info = SimpleNamespace(type=None, grid=None, particles=None, iteration=None, extra={})
    datasets = {}
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
