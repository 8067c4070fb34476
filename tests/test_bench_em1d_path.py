"""bench.py --workload em1d runs its whole body against a stand-in for the device library (call names and
argument counts are checked against the real library's declared prototypes; no device needed)."""
import ctypes as C
import json
import types

import bench
from zpic_b200 import load


class _Recorder:
    """answers every zdev_* call the em1d leg makes with a plausible value and checks it against the prototype"""

    def __init__(self, real):
        self.real, self.calls, self.n_launch = real, [], 0

    def __getattr__(self, name):
        fn = getattr(self.real, name)                    # AttributeError if the library has no such entry point
        argtypes = fn.argtypes

        def call(*args):
            assert argtypes is None or len(args) == len(argtypes), (name, len(args), len(argtypes))
            self.calls.append(name)
            if name in ("zdev_init", "zdev_spec1d_inject_lattice"):
                return 0
            if name.endswith("_create"):
                return 0x1000 + len(self.calls)
            if name == "zdev_launch_count":
                self.n_launch += 70
                return self.n_launch
            if name == "zdev_event_elapsed_ms":
                return 55.0
            if name == "zdev_spec1d_fetch":
                C.cast(args[1], C.POINTER(C.c_double))[0] = 1.0
                C.cast(args[2], C.POINTER(C.c_int64))[0] = (1 << 10) * 8
            return None
        return call


def test_em1d_leg_of_the_bench(capsys):
    rec = _Recorder(load("em1d"))
    args = types.SimpleNamespace(steps=2, warmup=3, log2_cells=10, ppc1d=8)
    out = bench.run_em1d(args, lib=rec)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["metric"] == bench.METRIC and line["n_gpus"] == 1 and line["steps"] == 2
    assert line["config"]["particles_per_gpu"] == 2 * 1024 * 8
    assert abs(line["value"] - 2 * 1024 * 8 * 2 / 55e-3) < 1e-6 * line["value"]
    assert abs(line["roofline"]["achieved"] - 40.0 * line["value"] / 1e9) < 1e-9
    assert rec.calls.count("zdev_spec1d_advance") == 2 * (3 + 2)
    assert out["gpu_launches"] == 70
