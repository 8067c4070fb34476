"""bench.py's `other_configs` (the other BASELINE configurations measured by the default run as child processes):
what is started for N = 1 and N > 1, what is kept of a child's JSON line, and that a failing or hanging child
costs an error entry, not the main line.  CPU only: the children are replaced by canned results."""
import json
import subprocess
import types

import bench


def _args(no_extras=False):
    return types.SimpleNamespace(no_extras=no_extras)


def _line(workload, n_gpus):
    return json.dumps({"metric": bench.METRIC, "value": 1.5e10, "unit": bench.UNIT, "n_gpus": n_gpus, "steps": 10,
                       "ms_per_step": 2.0, "scaling": "strong", "config": {"workload": workload},
                       "roofline": {"frac": 0.4, "kernel": "whole step"}, "cells": {"value": 3.0e9}})


def test_children_of_a_single_gpu_run(monkeypatch):
    seen = []

    def fake_run(cmd, env=None, timeout=None, **kw):
        seen.append((cmd, env, timeout))
        w = cmd[cmd.index("--workload") + 1]
        return types.SimpleNamespace(returncode=0, stdout="noise\n" + _line(w, 1) + "\n", stderr="")

    monkeypatch.setattr(subprocess, "run", fake_run)
    out = bench.other_configs(0, 1, _args())
    assert list(out) == ["em1d", "lwfa"]
    assert out["lwfa"]["value"] == 1.5e10 and out["lwfa"]["roofline_frac"] == 0.4 and out["lwfa"]["cell_updates_per_s"] == 3.0e9
    for cmd, env, timeout in seen:
        assert env["ZPIC_BENCH_CHILD"] == "1" and env["ZPIC_JOB"].startswith("bx") and timeout >= 60
        assert cmd[cmd.index("--gpus") + 1] == "1"
    assert len({env["ZPIC_JOB"] for _, env, _ in seen}) == 2        # every child job has its own shared-memory segment


def test_children_of_a_slab_run_and_failures(monkeypatch):
    def fake_run(cmd, env=None, timeout=None, **kw):
        w = cmd[cmd.index("--workload") + 1]
        if w == "kh":
            raise subprocess.TimeoutExpired(cmd, timeout)
        return types.SimpleNamespace(returncode=3, stdout="", stderr="(*error*) something went wrong")

    monkeypatch.setattr(subprocess, "run", fake_run)
    out = bench.other_configs(0, 8, _args())
    assert list(out) == ["lwfa", "kh"]
    assert "exit code 3" in out["lwfa"]["error"] and "went wrong" in out["lwfa"]["error"]
    assert "no result within" in out["kh"]["error"]
    assert bench.other_configs(1, 8, _args()) is None               # only rank 0 reports
    assert bench.other_configs(0, 8, _args(no_extras=True)) is None
