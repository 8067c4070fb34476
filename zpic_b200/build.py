"""Build recipe for the native libraries (in-tree, sm_100a only).

    python -m zpic_b200.build            # build everything that is out of date
    python -m zpic_b200.build --force

Produces zpic_b200/lib/libzpic_b200_em2d.so (and ..._em1d.so): the device kernels
(csrc/dev/*.cu, nvcc) plus the host C layer that exports the reference's own API
(csrc/host/<code>/*.c, gcc with strict IEEE flags because host scalars feed the
bit-exact parity checks).  The two codes export the same symbol names
(sim_iter, spec_advance, ...) exactly like the reference's em1d/ and em2d/
directories do, hence one shared object per code.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
OBJDIR = os.path.join(ROOT, "lib", "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")

# --fmad=false: no contraction, so the device arithmetic follows the reference's
# expression trees bit for bit (IEEE sqrt/div are nvcc defaults).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC",
    "-I" + os.path.join(REPO, "include"), "-I" + os.path.join(CSRC, "dev"),
] + os.environ.get("ZPIC_NVCC_EXTRA", "").split()
# host C: strict IEEE, no contraction (same flags as the strict oracle build)
CC_FLAGS = ["-O2", "-std=gnu99", "-ffp-contract=off", "-fPIC", "-Wall", "-Wno-unused-result",
            "-I" + os.path.join(REPO, "include")]

CODES = {
    "em2d": {
        "dev": ["zdev_runtime.cu", "zdev_refrng.cu", "zdev_grid2d.cu", "zdev_spec2d.cu"],
        "host_dir": "em2d",
    },
    "em1d": {
        "dev": ["zdev_runtime.cu", "zdev_refrng.cu", "zdev_grid1d.cu", "zdev_spec1d.cu"],
        "host_dir": "em1d",
    },
}
HOST_COMMON = "common"      # random.c, timer.c, zdf.c: identical in every reference code directory


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def _headers():
    hs = []
    for d in (os.path.join(REPO, "include"), os.path.join(CSRC, "dev"), os.path.join(CSRC, "host")):
        for base, _, files in os.walk(d):
            hs += [os.path.join(base, f) for f in files if f.endswith((".h", ".cuh"))]
    return hs


def lib_path(code="em2d"):
    return os.path.join(LIBDIR, "libzpic_b200_%s%s.so" % (code, os.environ.get("ZPIC_LIB_SUFFIX", "")))


def build(code="em2d", force=False, verbose=False):
    """Compile one code's shared library; returns its path."""
    cfg = CODES[code]
    os.makedirs(os.path.join(OBJDIR, code), exist_ok=True)
    hdrs = _headers()
    objs = []
    for f in cfg["dev"]:
        src = os.path.join(CSRC, "dev", f)
        obj = os.path.join(OBJDIR, code, f.replace(".cu", ".o"))
        if force or _newer(obj, [src] + hdrs):
            out = _run([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
            if verbose:
                print(out)
        objs.append(obj)
    hsrc = []
    for d in (os.path.join(CSRC, "host", HOST_COMMON), os.path.join(CSRC, "host", cfg["host_dir"])):
        hsrc += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".c")] if os.path.isdir(d) else []
    for src in hsrc:
        f = os.path.basename(src)
        obj = os.path.join(OBJDIR, code, "host_" + f.replace(".c", ".o"))
        if force or _newer(obj, [src] + hdrs):
            _run([CC] + CC_FLAGS + ["-I" + os.path.join(REPO, "include", cfg["host_dir"]),
                                    "-I" + os.path.join(CSRC, "host"), "-c", src, "-o", obj])
        objs.append(obj)
    lib = lib_path(code)
    if force or _newer(lib, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
             ["-Xlinker", "-Bsymbolic", "-lm"])
    return lib


def build_all(force=False, verbose=False):
    return [build(c, force=force, verbose=verbose) for c in CODES]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
