"""Build the reference's own Cython modules (python/source/em2d.pyx, em1d.pyx - UNMODIFIED, compiled from where
they lie under /root/reference) against this repository's headers and link them to libzpic_b200_em{2,1}d.so
instead of the reference C files: the drop-in at the Python boundary.

    python -m zpic_b200.cython.build_modules            # -> zpic_b200/cython/_build/em2d.*.so, em1d.*.so

Nothing of the reference is copied into the tree: the generated C file lives in the git-ignored _build
directory only while it is compiled.  The modules' `cdef extern from "../../em2d/particles.h"` resolve to
include/em2d through a symlink, which works because the public structs have the reference layout.

The unmodified .pyx hands out views of the host buffers without telling the library, so run it with
ZPIC_COHERENT=1 (host mirrors refreshed around every sim_iter); with the three-line patch of INTEGRATION.md 2
the mirrors are synchronised on demand instead."""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
REF = os.environ.get("ZPIC_REFERENCE", "/root/reference")


def build(code="em2d", verbose=False):
    """returns the path of the built extension module, or None when the reference tree is not around"""
    pyx = os.path.join(REF, "python", "source", code + ".pyx")
    if not os.path.exists(pyx):
        return None
    import numpy
    from zpic_b200 import build as zbuild
    lib = zbuild.build(code)
    out = os.path.join(BUILD, code + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(pyx), os.path.getmtime(lib)):
        return out
    gen_dir = os.path.join(BUILD, "gen", code)            # two levels below _build: "../../em2d/x.h" -> _build/em2d/x.h
    os.makedirs(gen_dir, exist_ok=True)
    link = os.path.join(BUILD, code)
    if os.path.islink(link):
        os.remove(link)
    os.symlink(os.path.join(REPO, "include", code), link)
    c_file = os.path.join(gen_dir, code + ".c")
    # the reference pins Cython 0.29 (binder/requirements.txt); Cython 3 needs the implicit-noexcept behaviour back
    cmd = [sys.executable, "-m", "cython", "-3", "--directive", "legacy_implicit_noexcept=True", pyx, "-o", c_file]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("cython failed:\n" + r.stdout[-2000:])
    cc = [os.environ.get("CC", "gcc"), "-shared", "-fPIC", "-O2", "-std=gnu99", "-w",
          "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), c_file,
          "-L" + os.path.dirname(lib), "-l:" + os.path.basename(lib),
          "-Wl,-rpath,$ORIGIN/../../lib", "-o", out]
    r = subprocess.run(cc, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("compiling the Cython module failed:\n" + r.stdout[-2000:])
    shutil.rmtree(os.path.join(BUILD, "gen"), ignore_errors=True)      # generated from the reference: not kept
    if verbose:
        print(out)
    return out


def build_all(verbose=False):
    return [build(c, verbose) for c in ("em2d", "em1d")]


if __name__ == "__main__":
    for p in build_all(verbose=True):
        if p is None:
            print("reference tree not found (%s): nothing built" % REF)
