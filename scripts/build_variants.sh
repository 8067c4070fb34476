#!/bin/bash
# CPU side of an A/B: build extra copies of the libraries with other compile-time settings next to the default
# ones (zpic_b200/lib/libzpic_b200_em2d<suffix>.so, ..._em1d<suffix>.so; selected at run time with
# ZPIC_LIB_SUFFIX), then restore the default objects.
# usage: scripts/build_variants.sh "_ps=-DPUSH_PRESORTED -DPUSH1_PRESORTED" "_h8=-DYF_H_N=8" ...
set -e
cd "$(dirname "$0")/.."
touch_and_build() {
python - <<'PY'
import os
from zpic_b200 import build
for f in ("zdev_spec2d.cu", "zdev_grid2d.cu", "zdev_spec1d.cu", "zdev_grid1d.cu"):
    os.utime(os.path.join(build.CSRC, "dev", f))
print(build.build("em2d"), build.build("em1d"))
PY
}
for spec in "$@"; do
  suffix="${spec%%=*}"; flags="${spec#*=}"
  ZPIC_LIB_SUFFIX="$suffix" ZPIC_NVCC_EXTRA="$flags" touch_and_build
done
touch_and_build        # the default build, last, so that zpic_b200/lib/obj holds its objects again
