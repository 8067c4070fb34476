# First GPU call of the next round (after `scripts/build_variants.sh "_ps=-DPUSH_PRESORTED -DPUSH1_PRESORTED"` on the CPU side):
# parity of the default build, parity of the pre-sorted push variant, A/B throughput of both, the three modes of
# the field advance, the reference program both ways.  Roughly 90 s on the box.
export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
if [ -f zpic_b200/lib/libzpic_b200_em2d_ps.so ]; then
  ZPIC_LIB_SUFFIX=_ps python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs.py -m gpu -q -x 2>&1 | tail -3
  for v in "" _ps; do echo "--- variant '$v'"; ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe.py 2048 8 5 | tail -1; done | tee gpurun_out/probe_ps.txt
fi
if [ -f zpic_b200/lib/libzpic_b200_em1d_ps.so ]; then
  ZPIC_LIB_SUFFIX=_ps python -m pytest tests/test_gpu_em1d.py -m gpu -q -x 2>&1 | tail -3
  for v in "" _ps; do echo "--- em1d variant '$v'"; ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe1d.py 20 256 20 | tail -1; done | tee gpurun_out/probe1d_ps.txt
fi
python scripts/grid_probe.py 4096 50 | tee gpurun_out/grid_probe_modes.txt
python scripts/gpu_decks.py em2d | cut -c1-500 | tee gpurun_out/decks_em2d.txt
python scripts/gpu_decks.py em1d | cut -c1-500 | tee gpurun_out/decks_em1d.txt
