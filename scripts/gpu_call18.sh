# A/B of the run-loop deposit (PUSH_RUN_LOOP_MAX = 1 is the old behaviour: one new cell per 32 lanes through the
# butterfly, more through the segmented scan), at 64 / 32 / 16 / 8 particles per cell; + the prefetching k_yee_march
export PYTHONPATH=$PWD
for ppc in 8 4x8 4 2x4; do
  echo "== ppc $ppc"
  bash scripts/gpu_ab.sh "_rl1 _rl2 _rl4" 2048 $ppc 10 2
done
python scripts/grid_probe.py 4096 50 | grep "fused=2"
for v in "" _rl1; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done
python -m pytest tests/test_gpu_em2d.py -m gpu -q -x 2>&1 | tail -3
