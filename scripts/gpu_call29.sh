# A/B: the one-boundary case of deposit32 straight-line (default) against the loop form of calls 18-20 (_loop)
export PYTHONPATH=$PWD
for ppc in 8 4x8 4 2x4; do
  echo "== ppc $ppc"
  bash scripts/gpu_ab.sh "_loop" 2048 $ppc 10 2
done
for v in "" _loop; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done
python -m pytest tests/test_gpu_em2d.py -m gpu -q -x 2>&1 | tail -3
