# A/B round 2 of the low-ppc deposit: default = run loop up to 2 cells + scan that carries its last run;
# _rl1 = round-2 behaviour (one cell through the butterfly, scan flushes everything), _rl1c = _rl1 + carry, _rl3c = loop up to 3 + carry
export PYTHONPATH=$PWD
for ppc in 4x8 4 2x4 8; do
  echo "== ppc $ppc"
  bash scripts/gpu_ab.sh "_rl1 _rl1c _rl3c" 2048 $ppc 10 2
done
for v in "" _rl1 _rl1c; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done
python -m pytest tests/test_gpu_em2d.py -m gpu -q -x 2>&1 | tail -3
