# A/B round 4 of the low-ppc deposit: default = the scan carries its last run in tiles with >= 24 particles per cell;
# _rl1 = never, _c2 = in every tile
export PYTHONPATH=$PWD
for ppc in 4x8 4 2x4 8; do
  echo "== ppc $ppc"
  bash scripts/gpu_ab.sh "_rl1 _c2" 2048 $ppc 10 2
done
for v in "" _rl1 _c2; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done
python -m pytest tests/test_gpu_em2d.py -m gpu -q -x 2>&1 | tail -3
