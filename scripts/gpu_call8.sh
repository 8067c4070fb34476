export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_em1d.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-200
for r in 1 2; do
for v in "_v6a" ""; do
echo "em1d variant '$v' round $r: $(ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe1d.py 22 256 5 2>&1 | grep -E 'Gpush|rror' | cut -c1-90)"
done; done
