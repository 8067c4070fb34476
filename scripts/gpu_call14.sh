export PYTHONPATH=$PWD
for f in 2 1; do
  for n in 1024 4096; do
    echo "fused=$f n=$n"; ZPIC_FUSED_YEE=$f python -m tests.full_size em2d $n 8 5 --device-init 2>&1 | tail -1 | cut -c1-600
  done
done
python -m pytest tests -m gpu -q --deselect "tests/test_gpu_full_size.py::test_properties_hold_at_any_size[em2d-4096-8-True]" 2>&1 | tail -8 | cut -c1-250
python scripts/grid_probe.py 2>&1 | tail -12
