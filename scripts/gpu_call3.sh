export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_guard.py tests/test_cython_module.py -m gpu -q -x > gpurun_out/guard.log 2>&1
head -50 gpurun_out/guard.log | cut -c1-220
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_guard.py --deselect tests/test_cython_module.py > gpurun_out/gpu_all.log 2>&1
tail -5 gpurun_out/gpu_all.log | cut -c1-220
