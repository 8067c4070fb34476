export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_guard.py tests/test_refstream.py tests/test_gpu_em1d.py tests/test_gpu_pyapi.py -m gpu -q > gpurun_out/guard.log 2>&1
head -60 gpurun_out/guard.log | cut -c1-220
