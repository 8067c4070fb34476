export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_guard.py tests/test_cython_module.py tests/test_gpu_pyapi.py -m gpu -q 2>&1 | tail -3
python scripts/e2e_breakdown.py 4096 | grep -v "call [1-5]"
