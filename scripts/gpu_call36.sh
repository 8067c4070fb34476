# the overlapped charge deposit + the regrow planner as a function: tests, timing of the host-buffer calls at 4096^2, LWFA line
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_full_size.py tests/test_gpu_pyapi.py -m gpu -q -x 2>&1 | tail -3
python scripts/e2e_breakdown.py 4096 | grep -E "deposit_charge|sync_emf" | tail -8
python bench.py --workload lwfa --steps 200 --warmup 5 2>/dev/null | cut -c1-200
