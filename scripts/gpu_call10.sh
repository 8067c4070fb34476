export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs_c.py tests/test_gpu_slabs.py tests/test_gpu_decks.py tests/test_refstream.py tests/test_gpu_guard.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-250
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
ZPIC_FUSED_SMOOTH=0 python scripts/lwfa_probe.py 4096 1024 200 | tail -1
