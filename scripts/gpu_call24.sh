export PYTHONPATH=$PWD
O=gpurun_out
ZPIC_VERBOSE=1 python bench.py --workload lwfa --steps 200 --warmup 5 > $O/lwfa1.json 2> $O/lwfa1.err
grep "found their" $O/lwfa1.err | sed 's/.*slots/slots/' | cut -c1-250; cut -c1-200 $O/lwfa1.json
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs_c.py tests/test_gpu_decks.py -m gpu -q -x 2>&1 | tail -3
