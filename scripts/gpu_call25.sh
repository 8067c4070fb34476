export PYTHONPATH=$PWD
O=gpurun_out
ZPIC_VERBOSE=1 python bench.py --workload lwfa --steps 200 --warmup 5 > $O/lwfa1.json 2> $O/lwfa1.err
grep -c "found their" $O/lwfa1.err; cut -c1-200 $O/lwfa1.json
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
