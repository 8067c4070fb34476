export PYTHONPATH=$PWD
O=gpurun_out
ZPIC_VERBOSE=1 python bench.py --workload lwfa --steps 100 --warmup 5 > $O/lwfa1.json 2> $O/lwfa1.err
grep "found their" $O/lwfa1.err | sed 's/.*slots/slots/' | cut -c1-250; cut -c1-200 $O/lwfa1.json
