export PYTHONPATH=$PWD
O=gpurun_out
python scripts/lwfa_probe.py 4096 1024 50 > /dev/null 2>&1   # (the first probe of a call runs slow: discarded)
ZPIC_VERBOSE=1 python bench.py --workload lwfa --lwfa-nx 2048 --steps 200 --warmup 5 > $O/lwfa2k.json 2> $O/lwfa2k.err
grep -c "found their" $O/lwfa2k.err; grep "found their" $O/lwfa2k.err | sed 's/.*slots/slots/' | head -30 | cut -c1-200; cut -c1-220 $O/lwfa2k.json
