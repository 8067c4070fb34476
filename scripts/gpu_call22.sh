export PYTHONPATH=$PWD
O=gpurun_out
ZPIC_VERBOSE=1 python bench.py --workload lwfa --steps 100 --warmup 5 > $O/lwfa1.json 2> $O/lwfa1.err
grep -c "found their" $O/lwfa1.err; grep "found their" $O/lwfa1.err | head -5 | cut -c1-250; cut -c1-200 $O/lwfa1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/lwfa1_launches.csv python bench.py --workload lwfa --steps 40 --warmup 5 > $O/lwfa1_under_ncu.log 2>&1
python scripts/launch_summary.py $O/lwfa1_launches.csv | head -30
