export PYTHONPATH=$PWD
python scripts/quick_push_probe.py 2048 8 5 | grep Gpush
for t in "16 16" "16 8" "8 8"; do set -- $t
echo "KH 4096 tile $1x$2: $(ZPIC_TILE_X=$1 ZPIC_TILE_Y=$2 python bench.py --workload kh --kh-n 4096 --steps 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.2f ms/step %.2f Gpush/s" % (d["ms_per_step"], d["value"]/1e9))')"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv --log-file gpurun_out/r02_lwfa_launches.csv python scripts/lwfa_probe.py 4096 1024 40 > gpurun_out/lwfa_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_lwfa_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 150 --csv --log-file gpurun_out/r02_kh_launches.csv python bench.py --workload kh --kh-n 4096 --steps 10 > gpurun_out/kh_under_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_kh_launches.csv
