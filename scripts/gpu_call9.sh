export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs_c.py tests/test_gpu_slabs.py tests/test_gpu_decks.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
python scripts/quick_push_probe.py 2048 8 5 | grep Gpush
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
echo "KH 4096: $(python bench.py --workload kh --kh-n 4096 --steps 20 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.2f ms/step %.2f Gpush/s" % (d["ms_per_step"], d["value"]/1e9))')"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_migrate2d -s 40 -c 6 --csv --log-file gpurun_out/mig_times.csv python scripts/lwfa_probe.py 4096 1024 40 > /dev/null 2>&1; grep -o '"[0-9.]*"$' gpurun_out/mig_times.csv | tr '\n' ' '
