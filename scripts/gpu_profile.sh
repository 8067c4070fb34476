export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o gpurun_out/push_v5 -f python scripts/quick_push_probe.py 1024 8 2 > gpurun_out/ncu_v5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_v5.csv python bench.py --steps 2 --warmup 1 --grid 2048 --e2e-grid 256 --cpu-seconds 1 > gpurun_out/bench_under_ncu_v5.log 2>&1
python bench.py > gpurun_out/bench_v5b.json 2> gpurun_out/bench_v5b.err
cat gpurun_out/bench_v5b.json | cut -c1-400
