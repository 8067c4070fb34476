# evidence refresh for the final kernels: --set full capture of k_push2d (64 ppc and 16 ppc), launch list of the bench command
export PYTHONPATH=$PWD
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o $O/r02c_push -f python scripts/quick_push_probe.py 1024 8 2 > $O/ncu_push_r02c.log 2>&1; tail -2 $O/ncu_push_r02c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o $O/r02c_push_ppc16 -f python scripts/quick_push_probe.py 2048 4 2 > $O/ncu_push_r02c16.log 2>&1; tail -2 $O/ncu_push_r02c16.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --cpu-seconds 1 > $O/bench_under_ncu_r02c.log 2>&1
python scripts/launch_summary.py $O/r02c_bench_launches.csv | head -24
