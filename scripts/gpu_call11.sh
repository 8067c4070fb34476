export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs_c.py tests/test_gpu_slabs.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_smooth -s 10 -c 4 --csv --log-file gpurun_out/sm_times.csv python scripts/lwfa_probe.py 4096 1024 20 > /dev/null 2>&1; grep -o 'k_smooth[a-z_]*\|"[0-9.]*"$' gpurun_out/sm_times.csv | tr '\n' ' '; echo
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
