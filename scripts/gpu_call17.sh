# verification after k_yee_march + the host-copy fix: GPU tests, bench default line, ncu capture of k_yee_march
export PYTHONPATH=$PWD
O=gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 ) > $O/r02b_gputests.log 2>&1; tail -6 $O/r02b_gputests.log
python bench.py > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; tail -2 $O/r02b_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_n1.json"))
print("value %.3e  ms/step %.2f  frac %.3f  e2e %.3e  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cells", d.get("cells"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_yee_march -s 10 -c 1 -o $O/r02_k_yee_march -f python scripts/grid_probe.py 4096 5 > $O/ncu_yee_march.log 2>&1
tail -3 $O/ncu_yee_march.log
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
