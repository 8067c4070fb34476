# final verification of the round: smoke, all GPU tests, the default bench line (with other_configs)
export PYTHONPATH=$PWD
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-250 ) > $O/r02d_gputests.txt 2>&1; tail -6 $O/r02d_gputests.txt
python bench.py > $O/r02d_bench_n1.json 2> $O/r02d_bench_n1.err; tail -2 $O/r02d_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_n1.json"))
print("value %.3e  ms/step %.2f  frac %.3f  e2e %.3e  launches %d  cells %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["cells"]["frac"], d["clocks"]))
print({k: (v.get("value"), v.get("ms_per_step"), v.get("roofline_frac"), v.get("error")) for k, v in d["other_configs"].items()})
PY
