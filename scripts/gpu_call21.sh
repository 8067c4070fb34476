export PYTHONPATH=$PWD
O=gpurun_out
( time python bench.py --cpu-seconds 3 > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err ) 2>&1 | tail -3
tail -3 $O/r02c_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02c_bench_n1.json"))
print("value %.3e  ms/step %.2f  frac %.3f  e2e %.3e  launches %d  cells %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["cells"]["frac"]))
print(json.dumps(d.get("other_configs"), indent=1))
PY
