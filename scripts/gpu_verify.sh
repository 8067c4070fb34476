# round-end style verification on one GPU box: smoke, GPU tests, bench (both arms), profile of the bench kernel
export PYTHONPATH=$PWD
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_final.json"))
print("value %.3e  ms/step %.2f  frac %.3f  e2e %.3e  coherent %.3e  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["coherent"]["value"], d["gpu_launches"], d["clocks"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o gpurun_out/push_final -f python scripts/quick_push_probe.py 1024 8 2 > gpurun_out/ncu_final.log 2>&1
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
python scripts/quick_push_probe1d.py 22 256 5 | tail -1
python scripts/gpu_decks.py em2d | cut -c1-400
python scripts/gpu_decks.py em1d | cut -c1-400
