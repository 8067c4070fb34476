# round 2 verification + evidence on one GPU box: smoke, GPU tests, both bench arms, ncu captures
# (k_push2d --set full on the Weibel probe; launch list of the bench command; the grid / migrate kernels on the LWFA probe)
export PYTHONPATH=$PWD
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/r02_gputests.log 2>&1; tail -6 $O/r02_gputests.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; cut -c1-200 $O/r02_bench_ref.json
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -2 $O/r02_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print("value %.3e  ms/step %.2f  frac %.3f  e2e %.3e  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o $O/r02_push_v6c -f python scripts/quick_push_probe.py 1024 8 2 > $O/ncu_push_v6c.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
for k in k_yee_fused k_smooth_x k_smooth_y k_migrate2d k_fold_x; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o $O/r02_lwfa_$k -f python scripts/lwfa_probe.py 4096 1024 20 > $O/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 60 -c 2 -o $O/r02_lwfa_k_push2d -f python scripts/lwfa_probe.py 4096 1024 40 > $O/ncu_lwfa_push.log 2>&1
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
python bench.py --workload em1d 2>/dev/null | cut -c1-600 | tee $O/r02_bench_em1d.json
ls -la $O | tail -30
