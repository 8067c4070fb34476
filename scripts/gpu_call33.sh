export PYTHONPATH=$PWD
O=gpurun_out
python scripts/lwfa_probe.py 2048 1024 50 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file $O/lwfa2k_launches.csv python bench.py --workload lwfa --lwfa-nx 2048 --steps 150 --warmup 5 > $O/lwfa2k_under_ncu.log 2>&1
python scripts/launch_summary.py $O/lwfa2k_launches.csv | head -24
python - <<PY
import csv,re,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/lwfa2k_launches.csv") if l.startswith('"'))]
h=rows[0]; iK=h.index("Kernel Name"); iV=h.index("Metric Value"); iG=h.index("Grid Size")
big=[(r[iG], float(r[iV].replace(",",""))) for r in rows[1:] if "k_push2d" in r[iK]]
print("push launches (grid, ns):", big[-8:])
PY
