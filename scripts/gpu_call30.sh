# why is the straight-line one-boundary path slower on the LWFA probe?  '' = straight + per-tile carry, _sn = straight, never carry, _loop = loop form + per-tile carry
export PYTHONPATH=$PWD
O=gpurun_out
for r in 1 2; do for v in "" _sn _loop; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done; done
for v in "" _loop; do
ZPIC_LIB_SUFFIX=$v timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:k_push2d -s 150 -c 4 --csv --log-file $O/lwfa_push_inst$v.csv python scripts/lwfa_probe.py 4096 1024 100 > /dev/null 2>&1
grep -E "k_push2d" $O/lwfa_push_inst$v.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-150
done
