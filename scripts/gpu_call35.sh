# --set full captures of the current small kernels on the LWFA probe
export PYTHONPATH=$PWD
O=gpurun_out
for k in k_smooth_x_multi k_migrate2d; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -o $O/r02c_lwfa_$k -f python scripts/lwfa_probe.py 4096 1024 20 > $O/ncu_$k.log 2>&1; tail -1 $O/ncu_$k.log
done
