export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python scripts/quick_push_probe.py 2048 8 5 > gpurun_out/probe_v5e.txt 2>&1
echo "--- t128 16x8" >> gpurun_out/probe_v5e.txt
ZPIC_LIB_SUFFIX=_t128 python scripts/quick_push_probe.py 2048 8 5 >> gpurun_out/probe_v5e.txt 2>&1
echo "--- t128 8x8" >> gpurun_out/probe_v5e.txt
ZPIC_LIB_SUFFIX=_t128 ZPIC_TILE_X=8 ZPIC_TILE_Y=8 python scripts/quick_push_probe.py 2048 8 5 >> gpurun_out/probe_v5e.txt 2>&1
echo "--- t256 16x16" >> gpurun_out/probe_v5e.txt
ZPIC_TILE_X=16 ZPIC_TILE_Y=16 python scripts/quick_push_probe.py 2048 8 5 >> gpurun_out/probe_v5e.txt 2>&1
echo "--- t256 8x8" >> gpurun_out/probe_v5e.txt
ZPIC_TILE_X=8 ZPIC_TILE_Y=8 python scripts/quick_push_probe.py 2048 8 5 >> gpurun_out/probe_v5e.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o gpurun_out/push_v5e -f python scripts/quick_push_probe.py 1024 8 2 > gpurun_out/ncu_v5e.log 2>&1
grep -E "^---|Gpush|error|Error" gpurun_out/probe_v5e.txt
