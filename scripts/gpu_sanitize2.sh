# compute-sanitizer memcheck of the code paths touched late in round 2: smoke, tile regrow (swapped overflow lists, kept scratch),
# the overlapped charge deposit, the field march kernel, small tiles with sized migrants segments
export PYTHONPATH=$PWD
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san2_mem.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|Invalid|smoke ok" gpurun_out/san2_mem.log | head -5
timeout 140 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_em2d.py -m gpu -q -x --no-gpu-retry -k "tiles_grow or charge_deposit or fast_beam or field_solver" > gpurun_out/san2_mem2.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/san2_mem2.log | head -6
