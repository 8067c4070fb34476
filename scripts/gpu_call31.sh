# outgrown tiles first on a side stream: LWFA probe (default vs _sn = the build before this change), tests that exercise tile growth and slabs
export PYTHONPATH=$PWD
for r in 1 2; do for v in "" _sn; do echo "lwfa '$v': $(ZPIC_LIB_SUFFIX=$v python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done; done
python bench.py --workload lwfa --steps 200 --warmup 5 2>/dev/null | cut -c1-200
python -m pytest tests/test_gpu_em2d.py tests/test_gpu_slabs_c.py tests/test_gpu_decks.py -m gpu -q -x 2>&1 | tail -3
