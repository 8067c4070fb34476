export PYTHONPATH=$PWD
python -m pytest tests/test_refstream.py -m gpu -q > gpurun_out/refstream.log 2>&1; tail -25 gpurun_out/refstream.log | cut -c1-200
python bench.py --workload em1d 2>gpurun_out/em1d.err | tee gpurun_out/r02_bench_em1d.json | cut -c1-900; tail -3 gpurun_out/em1d.err
python scripts/gpu_decks.py em2d > gpurun_out/r02_deck_program_em2d.json 2>&1; cut -c1-600 gpurun_out/r02_deck_program_em2d.json
python scripts/gpu_decks.py em1d > gpurun_out/r02_deck_program_em1d.json 2>&1; cut -c1-600 gpurun_out/r02_deck_program_em1d.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 30 -c 1 -o gpurun_out/r02_lwfa_k_push2d -f python scripts/lwfa_probe.py 4096 1024 40 > gpurun_out/ncu_lwfa_push.log 2>&1; tail -2 gpurun_out/ncu_lwfa_push.log
