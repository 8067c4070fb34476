# 8-GPU confirmation run (one call): Weibel weak scaling line, LWFA (config 3), KH (config 4) at 8 and 4 ranks
export PYTHONPATH=$PWD
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 400 $T --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -2 gpurun_out/r02_bench_n8.err | cut -c1-300; cut -c1-330 gpurun_out/r02_bench_n8.json
timeout 400 $T --nproc-per-node 8 --master-port 29602 bench.py --workload lwfa --gpus 8 --steps 200 --warmup 5 > gpurun_out/r02_lwfa_n8.json 2> gpurun_out/r02_lwfa_n8.err; tail -2 gpurun_out/r02_lwfa_n8.err | cut -c1-300; cut -c1-330 gpurun_out/r02_lwfa_n8.json
timeout 400 $T --nproc-per-node 8 --master-port 29603 bench.py --workload kh --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_kh_n8.json 2> gpurun_out/r02_kh_n8.err; tail -2 gpurun_out/r02_kh_n8.err | cut -c1-300; cut -c1-330 gpurun_out/r02_kh_n8.json
