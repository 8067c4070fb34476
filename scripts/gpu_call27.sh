# 8 GPUs: the default bench under torchrun (weak-scaling Weibel + slab parity + other_configs = LWFA and KH as children)
export PYTHONPATH=$PWD
O=gpurun_out
nvidia-smi -L | wc -l
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02c_bench_n8.json 2> $O/r02c_bench_n8.err ) 2>&1 | tail -3
tail -2 $O/r02c_bench_n8.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02c_bench_n8.json").read().strip().splitlines()[-1])
print("value %.3e  ms/step %.2f  frac %.3f  share %.3f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], d["slab_parity"]["ok"]))
print(json.dumps(d.get("other_configs"), indent=1))
PY
