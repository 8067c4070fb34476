export PYTHONPATH=$PWD
python scripts/lwfa_probe.py 2048 1024 50 > /dev/null 2>&1
for r in 1 2; do
for t in "16 16" "16 8" "8 8"; do
  set -- $t
  echo "tile $1x$2: $(ZPIC_TILE_X=$1 ZPIC_TILE_Y=$2 python bench.py --workload lwfa --lwfa-nx 2048 --steps 200 --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"]/1e9)')"
done; done
for t in "16 16" "16 8"; do set -- $t; echo "probe tile $1x$2: $(ZPIC_TILE_X=$1 ZPIC_TILE_Y=$2 python scripts/lwfa_probe.py 4096 1024 200 | tail -1)"; done
