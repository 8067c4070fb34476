# one GPU iteration of kernel development: parity tests, throughput probe of the build variants, ncu capture
# usage: gpu_iter.sh TAG "variant-suffixes" NCU(0/1) ["ENV=.. ENV=.." extra env for an additional probe line per variant]
export PYTHONPATH=$PWD
TAG=${1:-v5}
VARIANTS=${2:-""}
NCU=${3:-1}
EXTRA=${4:-""}
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
out=gpurun_out/probe_$TAG.txt
: > $out
for v in "" $VARIANTS; do
  if [ -f zpic_b200/lib/libzpic_b200_em2d$v.so ]; then
    echo "--- variant '$v'" >> $out
    ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe.py 2048 8 5 >> $out 2>&1
    if [ -n "$EXTRA" ]; then
      echo "--- variant '$v' $EXTRA" >> $out
      env $EXTRA ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe.py 2048 8 5 >> $out 2>&1
    fi
  fi
done
if [ "$NCU" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2d -s 2 -c 1 -o gpurun_out/push_$TAG -f python scripts/quick_push_probe.py 1024 8 2 > gpurun_out/ncu_$TAG.log 2>&1
fi
grep -E "^---|Gpush|rror" $out
