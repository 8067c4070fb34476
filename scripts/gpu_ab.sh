# A/B throughput of build variants (ZPIC_LIB_SUFFIX) of the em2d push, interleaved and repeated (run-to-run noise on
# a shared box is a few percent): usage gpu_ab.sh "suffix[:ENV=v,ENV=v] ..." [n ppc steps rounds]
export PYTHONPATH=$PWD
N=${2:-2048}; PPC=${3:-8}; STEPS=${4:-10}; ROUNDS=${5:-3}
for r in $(seq $ROUNDS); do
for spec in "" $1; do
  v="${spec%%:*}"; envs=""; [ "$spec" != "$v" ] && envs="$(echo "${spec#*:}" | tr ',' ' ')"
  if [ -f zpic_b200/lib/libzpic_b200_em2d$v.so ]; then
    echo "variant '$spec' round $r: $(env $envs ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe.py $N $PPC $STEPS 2>&1 | grep -E 'Gpush|rror' | cut -c1-60)"
  fi
done
done
