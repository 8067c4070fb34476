# A/B throughput of build variants (ZPIC_LIB_SUFFIX) of the em2d push: usage gpu_ab.sh "suffix suffix ..." [n ppc steps]
export PYTHONPATH=$PWD
N=${2:-2048}; PPC=${3:-8}; STEPS=${4:-5}
for v in "" $1; do
  if [ -f zpic_b200/lib/libzpic_b200_em2d$v.so ]; then
    echo "--- variant '$v'"
    ZPIC_LIB_SUFFIX=$v python scripts/quick_push_probe.py $N $PPC $STEPS 2>&1 | grep -E "Gpush|rror|tile"
  fi
done
