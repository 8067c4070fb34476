export PYTHONPATH=$PWD
python -m pytest tests/test_refstream.py -m gpu -q -x > gpurun_out/refstream.log 2>&1
head -60 gpurun_out/refstream.log
