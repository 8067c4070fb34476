export PYTHONPATH=$PWD
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
python scripts/lwfa_probe.py 4096 1024 200 | tail -1
