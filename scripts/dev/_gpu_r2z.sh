export PYTHONPATH=$PWD
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25
