export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -25
