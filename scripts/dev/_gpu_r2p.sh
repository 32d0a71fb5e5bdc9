export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -k "sharded" 2>&1 | tail -15
