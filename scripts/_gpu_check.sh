export PYTHONPATH=$PWD
free -g | head -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15
