export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -x -k "every_bucket_sort_path" 2>&1 | tail -15
