export PYTHONPATH=$PWD
timeout 120 python scripts/dense_probe.py 100000 16 100000 10
timeout 120 python scripts/dense_probe.py 100000 16 100000 10
timeout 60 python scripts/perf_probe.py 1e5 16 100000 10 4 | tail -8
timeout 60 python scripts/perf_probe.py 1e8 8 4194304 10 2 2>&1 | tail -5
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
