export PYTHONPATH=$PWD
echo "=== form tests"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5)" 2>&1 | tail -3
echo "=== skip 40"
QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
echo "=== skip 256"
QE_FORM=5 QE_SKIP=256 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
echo "=== skip 600"
QE_FORM=5 QE_SKIP=600 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
