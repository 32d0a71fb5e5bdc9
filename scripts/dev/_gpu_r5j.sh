export PYTHONPATH=$PWD
echo "=== fullsize + capi tests"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_capi.py -q -x 2>&1 | tail -4
echo "=== skip 40"
QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
echo "=== skip 256"
QE_FORM=5 QE_SKIP=256 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
echo "=== c4 single GPU"
QE_FORM=5 QE_SKIP=16 timeout 300 python scripts/perf_probe.py 1e8 8 4194304 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
