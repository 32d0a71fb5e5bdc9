export PYTHONPATH=$PWD
echo "=== form tests"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5)" 2>&1 | tail -3
for sk in 40 256; do echo "=== skip $sk"; QE_FORM=5 QE_SKIP=$sk timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -9 | cut -c1-300 | grep -v "slow by\|slowest\|in-order"; done
for sk in 40 256; do echo "=== stats skip $sk"; QE_LIBRARY=$PWD/build/libqe_fstats.so QE_FORM=5 QE_SKIP=$sk timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -9 | cut -c1-300 | grep "in-order\|phase A"; done
