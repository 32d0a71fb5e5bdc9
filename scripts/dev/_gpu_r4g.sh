export PYTHONPATH=$PWD
echo "=== form 5 skip 40 staged"
QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -6
echo "=== form 5 skip 40 unstaged"
QE_FLOW_FLAGS=1 QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -5
echo "=== form 5 skip 256 staged"
QE_FORM=5 QE_SKIP=256 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -6
echo "=== form 5 c4"
QE_FORM=5 QE_SKIP=16 timeout 300 python scripts/perf_probe.py 1e8 8 4194304 8 2 2>&1 | tail -6
echo "=== form tests"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5)" 2>&1 | tail -3
