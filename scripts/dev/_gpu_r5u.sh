export PYTHONPATH=$PWD
QE_FORM=5 QE_SKIP=8 timeout 60 python scripts/perf_probe.py 1e6 16 1048576 8 1 2>&1 | grep "best" || { echo "FAILED OR HUNG"; exit 1; }
timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5) or config3_full or chunking" 2>&1 | tail -2
for sk in 40 256; do
QE_FORM=5 QE_SKIP=$sk timeout 90 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep "best\|phase A"
done
