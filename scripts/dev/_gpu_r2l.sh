export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "all_forms and 3- or chunking" 2>&1 | tail -5
for skip in 40 256; do
echo "=== form 3 skip $skip"
QE_FORM=3 QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
done
echo "=== stats form 3 skip 40"
QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -4
echo "=== mb4 form 3 skip 40"
QE_LIBRARY=$PWD/build/libqe_mb4.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
