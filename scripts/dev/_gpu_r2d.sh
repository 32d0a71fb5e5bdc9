export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "all_forms and 3-" 2>&1 | tail -5
for skip in 40 256; do
  echo "=== form 3 skip $skip"
  QE_FORM=3 QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
  echo "=== stats form 3 skip $skip"
  QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -7
done
echo "=== no sm roles"
QE_PIPE_NO_SM_ROLES=1 QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
