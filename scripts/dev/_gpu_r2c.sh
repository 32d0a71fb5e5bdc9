export PYTHONPATH=$PWD
for skip in 40; do
  echo "=== stats form 3 skip $skip"
  QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -7
done
