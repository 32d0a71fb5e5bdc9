export PYTHONPATH=$PWD
build/microbench > gpurun_out/microbench_r2.json 2> gpurun_out/microbench_r2.err; cat gpurun_out/microbench_r2.json
for skip in 40 256; do
  echo "=== stats form 3 skip $skip"
  QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
done
