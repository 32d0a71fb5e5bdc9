export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "all_forms and 3-" 2>&1 | tail -15
for f in 3 0 1; do
  for skip in 40 256; do
    echo "=== form $f skip $skip"
    QE_FORM=$f QE_SKIP=$skip timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
  done
done
