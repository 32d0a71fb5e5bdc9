export PYTHONPATH=$PWD
QE_FORM=5 QE_SKIP=8 timeout 60 python scripts/perf_probe.py 1e6 16 1024 8 2 2>&1 | grep "best" || { echo "FAILED OR HUNG"; exit 1; }
for n in 8192 65536 262144; do
QE_FORM=5 QE_SKIP=8 timeout 60 python scripts/perf_probe.py 1e6 16 $n 8 2 2>&1 | grep "best"
done
echo "=== tests"
timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_capi.py tests/test_gpu_runtimes.py tests/test_gpu_distributed.py -q -x 2>&1 | tail -3
