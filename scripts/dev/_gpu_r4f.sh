export PYTHONPATH=$PWD
echo "=== form tests"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5)" 2>&1 | tail -5
echo "=== form 5 skip 40"
QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
echo "=== form 5 skip 256"
QE_FORM=5 QE_SKIP=256 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
echo "=== form 5 c4 single GPU"
QE_FORM=5 QE_SKIP=16 timeout 300 python scripts/perf_probe.py 1e8 8 4194304 8 3 2>&1 | tail -6
export QE_FORM=5 QE_SKIP=40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_flow_kernel -s 6 -c 1 -o gpurun_out/prof_flow_v3 python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r4f.log 2>&1
tail -2 gpurun_out/ncu_r4f.log
