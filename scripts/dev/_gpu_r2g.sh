export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "all_forms and 3-" 2>&1 | tail -5
echo "=== form 3 skip 40"
QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
echo "=== stats form 3 skip 40"
QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -7
echo "=== mb3 form 3 skip 40 / 256"
QE_LIBRARY=$PWD/build/libqe_mb3.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
QE_LIBRARY=$PWD/build/libqe_mb3.so QE_FORM=3 QE_SKIP=256 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
export QE_FORM=3 QE_SKIP=40 QE_LIBRARY=$PWD/build/libqe_mb3.so
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r2g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_pipe_kernel -s 6 -c 1 -o gpurun_out/prof_pipe_v3 python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r2g.log 2>&1
tail -2 gpurun_out/ncu_r2g.log
