export PYTHONPATH=$PWD
echo "=== stats form 3 skip 40"
QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -3
echo "=== mb4 form 3 skip 40"
QE_LIBRARY=$PWD/build/libqe_mb4.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -6
export QE_FORM=3 QE_SKIP=40
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r2j.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_pipe_kernel -s 6 -c 1 -o gpurun_out/prof_pipe_v4 python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r2j.log 2>&1
tail -2 gpurun_out/ncu_r2j.log
