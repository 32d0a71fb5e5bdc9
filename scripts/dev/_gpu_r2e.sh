export PYTHONPATH=$PWD
export QE_FORM=3 QE_SKIP=40
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r2e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_pipe_kernel -s 6 -c 1 -o gpurun_out/prof_pipe_v2 python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r2e.log 2>&1
tail -3 gpurun_out/plain_r2e.log; tail -5 gpurun_out/ncu_r2e.log
