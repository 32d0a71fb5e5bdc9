export PYTHONPATH=$PWD
export QE_FORM=5 QE_SKIP=40
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r4c.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_flow_kernel -s 6 -c 1 -o gpurun_out/prof_flow_v1 python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r4c.log 2>&1
tail -2 gpurun_out/ncu_r4c.log
