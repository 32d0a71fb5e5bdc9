export PYTHONPATH=$PWD
export QE_FORM=5 QE_SKIP=40
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r5a.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_flow_kernel -s 6 -c 1 -o gpurun_out/prof_flow_r2final python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r5a.log 2>&1
tail -2 gpurun_out/ncu_r5a.log
unset QE_FORM QE_SKIP
python bench.py --gpus 1 --steps 20 --warmup 5 --no-late > gpurun_out/plain_r5a_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_bench_launches_ncu.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-late > gpurun_out/ncu_r5a_list.log 2>&1
tail -2 gpurun_out/ncu_r5a_list.log | cut -c1-300
