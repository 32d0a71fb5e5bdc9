export PYTHONPATH=$PWD
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench default"
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -3 gpurun_out/bench_r2_n1.err
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py --steps 16 --warmup 8 --no-late > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_pipe_kernel -s 3 -c 1 -o gpurun_out/prof_pipe_r2 python bench.py --steps 16 --warmup 8 --no-late > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -5
