export PYTHONPATH=$PWD
export QE_FORM=5 QE_SKIP=40
python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/plain_r5m.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_flow_kernel -s 6 -c 1 -f -o /tmp/prof_flow python scripts/perf_probe.py 1e6 16 1048576 8 1 > gpurun_out/ncu_r5m.log 2>&1
tail -2 gpurun_out/ncu_r5m.log
python scripts/ncu_summary.py /tmp/prof_flow.ncu-rep "round 2 final: one-pass pipeline, config 3, one launch of 8 vector steps after 48 steps" > gpurun_out/r2_final_flow_kernel_ncu_summary.json
ncu -i /tmp/prof_flow.ncu-rep --page source --csv --print-source sass,cuda > /tmp/flow_src.csv 2>/dev/null; python scripts/ncu_lines.py /tmp/flow_src.csv 70 > gpurun_out/r2_final_flow_kernel_hot_lines.txt
unset QE_FORM QE_SKIP
python bench.py --gpus 1 --steps 20 --warmup 5 --no-late > gpurun_out/plain_r5m_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_bench_launches_ncu.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-late > gpurun_out/ncu_r5m_list.log 2>&1
tail -1 gpurun_out/ncu_r5m_list.log | cut -c1-200
echo "=== c2 kernel ncu"
python bench.py --workload c2 --steps 512 --warmup 256 > gpurun_out/plain_r5m_c2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_small_kernel -s 1 -c 1 -f -o /tmp/prof_small python bench.py --workload c2 --steps 512 --warmup 256 > gpurun_out/ncu_r5m_c2.log 2>&1
tail -1 gpurun_out/ncu_r5m_c2.log | cut -c1-200
python scripts/ncu_summary.py /tmp/prof_small.ncu-rep "round 2 final: one-CTA loop, config 2 (TicTacToe, 128 agents), one launch of 256 vector steps" > gpurun_out/r2_final_small_kernel_ncu_summary.json
ncu -i /tmp/prof_small.ncu-rep --page source --csv --print-source sass,cuda > /tmp/small_src.csv 2>/dev/null; python scripts/ncu_lines.py /tmp/small_src.csv 50 > gpurun_out/r2_final_small_kernel_hot_lines.txt
ls -la gpurun_out
