export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_tictactoe.py tests/test_gpu_runtimes.py tests/test_gpu_capi.py -q -x -k "small or ttt or tictactoe or runtime or all_forms" 2>&1 | tail -4
echo "=== small MDP probe 128 agents"
timeout 300 python scripts/perf_probe.py 19683 9 128 256 3 2>&1 | grep -E "rep|grid"
echo "=== 1 agent"
timeout 300 python scripts/perf_probe.py 19683 9 1 256 3 2>&1 | grep -E "rep 2|grid"
echo "=== c2 via bench"
timeout 900 python bench.py --workload c2 --steps 2048 --warmup 256 > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err; tail -3 gpurun_out/bench_r2_c2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_c2.json') if l.startswith('{')][-1])
print('c2 value %.2f M' % (d['value']/1e6), 'us/step %.2f' % (d['ms_per_step']*1e3), 'e2e %.1f k' % (d['e2e']['value']/1e3), d['roofline']['kernel'])
PY
