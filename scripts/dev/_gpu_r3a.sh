export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_tictactoe.py tests/test_gpu_runtimes.py tests/test_gpu_capi.py tests/test_gpu_eval_bench.py -q -x 2>&1 | tail -8
echo "=== c2 probe (TTT 128 agents) via bench"
timeout 900 python bench.py --workload c2 --steps 2048 --warmup 256 > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err; tail -3 gpurun_out/bench_r2_c2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_c2.json') if l.startswith('{')][-1])
print('c2 value %.2f M' % (d['value']/1e6), 'ms/step %.5f' % d['ms_per_step'], 'e2e %.1f k' % (d['e2e']['value']/1e3), d['roofline']['kernel'], d['config'].get('td_update_form'))
cb=d.get('cpu_baseline') or {}
print('cpu', cb.get('value'), (cb.get('multiprocessing') or {}).get('best'), (cb.get('c_port') or {}).get('value'))
PY
