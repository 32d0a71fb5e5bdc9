export PYTHONPATH=$PWD
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== smoke"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench default (driver flags)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; tail -3 gpurun_out/bench_r2_final_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final_n1.json') if l.startswith('{')][-1])
print('value %.3f G frac %.4f e2e %.3f G long %.3f late %.3f' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['value_long']['value']/1e9, d['config']['late_training']['value']/1e9))
print(d['roofline']['phase_us_per_step'], d['config']['td_update_form'], d['roofline']['kernel'])
PY
for w in c2 c1; do
echo "=== bench $w"
timeout 600 python bench.py --workload $w --steps 2048 --warmup 256 > gpurun_out/bench_r2_final_$w.json 2> gpurun_out/bench_r2_final_$w.err; tail -2 gpurun_out/bench_r2_final_$w.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final_$w.json') if l.startswith('{')][-1])
print('$w value %.3f M us/step %.2f e2e %.1f k' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e3), d['roofline']['kernel'], 'cpu ref', d['cpu_baseline']['value'], (d['cpu_baseline'].get('multiprocessing') or {}).get('best'), (d['cpu_baseline'].get('c_port') or {}).get('value'))
PY
done
echo "=== bench c4 single GPU"
timeout 900 python bench.py --workload c4 --steps 16 --warmup 8 > gpurun_out/bench_r2_final_c4.json 2> gpurun_out/bench_r2_final_c4.err; tail -2 gpurun_out/bench_r2_final_c4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final_c4.json') if l.startswith('{')][-1])
print('c4 value %.3f G frac %.4f' % (d['value']/1e9, d['roofline']['frac']), d['roofline']['kernel'])
PY
