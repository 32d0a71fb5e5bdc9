export PYTHONPATH=$PWD
echo "=== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== bench default (driver flags)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err; tail -2 gpurun_out/bench_r2_final_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final_n1.json') if l.startswith('{')][-1])
print('value %.3f G frac %.4f e2e %.3f G long %.3f late %.3f traffic %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['value_long']['value']/1e9, d['config']['late_training']['value']/1e9, d['roofline']['traffic']))
print(d['roofline']['phase_us_per_step'])
PY
