export PYTHONPATH=$PWD
echo "=== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== bench default (driver flags)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_v5.json 2> gpurun_out/bench_r2_v5.err; tail -3 gpurun_out/bench_r2_v5.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_v5.json') if l.startswith('{')][-1])
print('value %.3f G frac %.4f e2e %.3f G long %.3f late %.3f' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['value_long']['value']/1e9, d['config']['late_training']['value']/1e9))
print(d['roofline']['phase_us_per_step'], d['config']['td_update_form'], d['roofline']['kernel'])
PY
