export PYTHONPATH=$PWD
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-late > gpurun_out/bench_r2_x.json 2> gpurun_out/bench_r2_x.err; tail -2 gpurun_out/bench_r2_x.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_x.json') if l.startswith('{')][-1])
print('value %.3f G e2e %.3f G' % (d['value']/1e9, d['e2e']['value']/1e9), d['e2e_engine_rng'])
PY
timeout 200 python bench.py --workload c2 --steps 1024 --warmup 256 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 e2e %.1f k' % (d['e2e']['value']/1e3), d['e2e_engine_rng'])"
