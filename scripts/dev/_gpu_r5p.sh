export PYTHONPATH=$PWD
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_final_n$N.json 2> gpurun_out/bench_r2_final_n$N.err
echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final_n$N.json') if l.startswith('{')][-1])
print('N=$N value %.3f G'%(d['value']/1e9), 'ms/step %.3f'%d['ms_per_step'], d['config']['td_update_form'], 'e2e %.3f G'%(d['e2e']['value']/1e9), d['multi_gpu_parity'], (d.get('sharded_c4') or {}).get('value'))
PY
if [ "$2" = "sweep" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 scripts/c5_sweep.py ${3:-64} 2>&1 | grep "^{" | cut -c1-200
fi
