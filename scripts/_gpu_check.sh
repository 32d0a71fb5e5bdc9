export PYTHONPATH=$PWD
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 8 > gpurun_out/bench_r1_g2.json 2> gpurun_out/bench_r1_g2.err; echo "bench rc $?"; tail -5 gpurun_out/bench_r1_g2.err; cat gpurun_out/bench_r1_g2.json
