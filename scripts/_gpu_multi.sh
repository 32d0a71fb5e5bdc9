export PYTHONPATH=$PWD
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 40 > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err
tail -c 2500 gpurun_out/bench_r1_n$N.json; tail -5 gpurun_out/bench_r1_n$N.err
