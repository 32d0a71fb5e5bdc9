export PYTHONPATH=$PWD
echo "=== parity G=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -4
echo "=== shard bench C4, 8 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -12
echo "=== shard bench C4, 4 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -8
