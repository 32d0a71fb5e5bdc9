export PYTHONPATH=$PWD
echo "=== parity"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -3
echo "=== shard bench C4, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -5
echo "=== shard bench C4, 1 rank"
timeout 600 python scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -3
