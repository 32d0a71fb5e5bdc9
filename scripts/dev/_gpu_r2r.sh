export PYTHONPATH=$PWD
echo "=== shard bench C4, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -6
echo "=== shard bench C4, 1 rank (world 1)"
timeout 600 python scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -4
