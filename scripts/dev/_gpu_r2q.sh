export PYTHONPATH=$PWD
nvidia-smi topo -m | head -12
echo "=== NCCL/IPC parity check, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -8
echo "=== shard bench C4, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_bench.py 2>&1 | tail -6
echo "=== shard bench C4, 1 rank (world 1)"
timeout 600 python scripts/shard_bench.py 2>&1 | tail -5
