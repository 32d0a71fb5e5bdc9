export PYTHONPATH=$PWD
echo "=== one-GPU sharded parity tests (virtual ranks)"
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_multi.py -q -x 2>&1 | tail -4
echo "=== parity NCCL 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -3
echo "=== shard bench C4, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -5
echo "=== shard bench C4, 1 rank"
timeout 600 python scripts/shard_bench.py 1e8 8 4194304 8 2>&1 | tail -3
echo "=== c4 single GPU, one-pass form"
QE_FORM=5 QE_SKIP=16 timeout 300 python scripts/perf_probe.py 1e8 8 4194304 8 3 2>&1 | grep -v "slow by\|slowest\|in-order pass" | tail -6
