export PYTHONPATH=$PWD
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/shard_probe.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
