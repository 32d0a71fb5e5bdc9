export PYTHONPATH=$PWD
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 scripts/c5_sweep.py 32 2>&1 | grep "^{" | cut -c1-200
