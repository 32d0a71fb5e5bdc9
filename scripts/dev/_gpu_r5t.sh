export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_multi.py -q -x 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/replicated_profile.py 2>&1 | grep " ms"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -3
