export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; echo "bench rc $?"; tail -3 gpurun_out/bench_r1_c.err; cat gpurun_out/bench_r1_c.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r1_ref.json 2>&1; cat gpurun_out/bench_r1_ref.json
