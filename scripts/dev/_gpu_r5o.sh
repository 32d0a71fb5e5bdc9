export PYTHONPATH=$PWD
export QE_LIBRARY=$PWD/build/libqe_next.so
echo "=== variant: quick hang check"
QE_FORM=5 QE_SKIP=8 timeout 60 python scripts/perf_probe.py 1e6 16 1048576 8 1 2>&1 | grep "best" || { echo "VARIANT FAILED OR HUNG (flow)"; exit 1; }
timeout 60 python bench.py --workload c2 --steps 512 --warmup 256 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 value %.3f M us/step %.2f e2e %.1f k' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e3))" || { echo "VARIANT FAILED OR HUNG (small)"; exit 1; }
echo "=== variant: form tests + small batches"
timeout 200 python -m pytest tests/test_gpu_fullsize.py -q -x -k "(all_forms and 5) or automatic or (long_run and 5) or small_batches" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_tictactoe.py tests/test_gpu_runtimes.py -q -x 2>&1 | tail -2
for sk in 40 256 600; do
echo "=== variant skip $sk"
QE_FORM=5 QE_SKIP=$sk timeout 90 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep "best\|phase A"
done
echo "=== variant c1"
timeout 90 python bench.py --workload c1 --steps 2048 --warmup 256 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c1 value %.3f M us/step %.2f e2e %.1f k' % (d['value']/1e6, d['ms_per_step']*1e3, d['e2e']['value']/1e3))"
unset QE_LIBRARY
echo "=== committed skip 600"
QE_FORM=5 QE_SKIP=600 timeout 90 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | grep "best\|phase A"
