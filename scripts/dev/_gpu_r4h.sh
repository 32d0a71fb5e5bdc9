export PYTHONPATH=$PWD
for f in 0 8; do echo "=== flags $f"; QE_FLOW_FLAGS=$f QE_FORM=5 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -8 | cut -c1-330; done
