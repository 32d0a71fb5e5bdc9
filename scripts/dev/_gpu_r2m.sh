export PYTHONPATH=$PWD
echo "=== evaluate only (phase A of the list kernel), dense table"
QE_EVAL=1 QE_FORM=0 QE_SKIP=0 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
echo "=== atomics mode"
QE_ACC=1 QE_FORM=0 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
echo "=== lists form skip 40"
QE_FORM=0 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 3 2>&1 | tail -5
