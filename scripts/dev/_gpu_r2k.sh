export PYTHONPATH=$PWD
QE_LIBRARY=$PWD/build/libqe_stats.so QE_FORM=3 QE_SKIP=40 timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -8
