export PYTHONPATH=$PWD
for sk in 40 256; do echo "=== stats skip $sk"; QE_LIBRARY=$PWD/build/libqe_fstats.so QE_FORM=5 QE_SKIP=$sk timeout 300 python scripts/perf_probe.py 1e6 16 1048576 8 2 2>&1 | tail -9 | cut -c1-300 | grep -v "slow by\|slowest\|B1\|B2"; done
