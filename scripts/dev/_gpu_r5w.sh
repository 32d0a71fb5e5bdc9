export PYTHONPATH=$PWD
echo "=== pytest gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
