python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_arguments or feature_kernel" 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
