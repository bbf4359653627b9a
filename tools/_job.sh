python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/phase_times.py 8192 emu5 2>&1 | tail -1
