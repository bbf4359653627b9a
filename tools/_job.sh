python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/phase_times.py 8192 emu5 2>&1 | tail -1
SCFGP_POTRF_LOOKAHEAD=0 python tools/phase_times.py 8192 emu5 2>&1 | tail -1
python tools/scaling_probe.py emu5 2>&1 | tail -12
