python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py tests/test_gpu_reference_graph.py -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_features_digits5' -c 16 --csv --log-file gpurun_out/r2_launches_v12.csv python tools/ncu_target.py 65536 > gpurun_out/ncu_plain.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_v12.csv 12
