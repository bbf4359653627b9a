python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_digits_cache5|k_features_digits5|k_dw_splitk' -c 40 --csv --log-file gpurun_out/r2_launches_v11.csv python tools/ncu_target.py 65536 > gpurun_out/ncu_plain.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_v11.csv 12
