python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python tools/tile_timing.py 327680 > gpurun_out/r2_tile_timing_v3.txt 2>&1
python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; tail -c 600 gpurun_out/r2_bench_v3.err
python tools/envelope.py cfg4 gpurun_out/r2_envelope_cfg4.json > gpurun_out/r2_envelope_cfg4.txt 2>&1; tail -4 gpurun_out/r2_envelope_cfg4.txt
