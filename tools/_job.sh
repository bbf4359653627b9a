python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py tests/test_gpu_guard.py tests/test_gpu_reference_graph.py tests/test_gpu_reference_source.py -x -q 2>&1 | tail -8
python tools/tile_timing.py 327680 > gpurun_out/r2_tile_timing_v4.txt 2>&1
python bench.py --steps 4 --warmup 3 --no-fp64-compare --no-cpu-baseline > gpurun_out/r2_bench_v4.json 2> gpurun_out/r2_bench_v4.err; tail -c 600 gpurun_out/r2_bench_v4.err
python tools/trace_eval.py 327680 emu5 gpurun_out/r2_trace_v4.json > gpurun_out/r2_trace_v4_summary.txt 2>&1
