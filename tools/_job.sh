python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py tests/test_gpu_reference_graph.py tests/test_gpu_reference_source.py -x -q 2>&1 | tail -8
python tools/tile_timing.py 327680 > gpurun_out/r2_tile_timing_2phase.txt 2>&1
SCFGP_COLLECTOR=1 python tools/tile_timing.py 327680 > gpurun_out/r2_tile_timing_2phase_collector.txt 2>&1
SCFGP_FUSED_FEATURES=0 python tools/tile_timing.py 327680 > gpurun_out/r2_tile_timing_2phase_unfused.txt 2>&1
python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_2phase.json 2> gpurun_out/r2_bench_2phase.err; tail -c 600 gpurun_out/r2_bench_2phase.err
SCFGP_EPILOGUE=1 python bench.py --steps 4 --warmup 3 --no-fp64-compare --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_1phase.json 2> gpurun_out/r2_bench_1phase.err
SCFGP_COLLECTOR=1 python bench.py --steps 4 --warmup 3 --no-fp64-compare --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_collector.json 2> gpurun_out/r2_bench_collector.err
