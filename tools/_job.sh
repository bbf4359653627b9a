python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py tests/test_gpu_guard.py tests/test_gpu_reference_graph.py tests/test_gpu_reference_source.py tests/test_gpu_scaler.py -x -q 2>&1 | tail -4
python tools/scaler_objective_check.py 2>&1 | tail -8
python bench.py --steps 4 --warmup 3 --no-fp64-compare --no-cpu-baseline > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err; tail -c 300 gpurun_out/r2_bench_v5.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v5.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'], d['e2e']['value'])"
