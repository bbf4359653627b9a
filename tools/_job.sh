python -m pytest tests/test_gpu_emu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-fp64-compare --no-cpu-baseline > gpurun_out/r2_bench_v8.json 2> gpurun_out/r2_bench_v8.err; tail -c 300 gpurun_out/r2_bench_v8.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v8.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pass_ms_last'])"
