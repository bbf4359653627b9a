python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; tail -c 300 gpurun_out/r2_bench_final2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final2.json').read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],1), r['pass_ms_last'], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), 'iso', r['avg_launch_ms_isolated'], 'frac', r['frac'], r['frac_vs_measured_mma_stream'], 'traffic', r['traffic'], 'cpu', d['cpu_baseline']['value'], 'fp64', d['fp64_dmma_path']['ms_per_step'], d['fp64_dmma_path']['parity_of_emulated'], d['clocks'], d['gpu_launches'])"
