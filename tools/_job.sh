python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SCFGP_PHI_CACHE=0 python bench.py --rows 2000000 --steps 3 --warmup 2 --no-fp64-compare --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_nocache_2M.json 2> gpurun_out/r2_bench_nocache_2M.err; tail -c 200 gpurun_out/r2_bench_nocache_2M.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_nocache_2M.json').read().strip().splitlines()[-1]); print('nocache new', round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'])"
SCFGP_PHI_CACHE=0 SCFGP_FUSED_FEATURES=0 python bench.py --rows 2000000 --steps 3 --warmup 2 --no-fp64-compare --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_nocache_old_2M.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_nocache_old_2M.json').read().strip().splitlines()[-1]); print('nocache old path', round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'])"
python bench.py --rows 2000000 --steps 3 --warmup 2 --no-fp64-compare --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_cache_2M.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_cache_2M.json').read().strip().splitlines()[-1]); print('cache', round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'])"
