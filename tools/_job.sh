python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-fp64-compare --no-cpu-baseline > gpurun_out/r2_bench_v9.json 2> gpurun_out/r2_bench_v9.err; tail -c 300 gpurun_out/r2_bench_v9.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v9.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"
ncu --set full --import-source on --clock-control none -k regex:'k_features_digits5|k_gram_i8|k_tgemm_i8|k_digits_cache5|k_backward_i8_2p|k_dw_splitk|k_seeds_rows' -c 18 -o gpurun_out/r2_prof_final python tools/ncu_target.py 32768 > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log
