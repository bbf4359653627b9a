python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 4 --warmup 3 --no-fp64-compare --no-cpu-baseline > gpurun_out/r2_bench_v7.json 2> gpurun_out/r2_bench_v7.err; tail -c 300 gpurun_out/r2_bench_v7.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_v7.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],1), d['roofline']['pass_ms_last'], d['e2e']['value'], d['roofline']['avg_launch_ms_isolated'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v7.csv python tools/ncu_target.py 32768 > gpurun_out/ncu_plain.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_v7.csv 12
