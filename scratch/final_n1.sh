timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.log 2>&1; tail -1 gpurun_out/bench_r1_final.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['path_tracing']['value'], d['path_tracing']['ms_per_image'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_final2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-path > gpurun_out/ncu_bench_final2.log 2>&1; wc -l gpurun_out/launches_bench_final2.csv
