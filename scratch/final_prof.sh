set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.log 2>&1 || exit 1
tail -1 gpurun_out/bench_r1e.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-path > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_spec2 --launch-skip 4 --launch-count 2 -o gpurun_out/prof_r1_spec2_final python scratch/prof_default.py > gpurun_out/ncu_spec2.log 2>&1
tail -2 gpurun_out/ncu_spec2.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.log 2>&1
tail -1 gpurun_out/bench_r1_reference.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
