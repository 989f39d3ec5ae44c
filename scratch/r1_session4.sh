set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1f.log 2>&1; tail -1 gpurun_out/bench_r1f.log | cut -c1-600
python tools/run_config.py c3 --integrator whitted --li 2048 --crop 0.1 --json gpurun_out/r1_c3_whitted.json 2>&1 | tail -8
python tools/run_config.py c3 --integrator directlighting --strategy all --li 2048 --crop 0.1 --json gpurun_out/r1_c3_direct_all.json 2>&1 | tail -8
python tools/run_config.py c4 --li 0 --crop 0 --reps 2 --json gpurun_out/r1_c4_gpubuild.json 2>&1 | tail -6
