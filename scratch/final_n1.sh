set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference_b.log 2>&1; tail -1 gpurun_out/bench_r1_reference_b.log | cut -c1-200
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.log 2>&1; tail -1 gpurun_out/bench_r1_final.log | cut -c1-300
python tools/run_config.py c3 --li 2048 --crop 0.1 --json gpurun_out/r1_c3_final.json 2>&1 | grep "render\|parity\|crop"
python tools/run_config.py c1 --li 2048 --crop 1.0 --json gpurun_out/r1_c1_final.json 2>&1 | grep "render\|parity\|crop"
python tools/run_config.py c5 --li 1024 --crop 0.05 --json gpurun_out/r1_c5_final.json 2>&1 | grep "render\|parity\|crop\|preprocess"
