timeout 900 python -m pytest tests/test_render_gpu.py tests/test_spatial_lights.py tests/test_sobol.py -x -q 2>&1 | tail -3
for o in 0 1; do
  echo "== overlap $o"
  B200PT_OVERLAP=$o python tools/run_config.py c3 --li 0 --crop 0 --reps 4 2>&1 | grep "^render" | tail -2
  B200PT_OVERLAP=$o python tools/run_config.py c1 --li 0 --crop 0 --reps 6 2>&1 | grep "^render" | tail -2
done
