for b in 8 7; do echo "== 2l closest CTAs/SM $b"; B200PT_2L_BLOCKS=$b python tools/run_config.py c5 --li 0 --crop 0 --reps 2 2>&1 | grep "^render" | tail -1; done
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -k "instanc or c5 or nine" 2>&1 | tail -2
