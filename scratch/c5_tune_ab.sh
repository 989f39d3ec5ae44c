for t in 1 2 3 4; do echo "== tune $t"; B200PT_2L_TUNE=$t python tools/run_config.py c5 --li 0 --crop 0 --reps 2 2>&1 | grep "^render" | tail -1; done
