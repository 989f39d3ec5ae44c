ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python tools/run_config.py c5 --spp 4 --li 0 --reps 1 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
