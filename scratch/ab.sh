ncu --set full --clock-control none --import-source on -k regex:k_trace_spec -o gpurun_out/prof_r1_spec python scratch/prof_bounce.py 9 > gpurun_out/ncu_spec.log 2>&1
tail -2 gpurun_out/ncu_spec.log
