ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trace -o gpurun_out/prof_r1_spec2_final python tools/prof_traversal.py > gpurun_out/ncu_spec2.log 2>&1
tail -3 gpurun_out/ncu_spec2.log
