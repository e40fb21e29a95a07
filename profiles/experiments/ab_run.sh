python profiles/prof_secondary.py > gpurun_out/sec_plain.log 2>&1 && ncu --set full --clock-control none -k regex:"search_kernel|rollout_kernel" -o gpurun_out/r2_secondary -f python profiles/prof_secondary.py > gpurun_out/ncu_sec.log 2>&1
tail -3 gpurun_out/ncu_sec.log; tail -2 gpurun_out/sec_plain.log
