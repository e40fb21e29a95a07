python bench.py --steps 2 --warmup 1 --no_extras --no_cpu_baseline > gpurun_out/b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no_extras --no_cpu_baseline > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300
python profiles/prof_c5.py 1000 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tpt_search -c 1 -o gpurun_out/r2_tpt_v11 -f python profiles/prof_c5.py 1000 > gpurun_out/ncu_v11.log 2>&1
tail -3 gpurun_out/ncu_v11.log
