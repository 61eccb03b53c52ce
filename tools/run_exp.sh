timeout 300 python tools/profile_step.py 256 3 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches_b256_final.csv python tools/profile_step.py 256 3 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
python tools/launch_table.py gpurun_out/r02_step_launches_b256_final.csv > gpurun_out/r02_step_launches_b256_final.md 2>&1; tail -3 gpurun_out/r02_step_launches_b256_final.md
