mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r6c_cfg1_launches.csv python profiles/run_cfg.py cfg1 3 > gpurun_out/r6c_cfg1.log 2>&1
tail -1 gpurun_out/r6c_cfg1.log
