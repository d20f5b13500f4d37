# cfg4 launch list (forward + adjoint NUFFT, one repetition after a warm-up one)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5b_cfg4_launches.csv python profiles/run_cfg.py cfg4 2 > gpurun_out/r5b_cfg4.log 2>&1
tail -2 gpurun_out/r5b_cfg4.log
