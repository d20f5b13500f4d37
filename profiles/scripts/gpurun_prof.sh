mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_adj_tiled_3d|k_fwd_tiled_3d' -s 2 -c 2 -o gpurun_out/r1_prof_3d -f python profiles/run_cfg.py cfg4 2 > gpurun_out/r1_prof_3d.log 2>&1
tail -2 gpurun_out/r1_prof_3d.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv python profiles/run_cfg.py cfg4 2 2>/dev/null | grep -E '"k_|"void' | cut -d'"' -f10,30 | cut -c1-120 | tail -24
