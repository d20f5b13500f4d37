# round 2, call U: why is k_adj_own_3d slow -- plan statistics, launch list and one full ncu capture (1/8 of the spokes)
mkdir -p gpurun_out
timeout 600 python profiles/scripts/own3_probe.py > gpurun_out/r2u_probe.log 2>&1
grep -v Warn gpurun_out/r2u_probe.log | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches.csv python profiles/scripts/own3_probe.py 0.125 > gpurun_out/r2u_probe8.log 2>&1
grep -v Warn gpurun_out/r2u_probe8.log | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own_3d' -s 2 -c 1 -o gpurun_out/r2u_prof_own3 -f python profiles/scripts/own3_probe.py 0.125 > gpurun_out/r2u_prof.log 2>&1
tail -2 gpurun_out/r2u_prof.log | cut -c1-200
