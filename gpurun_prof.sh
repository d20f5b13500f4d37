set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fwd_tiled_2d|k_adj_tiled_2d' -s 6 -c 2 -o gpurun_out/r1_prof_tiled -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof.log 2>&1
ls -la gpurun_out/
