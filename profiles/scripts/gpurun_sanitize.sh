# round 2: compute-sanitizer over the new kernels (owner-tile spreads 2-D / cell / 3-D, pre-pass, fix-up, visit-list builders)
mkdir -p gpurun_out
SEL="cases_match_reference or integer_indices or modified_tables or ordered_tiled or (tiled_kernels_match and (16-20 or 18-27 or 64-19)) or graph_replay"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cases_match_reference and (d2_edge or d2_radial or d3-)" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -4 gpurun_out/r02_sanitizer_racecheck.log
