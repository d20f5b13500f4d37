# round 2, sanitizer over the kernels added in this session: streamed column pass, peer all-reduce (single rank: no
# cross-kernel waits, which a serialising tool would turn into a deadlock), fix-up kernel (3-D, one tile per CTA)
mkdir -p gpurun_out
SEL="streamed_fft or (peer_allreduce_protocol and 1] ) or (cases_match_reference and (d2_edge or d3-))"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "streamed_fft or peer_allreduce_protocol_on_one_device[1] or (cases_match_reference and (d2_edge or d3-))" > gpurun_out/r5d_sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r5d_sanitizer_memcheck.log
tail -4 gpurun_out/r5d_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed_fft and (N2 or N5)" > gpurun_out/r5d_sanitizer_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/r5d_sanitizer_racecheck.log
tail -4 gpurun_out/r5d_sanitizer_racecheck.log
