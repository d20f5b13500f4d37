# round 2, call 4a: streamed column passes A/B (cfg2, cfg3, cfg5)
mkdir -p gpurun_out
timeout 600 python profiles/scripts/stream_ab.py cfg2 cfg3 cfg5 0 1 2 3 17 34 51 > gpurun_out/r4a_stream_ab.log 2>&1
cat gpurun_out/r4a_stream_ab.log | tail -50
