# round 2, call 4b (1 GPU): peer all-reduce protocol on one device, streamed FFT option test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q -k "peer_allreduce" > gpurun_out/r4b_peer.log 2>&1
tail -5 gpurun_out/r4b_peer.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed_fft or fused_pruned" > gpurun_out/r4b_stream.log 2>&1
tail -3 gpurun_out/r4b_stream.log
