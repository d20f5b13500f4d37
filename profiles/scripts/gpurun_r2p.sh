# round 2, call P: host profile of the end-to-end step
mkdir -p gpurun_out
timeout 600 python profiles/scripts/e2e_host.py cfg2 > gpurun_out/r2p_e2e_host.log 2>&1
head -60 gpurun_out/r2p_e2e_host.log | cut -c1-180
