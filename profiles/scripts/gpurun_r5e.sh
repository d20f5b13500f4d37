# round 2, final multi-GPU evidence: N GPUs = $1
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r5e_pytest_2gpus.log 2>&1
  tail -2 gpurun_out/r5e_pytest_2gpus.log
fi
EXTRA=""
if [ "$N" != "2" ]; then EXTRA="--no-reference-cuda"; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/r5e_bench_${N}gpu.log 2>&1
tail -1 gpurun_out/r5e_bench_${N}gpu.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('value', d['value'], 'ms', d['ms_per_step'], 'eager', d['other_launch_mode']['ms_per_step'], 'e2e', d['e2e']['value'])
p = d['partitions']
print('cfg5', p['cfg5_strong']['value'], p['cfg5_strong']['ms_per_step'])
c = p['coil_sharded']
print('coil', c['ms_per_step'], 'allreduce alone', c['allreduce_ms_alone'], 'sep kernel', (c.get('peer_memory_allreduce') or {}).get('ms_per_step_allreduce_as_separate_kernel'), 'nccl', c['nccl_allreduce']['ms_per_step'], c['nccl_allreduce']['allreduce_ms_alone'], 'bit-identical', (c.get('peer_memory_allreduce') or {}).get('bit_identical_on_all_ranks'), (c.get('peer_memory_allreduce') or {}).get('error'))
"
