# round 2, call 3a: persistent column FFT pass with register prefetch (B2N_OPT_FFT_PERSIST) A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned or toeplitz" > gpurun_out/r3a_pytest0.log 2>&1
tail -2 gpurun_out/r3a_pytest0.log
for opt in 0 1; do
  for wl in cfg2 cfg3; do
    timeout 300 python bench.py --steps 30 --warmup 5 --workload $wl --opt 9=$opt --no-cpu-baseline --no-reference-cuda --no-partitions > gpurun_out/r3a_${opt}_$wl.log 2>&1
    tail -1 gpurun_out/r3a_${opt}_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('persist=$opt', '$wl', 'step %.1f us' % (d['ms_per_step']*1e3), {k: round(v*1e3,1) for k,v in d['stages_ms'].items()})"
  done
done 2>&1 | tee gpurun_out/r3a_fft_persist_ab.log
B2N_TEST_OPT=9=1 timeout 900 python -c "
import sys; sys.path.insert(0,'.')
from torchkbnufft_b200 import _lib
_lib.load().b2n_set_option(9, 1)
import pytest
sys.exit(pytest.main(['tests/test_gpu_parity.py','-m','gpu','-x','-q','-k','fused_pruned or toeplitz or cases_match']))
" > gpurun_out/r3a_pytest1.log 2>&1
tail -2 gpurun_out/r3a_pytest1.log
