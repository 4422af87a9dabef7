#!/bin/bash
# same-box A/B of the epilogue wait policy (spin vs suspending try_wait): step time, GEMM TF/s, SM clock under load
for ns in 0 1000 0 200 5000; do
  B200MM_GEMM_WAIT_NS=$ns python bench.py --steps 4 --warmup 3 --no-cpu-baseline --skip-e2e 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('wait_ns=$ns', 'ms/step', d['ms_per_step'], 'pairs/s', d['value'], 'gemm TF/s', d['roofline']['achieved'], 'sm_mhz', d['clocks']['sm_mhz'])"
done
