#!/bin/bash
# same-box A/B of the ViT memory policies at the bench shape + the new GPU tests
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -n 3
: > $out/r01l_ab_keep_ln.log
for cfg in "8 0" "0 0" "0 16" "0 20" "0 16"; do
  set -- $cfg
  timeout 200 python bench.py --keep-act $1 --keep-ln $2 --steps 5 --warmup 3 --no-cpu-baseline --skip-e2e > $out/tmp_bench.log 2>&1
  python - "$1" "$2" <<'PY' >> $out/r01l_ab_keep_ln.log
import json, sys
try:
    d = json.loads(open("gpurun_out/tmp_bench.log").read().strip().splitlines()[-1])
    print("keep_act", sys.argv[1], "keep_ln", sys.argv[2], "ms/step", d["ms_per_step"], "pairs/s", d["value"], "sm_mhz", d["clocks"]["sm_mhz"], "gemm TF/s", d["roofline"]["achieved"], "peak GiB", d["config"]["peak_mem_gib"])
except Exception as e:
    print("keep_act", sys.argv[1], "keep_ln", sys.argv[2], "FAILED", repr(e), open("gpurun_out/tmp_bench.log").read()[-400:])
PY
done
cat $out/r01l_ab_keep_ln.log
