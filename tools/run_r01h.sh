#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python tools/contrast_bench.py > $out/r01h_contrast_bench.log 2>&1; echo "contrast bench rc=$?"
cat $out/r01h_contrast_bench.log | cut -c1-1100
ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|lse_merge' -c 40 -o /tmp/r01h_contrast python tools/contrast_bench.py 768 > $out/r01h_ncu_contrast.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r01h_contrast.ncu-rep --page raw --csv > $out/r01h_contrast_raw.csv 2>/dev/null
timeout 300 python tools/subln_bench.py > $out/r01h_subln_sweep.log 2>&1
grep -E "W=4096.*(128,6|128,8|512,2|256,2)" $out/r01h_subln_sweep.log
