#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_vtp_gpu.py tests/test_m2_gpu.py tests/test_cross_gpu.py -q -m gpu > $out/r01f_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -n 30 $out/r01f_pytest_new.log | cut -c1-260
timeout 300 python bench.py --model base_vtp-ViT-B-16 --batch 32 --keep-act 0 --steps 3 --warmup 2 --no-cpu-baseline --profile > $out/r01f_bench_vtp.log 2>&1; echo "vtp bench rc=$?"
tail -n 2 $out/r01f_bench_vtp.log | cut -c1-1800
cp $out/op_profile_b32.json $out/r01f_op_profile_vtp_b32.json 2>/dev/null
timeout 300 python bench.py --model M2-Encoder-1B --batch 512 --seq-len 52 --keep-act 24 --steps 3 --warmup 2 --no-cpu-baseline --profile > $out/r01f_bench_m2_keep.log 2>&1; echo "m2 bench rc=$?"
tail -n 1 $out/r01f_bench_m2_keep.log | cut -c1-900
cp $out/op_profile_b512.json $out/r01f_op_profile_m2_b512.json 2>/dev/null
