#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_m2_gpu.py tests/test_vtp_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu > $out/r01g_pytest.log 2>&1; echo "tests rc=$?"
tail -n 12 $out/r01g_pytest.log | cut -c1-260
timeout 200 python tools/subln_bench.py > $out/r01g_subln_sweep.log 2>&1; echo "sweep rc=$?"
grep -E "W=4096" $out/r01g_subln_sweep.log
timeout 300 python bench.py --model M2-Encoder-1B --batch 512 --seq-len 52 --keep-act 24 --steps 5 --warmup 3 --no-cpu-baseline --profile > $out/r01g_bench_m2.log 2>&1; echo "m2 bench rc=$?"
tail -n 1 $out/r01g_bench_m2.log | cut -c1-700
cp $out/op_profile_b512.json $out/r01g_op_profile_m2_b512.json 2>/dev/null
timeout 300 python bench.py --model base_vtp-ViT-B-16 --batch 32 --keep-act 0 --steps 20 --warmup 3 --no-cpu-baseline > $out/r01g_bench_vtp.log 2>&1; echo "vtp bench rc=$?"
tail -n 1 $out/r01g_bench_vtp.log | cut -c1-400
timeout 400 python bench.py --model ViT-H-14 --batch 512 --keep-act 0 --steps 3 --warmup 2 --no-cpu-baseline --profile > $out/r01g_bench_vith.log 2>&1; echo "vit-h bench rc=$?"
tail -n 1 $out/r01g_bench_vith.log | cut -c1-900
cp $out/op_profile_b512.json $out/r01g_op_profile_vith_b512.json 2>/dev/null
