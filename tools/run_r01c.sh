#!/bin/bash
# new-feature GPU check: M2-Encoder + stage-2 cross-modal tests, then the whole GPU suite, then a short M2-Encoder-1B bench
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_m2_gpu.py tests/test_cross_gpu.py -q -m gpu > $out/r01c_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -n 40 $out/r01c_pytest_new.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -x -q > $out/r01c_pytest_gpu.log 2>&1; echo "all gpu tests rc=$?"
tail -n 3 $out/r01c_pytest_gpu.log
timeout 300 python bench.py --model M2-Encoder-1B --batch 512 --seq-len 52 --keep-act 0 --steps 3 --warmup 2 --no-cpu-baseline --profile > $out/r01c_bench_m2.log 2>&1; echo "m2 bench rc=$?"
tail -n 2 $out/r01c_bench_m2.log | cut -c1-1500
