#!/bin/bash
# bring-up probe for the tcgen05 GEMM (run under gpurun); each config in its own process with a timeout so a hang
# in one layout does not hide the others.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/probe_gemm.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
run() { echo "## $*" >> $LOG; timeout 60 ./tools/gemm_probe "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run 0 0 128 256 64 1
run 0 0 128 256 256 1
run 0 0 256 512 512 1
run 0 0 1000 776 520 1
run 0 1 128 256 64 1
run 0 1 1000 776 520 1
run 1 0 128 256 64 1
run 1 0 1000 776 520 1
run 1 1 128 256 64 1
run 1 1 1000 776 520 1
run 1 1 512 512 4096 4
run 0 0 300 264 4096 3
echo "## swapped MN lbo/sbo" >> $LOG
B200MM_DBG_DESC="1024,8192,16,1024" run 1 1 128 256 64 1
B200MM_DBG_DESC="1024,8192,16,1024" run 0 1 128 256 64 1
echo "## perf" >> $LOG
run 0 0 16384 4096 1024 1 20
run 0 0 32768 1024 4096 1 20
run 0 1 16384 1024 4096 1 20
run 1 1 4096 1024 32768 4 20
run 1 1 1024 1024 32768 16 20
cat $LOG
