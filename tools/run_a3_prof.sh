#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/tmem_mufu > gpurun_out/ubench.log 2>&1
timeout 600 python tools/attn_v3_check.py time > gpurun_out/a3_time.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn3_ -s 4 -c 4 -f -o gpurun_out/a3_prof python tools/attn_prof3.py > gpurun_out/a3_prof.log 2>&1
cat gpurun_out/ubench.log gpurun_out/a3_time.log
tail -3 gpurun_out/a3_prof.log
