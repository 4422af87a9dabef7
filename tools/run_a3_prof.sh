#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn -s 3 -c 3 -f -o gpurun_out/a3_prof python tools/attn_prof3.py "$@" > gpurun_out/a3_prof.log 2>&1
tail -3 gpurun_out/a3_prof.log
