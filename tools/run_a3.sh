#!/bin/bash
# attention v3 on a B200 (under gpurun): correctness per mode, then timing; each step has its own (short) timeout
mkdir -p gpurun_out
{
timeout 120 python tools/attn_v3_check.py fwd 2>&1 | grep -v Warning | tail -3
timeout 120 python tools/attn_v3_check.py bwd 2>&1 | grep -v Warning | grep -v "OK | dq.*OK | dk.*OK | dv.*OK" | tail -8
timeout 120 python tools/attn_v3_check.py time 2>&1 | tail -20
} > gpurun_out/a3_check.log 2>&1
tail -40 gpurun_out/a3_check.log
if [ "$1" == "prof" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn3_ -s 4 -c 4 -f -o gpurun_out/a3_prof python tools/attn_prof3.py > gpurun_out/a3_prof.log 2>&1
tail -2 gpurun_out/a3_prof.log
fi
