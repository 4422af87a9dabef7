#!/bin/bash
# Round-2 evidence on one B200: GPU tests, smoke, reference arm, DRAM-traffic capture at the bench shapes, bench (quotes that traffic, runs the
# eager-GPU baseline), ncu launch list of the bench command, one --set full capture of the hot kernels, attention timings.
#   gpurun --timeout 1500 -- 'bash tools/evidence_r02.sh r02a'
set -u
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nproc > $out/${tag}_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -n 3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python tools/attn_v3_check.py time > $out/${tag}_attn_time.log 2>&1; echo "attn time rc=$?"; tail -n 8 $out/${tag}_attn_time.log
PROF_ROWS=263168 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'gemm_tcgen05|attn|ln_|rowsum|act_' --csv --log-file $out/${tag}_traffic.csv python tools/prof_kernels.py 1 > $out/${tag}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python tools/summarize_ncu.py traffic $out/${tag}_traffic.csv 263168 > $out/${tag}_ncu_traffic.json 2>/dev/null && cp $out/${tag}_ncu_traffic.json profiles/ncu_traffic_latest.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.log 2>&1; echo "ref rc=$?"; tail -n 1 $out/${tag}_bench_reference.log | cut -c 1-600
timeout 600 python bench.py --profile > $out/${tag}_bench.log 2>&1; echo "bench rc=$?"
tail -n 1 $out/${tag}_bench.log | cut -c 1-3200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --skip-e2e > $out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn|ln_|rowsum|act_fwd|gather_rows|lse_merge|dropout' -o /tmp/${tag}_kernels \
    python tools/prof_kernels.py 1 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${tag}_kernels.ncu-rep --page raw --csv > $out/${tag}_kernels_raw.csv 2>/dev/null
du -sh $out
