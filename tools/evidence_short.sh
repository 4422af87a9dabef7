#!/bin/bash
# Round evidence on one B200 (GPU-minute budget ~8 min): GPU tests, smoke, DRAM-traffic capture of the hot kernels at the bench shapes,
# bench (quotes that traffic), ncu launch list of the bench command, one --set full capture of the hot kernels.
#   gpurun --timeout 1500 -- 'bash tools/evidence_short.sh r01k'
set -u
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
nproc > $out/${tag}_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -n 3 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
# DRAM traffic of one launch per kernel at the bench's own shapes (1024 images): three metrics only, a single replay pass
PROF_ROWS=263168 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'gemm_tcgen05|attn_|ln_|rowsum|act_' --csv --log-file $out/${tag}_traffic.csv python tools/prof_kernels.py 1 > $out/${tag}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python tools/summarize_ncu.py traffic $out/${tag}_traffic.csv 263168 > $out/${tag}_ncu_traffic.json 2>/dev/null && cp $out/${tag}_ncu_traffic.json profiles/ncu_traffic_latest.json
timeout 600 python bench.py --profile > $out/${tag}_bench.log 2>&1; echo "bench rc=$?"
tail -n 1 $out/${tag}_bench.log | cut -c 1-2800
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > $out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn_|ln_|rowsum|act_fwd|gather_rows|lse_merge' -o /tmp/${tag}_kernels \
    python tools/prof_kernels.py 1 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${tag}_kernels.ncu-rep --page raw --csv > $out/${tag}_kernels_raw.csv 2>/dev/null
du -sh $out
