#!/bin/bash
# Round evidence on one B200 box: GPU test suite, smoke, both bench arms, the ncu launch list of the bench command and
# one `ncu --set full` capture of the hot kernels.  Everything lands in gpurun_out/ (copy what is to be judged to profiles/).
#   gpurun --timeout 1500 -- 'bash tools/evidence.sh r01'
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.log 2>&1; echo "ref rc=$?"
python bench.py --profile > $out/${tag}_bench.log 2>&1; echo "bench rc=$?"
tail -n 1 $out/${tag}_bench.log | cut -c 1-1800
# launch list: per-launch durations of one warm-up + one timed step of the same workload (cold-cache, serialised: use the SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > $out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# full-set capture of one launch of each hot kernel; only the raw-metrics CSV travels back (gpurun_out is capped at 64 MiB)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn_|ln_|rowsum|act_fwd' -o /tmp/${tag}_kernels \
    python tools/prof_kernels.py 1 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${tag}_kernels.ncu-rep --page raw --csv > $out/${tag}_kernels_raw.csv 2>/dev/null
if [ "${KEEP_REP:-0}" = "1" ] && [ $(stat -c %s /tmp/${tag}_kernels.ncu-rep) -lt 50000000 ]; then cp /tmp/${tag}_kernels.ncu-rep $out/; fi
du -sh $out; ls -la $out | tail -n 20
