#!/bin/bash
# Round-2 multi-GPU evidence on one 8 x B200 node:   gpurun --gpus 8 --timeout 1500 -- 'bash tools/run_r02_8gpu.sh r02e'
#   1. every sharded path vs the oracle on the gathered batch at 8 ranks (tests/mgpu_worker.py)
#   (the r02e logs under profiles/ were taken when `--grad-sync ddp` was still the default; the flags below reproduce them)
#   2. configs[2] (ViT-L/14 + BERT-base, global batch 8192): DDP bucketed overlap vs flat all-reduce after backward, with per-rank step time / GEMM rate
#   3. configs[4] as written: ViT-H/14 at 336^2 + BERT-large, global batch 4096 (512 per GPU)
#   4. configs[3] as written: base_vtp ViT-B/16 x 8 frames + BERT-base, global batch 256 (32 per GPU)
set -u
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
N=${NGPU:-8}
run() { timeout ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29601 tests/mgpu_worker.py > $out/${tag}_mgpu_parity_${N}ranks.log 2>&1; echo "mgpu rc=$?"; grep MGPU $out/${tag}_mgpu_parity_${N}ranks.log
run 29602 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --grad-sync ddp > $out/${tag}_bench_${N}gpu_ddp.log 2>&1; echo "ddp rc=$?"
tail -n 1 $out/${tag}_bench_${N}gpu_ddp.log | cut -c 1-260; tail -n 1 $out/${tag}_bench_${N}gpu_ddp.log | grep -o '"per_rank.*' | cut -c 1-700
run 29603 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --skip-e2e --grad-sync flat > $out/${tag}_bench_${N}gpu_flat.log 2>&1; echo "flat rc=$?"
tail -n 1 $out/${tag}_bench_${N}gpu_flat.log | cut -c 1-260; tail -n 1 $out/${tag}_bench_${N}gpu_flat.log | grep -o '"per_rank.*' | cut -c 1-700
NCCL_MAX_CTAS=8 run 29604 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --skip-e2e --grad-sync ddp > $out/${tag}_bench_${N}gpu_ddp_maxctas8.log 2>&1; echo "ddp maxctas rc=$?"
tail -n 1 $out/${tag}_bench_${N}gpu_ddp_maxctas8.log | cut -c 1-260; tail -n 1 $out/${tag}_bench_${N}gpu_ddp_maxctas8.log | grep -o '"per_rank.*' | cut -c 1-700
TMO=600 run 29605 bench.py --gpus $N --model ViT-H-14 --image-res 336 --batch 512 --ckpt-every 2 --keep-ln 0 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_${N}gpu_vith336.log 2>&1; echo "vith rc=$?"
if ! tail -n 1 $out/${tag}_bench_${N}gpu_vith336.log | grep -q '"metric"'; then  # out of memory with half of the blocks kept: re-run every block in backward
  tail -n 3 $out/${tag}_bench_${N}gpu_vith336.log | cut -c 1-300
  TMO=600 run 29615 bench.py --gpus $N --model ViT-H-14 --image-res 336 --batch 512 --ckpt-every 1 --keep-ln 0 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_${N}gpu_vith336.log 2>&1; echo "vith (ckpt 1) rc=$?"
fi
tail -n 1 $out/${tag}_bench_${N}gpu_vith336.log | cut -c 1-1500
run 29606 bench.py --gpus $N --model base_vtp-ViT-B-16 --batch 32 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/${tag}_bench_${N}gpu_vtp.log 2>&1; echo "vtp rc=$?"
tail -n 1 $out/${tag}_bench_${N}gpu_vtp.log | cut -c 1-1500
