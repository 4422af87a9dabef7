"""Renders profiles/parity_r02.md from the measurements the GPU tests record (gpurun_out/parity_measured.jsonl, committed copy
profiles/r02_parity_measured.jsonl):   python tools/parity_report.py > profiles/parity_r02.md"""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "r02_parity_measured.jsonl")
rows = [json.loads(l) for l in open(path) if l.strip()]
d = [r for r in rows if r.get("test") == "vit_l14_bert_base_full_depth"][-1]
g = d["grad_rel_l2"]
print("# Parity at BASELINE depth, measured on B200 (round 2)\n")
print(f"Source: `{os.path.relpath(path, root)}` (written by `tests/test_depth_gpu.py`; {len(rows)} recorded run(s), the last one shown). "
      "Model: the real `CONFIGS[\"ViT-L-14\"]` towers — 24 ViT-L/14 blocks + 12 BERT-base layers (BASELINE.json configs[1]) — random reference "
      f"initialisation, B = {d['B']}, text length {d['L_text']} with padding. Oracle: `oracle/restated.py` in fp32 on the CPU on the same bf16-rounded "
      f"weights and inputs ({d['oracle_seconds']} s). Calibrator: the same oracle arithmetic executed in bf16 by torch eager on the GPU = what the "
      "reference's modules produce after `.cuda().bfloat16()`. All figures are rel-L2 = ‖got − oracle‖ / ‖oracle‖.\n")
print("| quantity | b200mm (bf16 storage, fp32 accumulation) | eager bf16 reference arithmetic |")
print("|---|---|---|")
for k, (a, b) in d["features"].items():
    print(f"| {k} (L2-normalised [B, 768]) | {a:.3e} | {b:.3e} |")
print(f"| contrastive loss, absolute error (loss = {d['loss']['oracle']:.5f}) | {d['loss_abs_err'][0]:.2e} | {d['loss_abs_err'][1]:.2e} |")
print(f"| parameter gradients, median over {g['n_params']} tensors | {g['median'][0]:.3f} | {g['median'][1]:.3f} |")
print(f"| parameter gradients, 90th percentile | {g['p90'][0]:.3f} | {g['p90'][1]:.3f} |")
print(f"| parameter gradients, maximum | {g['max'][0]:.3f} | {g['max'][1]:.3f} |")
w = g["worst_vs_eager"]
print(f"| worst ratio ours / eager: `{w['name']}` | {w['ours']:.3f} | {w['eager']:.3f} |")
print("\nError by depth (largest weight-gradient error of a block):\n")
print("| block | b200mm | eager bf16 |")
print("|---|---|---|")
for k, (a, b) in d["by_depth_max_weight_grad"].items():
    print(f"| {k} | {a:.3f} | {b:.3f} |")
print("""
Reading: the forward error after 24 + 12 layers is 1.1–1.2e-2, below the eager-bf16 reference arithmetic (1.4–1.5e-2) and flat in depth —
one bf16 rounding is 2⁻⁹ = 2e-3 relative, so the north star's "1e-3 rel" (an fp16 / single-kernel figure) is not attainable for a 36-layer
bf16 pipeline against an fp32 oracle; the bar asserted by `tests/test_depth_gpu.py` is 1.5 × these measured values and never worse than
2 × the eager-bf16 arithmetic. The gradient errors (≈ 10 %) are a property of the test point, not of the kernels: at random
initialisation and B = 4 the contrastive gradient is a difference of nearly equal terms, and the reference's own arithmetic in bf16 is 11.6 %
from fp32; b200mm is closer to the oracle than the calibrator at the median, the 90th percentile and in every block.

Other measured parity points of round 2 (logs under `profiles/`):

| check | result | log |
|---|---|---|
| full GPU suite (C-ABI kernels vs oracle / reference golden vectors) | 122 passed, 1 skipped (needs 2 GPUs) | `r02a_pytest_gpu.log` |
| dropout p > 0: masks bit-equal to the oracle's hash; fused GEMM-epilogue / attention / embedding sites vs the masked fp32 reference; training-mode BertModel vs the unmodified reference run with the same preset masks (`tests/golden/bert_dropout.pt`) | 63 passed (incl. the kernel suite) | `r02c_dropout_tests.log` |
| attention on tcgen05 for every BASELINE shape: (1,577,1,80), (2,50,2,80), (2,577,2,64), head_dim 128, L = 1000 | in the 122 | `r02a_pytest_gpu.log` |
| sharded losses vs the oracle on the gathered batch, 2 ranks: MIL-NCE (1 and 2 clips) loss 1e-7, feature gradients 2e-3; MoCo gather + queue NCE exact / 1.7e-3 | OK | `r02d_mgpu_parity_2ranks.log` |
| the same at 8 ranks (global batch 80 / 96) | OK (2.5e-3 / 2.6e-3 / 1.8e-3) | `r02e_mgpu_parity_8ranks.log` |
| sharded symmetric InfoNCE through the 2-layer model, loss vs oracle | 3.20072 vs 3.20030 (2 ranks), 4.59452 vs 4.59441 (8 ranks) | same logs |
| ... its parameter gradients vs the fp32 oracle (mid-round build) | worst rel-L2 0.12 (2 ranks, `visual.ln_post.bias`), 0.35 (8 ranks, a BERT value bias; 10 × the eager-bf16 calibrator) | same logs |

The last row was traced after the GPU budget had run out, on the CPU: `tests/test_sharded_grad_storage_cpu.py` replays what every rank does
over the emulated kernels and, with the mid-round code, reproduced the hardware figures (same worst parameters, 0.19 and 0.355). Cause: the
softmax-gradient tiles G were stored in bf16 INCLUDING the positive's "minus one-hot" entry; at random initialisation batch-summed parameter
gradients are residuals ≈ 20 × smaller than one rank's partial sum (|partial| 0.16 vs |global| 0.0077 for `ln_post.bias`), and 2⁻⁹ of that one
large entry per row does not cancel. Fix (both contrastive backends, host side only — the kernels are called with `diag_sub = 0` and the
positives' term is added to the row gradients and to the log-temperature gradient in fp32): worst parameter **0.055 / 0.049** at 2 / 8 ranks in
the same replay, level with the eager-bf16 reference arithmetic (0.05). The hardware re-run of `tests/mgpu_worker.py` with the fix is the
first item of the next GPU session; the round-end single-GPU suite exercises the same backward through `tests/test_model_gpu.py`.
""")
