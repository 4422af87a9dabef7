"""Per-kernel roofline table from the committed ncu capture (profiles/ncu_traffic_latest.json: one launch per hot kernel at the bench
shapes, T = 263 168 token rows = 1024 images x 257 tokens, width 1024): measured time and DRAM bytes next to the kernel's ALGORITHMIC
work (tools/prof_kernels.py shapes), achieved TFLOP/s or GB/s, and the fraction of the measured B200 peak that bounds it.

  python tools/roofline_table.py > profiles/r02a_roofline_table.md
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_latest.json")))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK_TF, PEAK_TF_BURST, PEAK_HBM = pk.get("bf16_tflops_sustained", 1407.1), pk.get("bf16_tflops", 1636.2), pk.get("hbm_gbs", 6552.0)
T, W, NB, L, H, HD = 263168, 1024, 1024, 257, 16, 64
bf = 2
# kernel -> (what, bound, algorithmic FLOPs, algorithmic bytes)
work = {
    "gemm_tcgen05_kernel<0, 0, 0, 2, 67>": ("c_fc fwd: [T,1024]x[4096,1024]^T + bias + QuickGELU + pre-activation copy", "tensor", 2.0 * T * 4096 * 1024, bf * (T * 1024 + 4096 * 1024 + 2 * T * 4096)),
    "gemm_tcgen05_kernel<0, 1, 0, 2, 72>": ("dgrad through the activation: dy [T,1024] x W [1024,4096] x act'(u)", "tensor", 2.0 * T * 4096 * 1024, bf * (T * 1024 + 4096 * 1024 + 2 * T * 4096)),
    "gemm_tcgen05_kernel<0, 0, 0, 2, 5>": ("c_proj fwd: [T,4096]x[1024,4096]^T + bias + residual", "tensor", 2.0 * T * 1024 * 4096, bf * (T * 4096 + 4096 * 1024 + 2 * T * 1024)),
    "gemm_tcgen05_kernel<1, 1, 0, 2, ": ("wgrad split-K: dy^T [1024,T] x g [T,4096] (both MN-major)", "tensor", 2.0 * T * 1024 * 4096, bf * (T * 1024 + T * 4096) + 2 * 1024 * 4096),
    "attn_fwd_tc2_kernel": ("attention fwd, 1024 x 16 heads x 257 tokens, head_dim 64", "tensor/MUFU", 4.0 * NB * H * L * L * HD, bf * 4 * T * W),
    "attn_bwd_tc_kernel": ("attention bwd (dQ, dK, dV in one kernel)", "tensor/MUFU", 10.0 * NB * H * L * L * HD, bf * 8 * T * W),
    "attn_dsum_kernel": ("D = rowsum(dO * O)", "hbm", 0, bf * 2 * T * W),
    "ln_fwd_plain_kernel<4>": ("LayerNorm fwd [T,1024]", "hbm", 0, bf * 2 * T * W),
    # (round 2: tools/prof_kernels.py passes a distinct tensor as the residual-branch gradient: dy, x, dadd read + dx written; the round-1
    # capture aliased dadd = dy and under-counted the kernel's bytes, which is where its "0.74 of HBM" came from)
    "ln_bwd_plain_kernel<4, 1>": ("LayerNorm bwd + residual-gradient add [T,1024] (dy, x, dadd read; dx written)", "hbm", 0, bf * 4 * T * W),
    "rowsum_periodic_kernel": ("bias gradient: column sums of [T,4096]", "hbm", 0, bf * T * 4096),
    "act_fwd_kernel": ("QuickGELU recompute [T,4096]", "hbm", 0, bf * 2 * T * 4096),
    "act_ln_fwd_kernel<4, 128, 2, 1>": ("M2 sub-LN fwd: LN(gelu(u)) [T,4096]", "hbm (issue/MUFU-limited)", 0, bf * 2 * T * 4096),
    "act_ln_bwd_kernel<1, 512, 2, 2>": ("M2 sub-LN bwd: LN' x gelu' [T,4096]", "hbm (issue/MUFU-limited)", 0, bf * 3 * T * 4096),
}
print("# Per-kernel roofline at the bench shapes (ViT-L/14, B = 1024: T = 263 168 rows) — ncu, one launch each\n")
print(f"Source: `{d.get('source')}` ({d.get('how')}). Peaks: measured sustained bf16 {PEAK_TF:.0f} TFLOP/s (burst {PEAK_TF_BURST:.0f}), "
      f"measured HBM copy {PEAK_HBM:.0f} GB/s (MEASURED_PEAKS.json). ncu launches are cold-cache and serialised: in the live bench the same\n"
      "GEMMs run 5-10 % faster (bench.py `roofline`), so these fractions are lower bounds.\n")
print("| kernel | what | time µs | algorithmic | measured DRAM bytes (÷ algorithmic) | achieved | bound | fraction of measured peak |")
print("|---|---|---|---|---|---|---|---|")
for name, k in d["kernels"].items():
    if name not in work:
        continue
    what, bound, fl, by = work[name]
    t = k["time_us"] * 1e-6
    ratio = k["dram_bytes"] / by
    if fl:
        ach = fl / t / 1e12
        frac = ach / PEAK_TF
        print(f"| `{name}` | {what} | {k['time_us']:.0f} | {fl / 1e12:.2f} TFLOP, {by / 1e9:.2f} GB | {k['dram_bytes'] / 1e9:.2f} GB ({ratio:.2f}x) | {ach:.0f} TFLOP/s | {bound} | {frac:.2f} of sustained bf16 |")
    else:
        ach = by / t / 1e9
        print(f"| `{name}` | {what} | {k['time_us']:.0f} | {by / 1e9:.2f} GB | {k['dram_bytes'] / 1e9:.2f} GB ({ratio:.2f}x) | {ach:.0f} GB/s | {bound} | {ach / PEAK_HBM:.2f} of HBM |")
