"""Sweep of the sub-LayerNorm launch shapes (threads per row-CTA, CTAs per SM) at the M2-Encoder widths; prints algorithmic GB/s
(fwd 4 B/element, bwd 6 B/element) against the measured HBM peak. Bring-up tool: results go to profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

BF = torch.bfloat16
peak = 6552.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100864
for W in (4096, 3072):
    torch.manual_seed(0)
    u = torch.randn(rows, W, device="cuda").to(BF)
    dy = torch.randn(rows, W, device="cuda").to(BF)
    w = torch.ones(W, device="cuda", dtype=BF)
    b = torch.zeros(W, device="cuda", dtype=BF)
    dw = torch.zeros(W, device="cuda")
    db = torch.zeros(W, device="cuda")
    for act, name in ((ops.ACT_GELU_ERF, "gelu"), (ops.ACT_NONE, "none")):
        _, mean, rstd = ops.act_layernorm_fwd(u, act, w, b, 1e-5)
        for cfg in ("128,8", "128,6", "128,4", "256,4", "256,3"):
            os.environ["B200MM_SUBLN_FWD"] = cfg
            ms = timeit(lambda: ops.act_layernorm_fwd(u, act, w, b, 1e-5))
            gbs = rows * W * 4 / ms / 1e6
            print(f"fwd W={W} act={name:5s} cfg={cfg:6s} {ms:7.3f} ms  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of HBM peak", flush=True)
        for cfg in ("256,2", "256,4", "512,2", "512,3"):
            os.environ["B200MM_SUBLN_BWD"] = cfg
            ms = timeit(lambda: ops.act_layernorm_bwd(dy, u, act, mean, rstd, w, dw, db))
            gbs = rows * W * 6 / ms / 1e6
            print(f"bwd W={W} act={name:5s} cfg={cfg:6s} {ms:7.3f} ms  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of HBM peak", flush=True)
    # the existing warp-per-row LayerNorm at width 1024 for comparison (same bytes per element)
x = torch.randn(rows, 1024, device="cuda").to(BF)
w1 = torch.ones(1024, device="cuda", dtype=BF)
b1 = torch.zeros(1024, device="cuda", dtype=BF)
ms = timeit(lambda: ops.layernorm_fwd(x, w1, b1, 1e-5))
print(f"ref layernorm_fwd W=1024 {ms:7.3f} ms {rows * 1024 * 4 / ms / 1e6:7.0f} GB/s")
