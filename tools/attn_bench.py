"""Correctness (tcgen05 vs legacy mma.sync path) and micro-benchmark of the attention kernels (diagnostics; run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

CASES = [(1, 30, 1, 64, True), (2, 129, 2, 64, False), (3, 160, 2, 64, True), (1, 128, 1, 64, False), (2, 77, 3, 64, True), (3, 257, 2, 64, False), (2, 288, 2, 64, True), (2, 197, 3, 64, False), (40, 257, 16, 64, False)]


def make(B, L, H, hd, masked, seed=0):
    W = H * hd
    torch.manual_seed(seed)
    qkv = torch.randn(B * L, 3 * W, device="cuda").to(torch.bfloat16)
    d_o = torch.randn(B * L, W, device="cuda").to(torch.bfloat16)
    kb = None
    if masked:
        kb = torch.zeros(B, L, device="cuda")
        kb[0, L // 2:] = -10000.0
    return qkv, d_o, kb


def legacy(fn, var="B200MM_ATTN_LEGACY"):
    os.environ[var] = "1"
    try:
        return fn()
    finally:
        del os.environ[var]


def check_fwd(B, L, H, hd, masked):
    qkv, d_o, kb = make(B, L, H, hd, masked)
    o_ref, lse_ref = legacy(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb))
    o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
    torch.cuda.synchronize()
    eo = (o.float() - o_ref.float()).abs().max().item()
    el = (lse - lse_ref).abs().max().item()
    print(f"fwd  B={B} L={L} H={H} masked={masked}: max|o-o_legacy|={eo:.4g} max|lse diff|={el:.4g}", "OK" if eo < 2e-2 and el < 1e-3 else "MISMATCH", flush=True)


def check_bwd(B, L, H, hd, masked):
    qkv, d_o, kb = make(B, L, H, hd, masked)
    W = H * hd
    o, lse = legacy(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb))
    g_ref = legacy(lambda: ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb), "B200MM_ATTN_BWD_LEGACY")
    g = ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb)
    torch.cuda.synchronize()
    res = []
    for nm, sl in [("dq", slice(0, W)), ("dk", slice(W, 2 * W)), ("dv", slice(2 * W, 3 * W))]:
        sc = g_ref[:, sl].float().abs().max().item()
        eg = (g[:, sl].float() - g_ref[:, sl].float()).abs().max().item()
        res.append(f"{nm} {eg:.3g}/{sc:.3g} " + ("OK" if eg < 3e-2 * sc else "MISMATCH"))
    print(f"bwd  B={B} L={L} H={H} masked={masked}: " + " | ".join(res), flush=True)


def time_it(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    iters = 10
    shapes = [(128, 257, 16, 64, False), (256, 77, 12, 64, True)]
    for c in CASES + [(5, 32, 2, 64, True), (3, 320, 2, 64, False), (300, 257, 16, 64, False)]:
        check_fwd(*c)
    for B, L, H, hd, m in shapes + [(1024, 257, 16, 64, False)]:
        qkv, d_o, kb = make(B, L, H, hd, m)
        t = time_it(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb), iters)
        t1 = legacy(lambda: time_it(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb), iters), "B200MM_ATTN_FWD_V1")
        fl = 4.0 * B * H * L * L * hd
        t2 = legacy(lambda: time_it(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb), iters), "B200MM_ATTN_NOSHORT")
        print(f"time fwd B={B} L={L} H={H}: {t:.3f} ms ({fl / t / 1e9:.0f} TF/s)   [v1 kernel: {t1:.3f} ms; v2 without short-tile path: {t2:.3f} ms]", flush=True)
    if "--fwd-only" in sys.argv:
        sys.exit(0)
    for c in CASES:
        check_bwd(*c)
    for B, L, H, hd, m in shapes + [(1024, 257, 16, 64, False)]:
        qkv, d_o, kb = make(B, L, H, hd, m)
        o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
        t = time_it(lambda: ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb), iters)
        fl = 10.0 * B * H * L * L * hd
        t2 = legacy(lambda: time_it(lambda: ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb), iters), "B200MM_ATTN_NOSHORT")
        print(f"time bwd B={B} L={L} H={H}: {t:.3f} ms ({fl / t / 1e9:.0f} TF/s alg)   [without short-tile path: {t2:.3f} ms]", flush=True)
