"""Correctness of the tcgen05 attention kernels against an fp32 torch restatement on the same bf16 inputs, and their device time
(diagnostics; run under gpurun). Usage: attn_v3_check.py fwd|bwd|time [case indices...]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

# (B, L, H, hd, masked)
CASES = [
    (1, 16, 1, 64, False), (1, 128, 1, 64, False), (2, 77, 3, 64, True), (3, 5, 2, 32, False), (2, 12, 2, 16, True),
    (1, 129, 1, 64, False), (2, 257, 2, 64, False), (3, 197, 2, 64, False), (2, 288, 2, 64, True), (2, 300, 1, 64, False),
    (2, 50, 2, 80, True), (1, 577, 1, 80, False), (2, 577, 2, 64, False), (2, 86, 12, 64, True), (2, 200, 2, 128, False),
    (1, 1000, 2, 64, True), (150, 257, 2, 64, False), (40, 577, 16, 80, False), (3, 160, 4, 96, True),
]


def make(B, L, H, hd, masked, seed=0):
    W = H * hd
    g = torch.Generator(device="cuda").manual_seed(seed + B * L + hd)
    qkv = torch.randn(B * L, 3 * W, device="cuda", generator=g).to(torch.bfloat16)
    d_o = torch.randn(B * L, W, device="cuda", generator=g).to(torch.bfloat16)
    kb = None
    if masked:
        kb = torch.zeros(B, L, device="cuda")
        kb[0, L // 2:] = -10000.0
        kb[-1, L - 3:] = -10000.0
    return qkv, d_o, kb


def reference(qkv, d_o, kb, B, L, H, hd):
    W = H * hd
    q = qkv.float().view(B, L, 3, H, hd).requires_grad_()
    s = torch.einsum("blhd,bmhd->bhlm", q[:, :, 0], q[:, :, 1]) / math.sqrt(hd)
    if kb is not None:
        s = s + kb[:, None, None, :]
    lse = torch.logsumexp(s, -1)
    o = torch.einsum("bhlm,bmhd->blhd", torch.softmax(s, -1), q[:, :, 2]).reshape(B * L, W)
    o.backward(d_o.float())
    return o.detach(), lse.detach(), q.grad.reshape(B * L, 3 * W)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def check(idx, do_bwd):
    B, L, H, hd, masked = CASES[idx]
    qkv, d_o, kb = make(B, L, H, hd, masked)
    o_ref, lse_ref, g_ref = reference(qkv, d_o, kb, B, L, H, hd)
    o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
    torch.cuda.synchronize()
    eo, el = rel(o, o_ref), (lse - lse_ref).abs().max().item()
    ok = eo < 1.5e-2 and el < 2e-3 and bool(torch.isfinite(o.float()).all())
    msg = f"case {idx} B={B} L={L} H={H} hd={hd} masked={masked}: fwd o {eo:.3g} lse {el:.3g} {'OK' if ok else 'MISMATCH'}"
    if do_bwd:
        W = H * hd
        g = ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb)
        torch.cuda.synchronize()
        for nm, sl in [("dq", slice(0, W)), ("dk", slice(W, 2 * W)), ("dv", slice(2 * W, 3 * W))]:
            e = rel(g[:, sl], g_ref[:, sl])
            good = e < 2e-2 and bool(torch.isfinite(g[:, sl].float()).all())
            ok = ok and good
            msg += f" | {nm} {e:.3g} {'OK' if good else 'MISMATCH'}"
    print(msg, flush=True)
    return ok


def time_it(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timing():
    shapes = [(1024, 257, 16, 64, False), (256, 77, 12, 64, True), (512, 197, 12, 64, False), (128, 577, 16, 80, False), (256, 257, 16, 80, False),
              (64, 249, 32, 128, False)]
    for B, L, H, hd, m in shapes:
        qkv, d_o, kb = make(B, L, H, hd, m)
        for tag in ("v3",):
            try:
                o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
                tf = time_it(lambda: ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb))
                tb = time_it(lambda: ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb))
                ff, fb = 4.0 * B * H * L * L * hd, 10.0 * B * H * L * L * hd
                print(f"time {tag} B={B} L={L} H={H} hd={hd}: fwd {tf:.3f} ms ({ff / tf / 1e9:.0f} TF/s)  bwd {tb:.3f} ms ({fb / tb / 1e9:.0f} TF/s alg)", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"time {tag} B={B} L={L} H={H} hd={hd}: FAILED {e}", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "time":
        timing()
        sys.exit(0)
    idxs = [int(a) for a in sys.argv[2:]] or list(range(len(CASES)))
    bad = 0
    for i in idxs:
        try:
            bad += 0 if check(i, mode == "bwd") else 1
        except Exception as e:  # noqa: BLE001
            print(f"case {i} {CASES[i]}: EXCEPTION {e}", flush=True)
            bad += 1
            break  # a trapped kernel poisons the context
    print(f"{mode}: {len(idxs) - bad}/{len(idxs)} ok", flush=True)
