"""Micro-benchmark of the attention kernels at the ViT-L/14 and BERT-base shapes (diagnostics; run under gpurun / ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200mm
from b200mm import ops

def run(B, L, H, hd, masked, iters=10):
    W = H * hd
    qkv = torch.randn(B * L, 3 * W, device="cuda").to(torch.bfloat16)
    d_o = torch.randn(B * L, W, device="cuda").to(torch.bfloat16)
    kb = None
    if masked:
        kb = torch.zeros(B, L, device="cuda")
        kb[:, L // 2:] = -10000.0
    o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
    ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(iters):
        o, lse = ops.attention_fwd(qkv, B, L, H, hd, key_bias=kb)
    e[1].record()
    for _ in range(iters):
        ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=kb)
    e[2].record()
    torch.cuda.synchronize()
    f = e[0].elapsed_time(e[1]) / iters
    b = e[1].elapsed_time(e[2]) / iters
    flops = 4.0 * B * H * L * L * hd
    print(f"B={B} L={L} H={H} hd={hd}: fwd {f:.3f} ms ({flops / f / 1e9:.0f} TF/s)  bwd {b:.3f} ms ({2.5 * flops / b / 1e9:.0f} TF/s alg)")

if __name__ == "__main__":
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    run(128, 257, 16, 64, False, it)
    run(256, 77, 12, 64, True, it)
