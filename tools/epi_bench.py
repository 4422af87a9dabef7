"""Epilogue cost probe: the GEMM at K = 64 (mainloop negligible) vs K = 1024 for each fused-epilogue flavour (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops

BF = torch.bfloat16
M, N = 65536, 4096
torch.manual_seed(0)
u = torch.randn(M, N, device="cuda").to(BF)
res = torch.randn(M, N, device="cuda").to(BF)
bias = torch.randn(N, device="cuda").to(BF)

def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

for K in (64, 1024):
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(BF)
    tiles = (M // 256) * (N // 256)
    flav = {
        "plain": lambda: ops.gemm(a, w),
        "bias": lambda: ops.gemm(a, w, bias=bias),
        "bias+quickgelu": lambda: ops.gemm(a, w, bias=bias, act=ops.ACT_QUICKGELU),
        "bias+quickgelu+aux": lambda: ops.gemm(a, w, bias=bias, act=ops.ACT_QUICKGELU, aux_out=True),
        "residual": lambda: ops.gemm(a, w, residual=res),
        "dact quickgelu": lambda: ops.gemm(a, w, act=ops.ACT_QUICKGELU, dact_in=u),
        "f32 out": lambda: ops.gemm(a, w, out_f32=True),
    }
    for name, fn in flav.items():
        ms = t(fn)
        per_tile_us = ms * 1e3 / (tiles / 74)  # macro-tiles per CTA pair
        print(f"K={K:5d} {name:22s} {ms:7.3f} ms  {2.0*M*N*K/ms/1e9:7.0f} TF/s  {per_tile_us:6.2f} us per 256x256 macro-tile per pair", flush=True)
