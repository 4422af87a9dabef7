"""One launch of each hot kernel at (a slice of) the ViT-L/14 bench shapes — the target of the per-kernel ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

torch.manual_seed(0)
BF = torch.bfloat16
T, W = 65792, 1024  # 256 images x 257 tokens
x = torch.randn(T, W, device="cuda").to(BF)
dy = torch.randn(T, W, device="cuda").to(BF)
w = torch.ones(W, device="cuda", dtype=BF)
b = torch.zeros(W, device="cuda", dtype=BF)
wfc = (torch.randn(4 * W, W, device="cuda") * 0.02).to(BF)
bfc = torch.zeros(4 * W, device="cuda", dtype=BF)
wproj = (torch.randn(W, 4 * W, device="cuda") * 0.02).to(BF)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    # LayerNorm fwd / bwd
    y, _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    dw = torch.zeros(W, device="cuda")
    db = torch.zeros(W, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, w, dw, db, dadd=dy)
    # MLP up-projection with fused bias + QuickGELU + pre-activation copy (K = 1024: the epilogue-heavy GEMM)
    g, u = ops.gemm(y, wfc, bias=bfc, act=ops.ACT_QUICKGELU, aux_out=True)
    # dgrad through the activation: du = (dy @ Wproj) * act'(u)
    du = ops.gemm(dy, wproj, b_mn=True, act=ops.ACT_QUICKGELU, dact_in=u)
    # plain K = 4096 GEMM and a split-K weight gradient
    h = ops.gemm(g, wproj, bias=b, residual=x)
    dwp = ops.gemm(dy, g, a_mn=True, b_mn=True)
    # attention fwd / bwd, 256 x 16 heads x 257 tokens
    qkv = torch.randn(T, 3 * W, device="cuda").to(BF)
    o, lse = ops.attention_fwd(qkv, 256, 257, 16, 64)
    dqkv = ops.attention_bwd(qkv, o, dy, lse, 256, 257, 16, 64)
    ops.rowsum_periodic(du, torch.zeros(4 * W, device="cuda"))
    ops.act_fwd(u, ops.ACT_QUICKGELU)
torch.cuda.synchronize()
print("done")
