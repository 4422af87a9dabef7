"""One launch of each hot kernel at the ViT-L/14 bench shapes — the target of the per-kernel ncu captures. PROF_ROWS = 65 792 (256 images,
default: the `--set full` capture replays every kernel ~40 times) or 263 168 (1024 images = the bench's per-launch shapes, used with the
two-metric DRAM-traffic capture that `tools/summarize_ncu.py traffic` turns into profiles/ncu_traffic_latest.json for bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

torch.manual_seed(0)
BF = torch.bfloat16
T, W = int(os.environ.get("PROF_ROWS", 65792)), 1024  # 1024 images x 257 tokens = the bench's per-launch shapes (configs[1])
NB = T // 257
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
    dx = ops.layernorm_bwd(dy, x, mean, rstd, w, dw, db, dadd=y)  # a distinct tensor as the residual-branch gradient
    # MLP up-projection with fused bias + QuickGELU + pre-activation copy (K = 1024: the epilogue-heavy GEMM)
    g, u = ops.gemm(y, wfc, bias=bfc, act=ops.ACT_QUICKGELU, aux_out=True)
    # dgrad through the activation: du = (dy @ Wproj) * act'(u)
    du = ops.gemm(dy, wproj, b_mn=True, act=ops.ACT_QUICKGELU, dact_in=u)
    # ... and the form the training step launches: the same dgrad also emitting act(u) (flavour 74, the bench's dominant launch)
    du, g2 = ops.gemm(dy, wproj, b_mn=True, act=ops.ACT_QUICKGELU, dact_in=u, aux_out=True)
    del g2
    # plain K = 4096 GEMM and a split-K weight gradient
    h = ops.gemm(g, wproj, bias=b, residual=x)
    dwp = ops.gemm(dy, g, a_mn=True, b_mn=True)
    # attention fwd / bwd, 256 x 16 heads x 257 tokens
    qkv = torch.randn(T, 3 * W, device="cuda").to(BF)
    o, lse = ops.attention_fwd(qkv, NB, 257, 16, 64)
    dqkv = ops.attention_bwd(qkv, o, dy, lse, NB, 257, 16, 64)
    ops.rowsum_periodic(du, torch.zeros(4 * W, device="cuda"))
    ops.act_fwd(u, ops.ACT_QUICKGELU)
    # M2-Encoder sub-LayerNorm over the 4W-wide FFN hidden: gelu + LN forward, LN' * gelu' backward
    w4 = torch.ones(4 * W, device="cuda", dtype=BF)
    b4 = torch.zeros(4 * W, device="cuda", dtype=BF)
    gn, m4, r4 = ops.act_layernorm_fwd(u, ops.ACT_GELU_ERF, w4, b4, 1e-5)
    ops.act_layernorm_bwd(du, u, ops.ACT_GELU_ERF, m4, r4, w4, torch.zeros(4 * W, device="cuda"), torch.zeros(4 * W, device="cuda"))
    # stage-2 pair gather: 1024 pairs x 79 tokens x 768 from a 64 x 79-row table
    tab = torch.randn(64 * 79, 768, device="cuda").to(BF)
    ids = torch.randint(0, 64 * 79, (1024 * 79,), device="cuda")
    ops.gather_rows(tab, ids)
    # contrastive LSE partials + merge at the configs[2] per-rank size
    ia = torch.nn.functional.normalize(torch.randn(8192, 768, device="cuda"), dim=-1).to(BF)
    pm = ops.contrast_lse_partials(ia[:1024].contiguous(), ia, 14.3, 0)
    ops.contrast_lse_merge(pm[:2], None, pm[2], False, torch.zeros(1, device="cuda"))
    # qkv projection (K = 1024 -> N = 3072, flavour 1) and the plain dgrad back through the out-projection (flavour 0, B MN-major)
    win = (torch.randn(3 * W, W, device="cuda") * 0.02).to(BF)
    ops.gemm(y, win, bias=torch.zeros(3 * W, device="cuda", dtype=BF))
    wout = (torch.randn(W, W, device="cuda") * 0.02).to(BF)
    ops.gemm(dy, wout, b_mn=True)
torch.cuda.synchronize()
print("done")
