"""one forward + backward attention launch at the ViT-L/14 shape (for ncu captures)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
B, L, H, hd = 128, 257, 16, 64
qkv = torch.randn(B * L, 3 * H * hd, device="cuda").to(torch.bfloat16)
d_o = torch.randn(B * L, H * hd, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, lse = ops.attention_fwd(qkv, B, L, H, hd)
    ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd)
torch.cuda.synchronize()
