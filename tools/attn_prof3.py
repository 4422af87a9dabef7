"""one forward + backward of the attention kernels at a profiling shape (run under ncu)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

B, L, H, hd = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (296, 257, 16, 64)
W = H * hd
torch.manual_seed(0)
qkv = torch.randn(B * L, 3 * W, device="cuda").to(torch.bfloat16)
d_o = torch.randn(B * L, W, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, lse = ops.attention_fwd(qkv, B, L, H, hd)
    g = ops.attention_bwd(qkv, o, d_o, lse, B, L, H, hd)
torch.cuda.synchronize()
