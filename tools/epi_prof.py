import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
BF = torch.bfloat16
M, N, K = 65536, 4096, 64
a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.05).to(BF)
bias = torch.randn(N, device="cuda").to(BF)
for _ in range(3):
    y = ops.gemm(a, w, bias=bias)
torch.cuda.synchronize()
