import os, sys
sys.path.insert(0, "/root/repo")
import torch
from aptai_b200 import ops
dev = torch.device("cuda:0")
M, K, N = 47880, 4096, 1024
a = torch.randn((M, K), device=dev).bfloat16(); w = (torch.randn((N, K), device=dev) * 0.02).bfloat16()
h = torch.randn((M, N), device=dev); b = torch.zeros((N,), device=dev); g = torch.ones((N,), device=dev)
for _ in range(3):
    ops.linear(a, w, b, residual=h, out_f32=h, want_bf16=False, row_ln=(g, b, 1e-5))
torch.cuda.synchronize()
