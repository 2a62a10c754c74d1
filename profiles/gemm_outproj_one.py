"""out-proj shaped GEMM (K = N = 1024, fp32 residual in place) for an ncu capture."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
M, K, N = 75776, 1024, 1024
a = torch.randn((M, K), device=dev).bfloat16()
w = (torch.randn((N, K), device=dev) * 0.05).bfloat16()
b = torch.randn((N,), device=dev)
h = torch.randn((M, N), device=dev)
for _ in range(3):
    ops.linear(a, w, b, residual=h, out_f32=h, want_bf16=False, cta_pair=2)
torch.cuda.synchronize()
print("ok")
