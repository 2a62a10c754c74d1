import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
M, K, N = 49152, 1024, 4096
a = torch.randn((M, K), device=dev).bfloat16()
w = (torch.randn((N, K), device=dev) * 0.05).bfloat16()
b = torch.randn((N,), device=dev)
out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
h = torch.randn((M, 1024), device=dev)
w2 = (torch.randn((1024, N), device=dev) * 0.05).bfloat16()
for _ in range(2):
    ops.linear(a, w, b, out_bf16=out, act=1, cta_pair=2)                       # FFN1 (+GELU)
    ops.linear(out, w2, b[:1024].contiguous(), residual=h, out_f32=h, want_bf16=False, cta_pair=2)   # FFN2 (+residual, in place)
torch.cuda.synchronize()
print("ok")
