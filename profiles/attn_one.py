import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
B, T, heads, H = 120, 399, 16, 1024
qkv = (torch.randn((B * T, 3 * H), device=dev) * 0.5).bfloat16()
kl = torch.full((B,), T, dtype=torch.int32, device=dev)
out = torch.empty((B * T, H), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attention(qkv, kl, B, T, heads, out=out)
torch.cuda.synchronize()
print("ok")
