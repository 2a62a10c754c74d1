"""Ablation timing of the tcgen05 attention kernel (APTAI_ATTN_DBG bitmask: 1 no exp, 2 no P store, 4 no S load)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
B, T, heads, H = 120, 399, 16, 1024
qkv = (torch.randn((B * T, 3 * H), device=dev) * 0.5).bfloat16()
kl = torch.full((B,), T, dtype=torch.int32, device=dev)
out = torch.empty((B * T, H), dtype=torch.bfloat16, device=dev)
for _ in range(3): ops.attention(qkv, kl, B, T, heads, out=out)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention(qkv, kl, B, T, heads, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("dbg", os.environ.get("APTAI_ATTN_DBG", "0"), "median ms", sorted(ts)[5])
