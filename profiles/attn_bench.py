"""Attention micro-benchmark: tcgen05 kernel vs the legacy mma.sync kernel (CUDA events, L2 flushed)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=8):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for B, T in [(120, 399), (48, 999), (80, 600), (189, 250)]:
    heads, H = 16, 1024
    qkv = (torch.randn((B * T, 3 * H), device=dev) * 0.5).bfloat16()
    kl = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = torch.empty((B * T, H), dtype=torch.bfloat16, device=dev)
    fl = 4.0 * B * heads * T * T * 64
    a = timeit(lambda: ops.attention(qkv, kl, B, T, heads, out=out, impl=1))
    c = timeit(lambda: ops.attention(qkv, kl, B, T, heads, out=out, impl=2))
    line = (f"B={B:4d} T={T:4d}: tcgen05 v1 {a:7.3f} ms {fl / a / 1e9:7.1f} TFLOP/s | v2 (q-tile pairs) {c:7.3f} ms "
            f"{fl / c / 1e9:7.1f} TFLOP/s | v3 (P in TMEM), polynomial pairs per 8:")
    for poly in (0, 2, 3, 4):
        ops.ATTENTION_POLY8 = poly
        d = timeit(lambda: ops.attention(qkv, kl, B, T, heads, out=out, impl=3))
        line += f" [{poly & 255}] {d:7.3f} ms {fl / d / 1e9:7.1f}"
    ops.ATTENTION_POLY8 = 3
    print(line, flush=True)
