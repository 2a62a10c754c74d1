"""Grouped positional conv at B=120 x T=399 (24x1024 model): slab kernel vs the generic implicit-GEMM path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aptai_b200 import ops
dev = torch.device("cuda:0")
B, T, H, groups, taps = 120, 399, 1024, 16, 128
x = torch.randn((B, T, H), device=dev)
wf = (torch.randn((H, taps * 64), device=dev) * 0.01).bfloat16()
bias = torch.randn((H,), device=dev) * 0.1
xp = ops.cast_pad(x, taps // 2)
h = x.clone().view(B * T, H)
fl = 2.0 * B * T * H * 64 * taps
for slab in (1, 0):
    ops.POSCONV_SLAB = slab
    for _ in range(3):
        ops.posconv(xp, wf, bias, h, T, H, groups, taps, h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.posconv(xp, wf, bias, h, T, H, groups, taps, h)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{'slab kernel' if slab else 'generic BN=64 path'}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
