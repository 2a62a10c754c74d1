"""Fused tail kernel (final LayerNorm + both heads + argmax + log-softmax) at the bench's batch sizes.
Usage: python profiles/tail_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
# A/B against another build of the library: APTAI_LIB_ALT=path/to/other/libaptai_b200.so (aptai_b200/lib.py)
from aptai_b200 import ops

dev = torch.device("cuda:0")
H = 1024
g = torch.ones((H,), device=dev); b = torch.zeros((H,), device=dev)
tvw, tvb = torch.randn((9, H), device=dev) * 0.03, torch.zeros((9,), device=dev)
pw, pb = torch.randn((46, H), device=dev) * 0.03, torch.zeros((46,), device=dev)
for rows in (47880, 75776):
    hs = [torch.randn((rows, H), device=dev) for _ in range(3)]
    fn = lambda i: ops.tail(hs[i % 3], g, b, 1e-5, tvw, tvb, ops.ACT_TANH, pw, pb, ops.ACT_LEAKY, want_logp=True)
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 12 * 1e3
    print(f"tail rows={rows}: {us:.1f} us  ({rows * 64 * H * 2 / us / 1e6:.1f} TFLOP/s fp32 of the 64-slot tile, "
          f"h read {rows * H * 4 / us / 1e3:.0f} GB/s)", flush=True)
