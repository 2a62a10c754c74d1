"""Micro-benchmark of the tcgen05 GEMM family: TFLOP/s per shape, single-CTA vs CTA-pair tiles.
CUDA events on the launching stream, 3 warm-ups, L2 flushed (256 MB write) between timed launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(49152, 1024, 3072, "qkv"), (49152, 1024, 1024, "out"), (49152, 1024, 4096, "ffn1"), (49152, 4096, 1024, "ffn2"),
          (6384, 1024, 3072, "qkv-b16"), (6384, 4096, 1024, "ffn2-b16"), (8192, 8192, 8192, "square")]
for M, K, N, name in shapes:
    a = torch.randn((M, K), device=dev).bfloat16()
    w = (torch.randn((N, K), device=dev) * 0.05).bfloat16()
    b = torch.randn((N,), device=dev)
    out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    res = {}
    act = 1 if name.startswith("ffn1") else 0
    resid = torch.randn((M, N), device=dev) if name.startswith(("out", "ffn2")) else None
    for pair in (1, 2):
        if resid is not None:
            ms = timeit(lambda: ops.linear(a, w, b, residual=resid, out_f32=resid, want_bf16=False, cta_pair=pair))
        else:
            ms = timeit(lambda: ops.linear(a, w, b, out_bf16=out, act=act, cta_pair=pair))
        res[pair] = 2.0 * M * N * K / ms / 1e9
    ms = timeit(lambda: torch.nn.functional.linear(a, w, b.bfloat16()))
    print(f"{name:10s}{" +gelu" if act else (" +res32" if resid is not None else "")} M={M:6d} K={K:5d} N={N:5d}  single {res[1]:7.1f}  pair {res[2]:7.1f}  cublas {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s")
