"""Isolated CUDA-event timings of the training-step kernels at BASELINE config-4 shapes (B=32 x 8 s, 24x1024):
wgrad (tcgen05 MN-major), dgrad epilogues, attention backward, LayerNorm backward, column sums, Adam.
L2 is flushed between iterations by the working set itself (>= 150 MB per call) plus a 256 MB scratch write."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, H, F, heads = 32, 399, 1024, 4096, 16
M = B * T
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ev(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        scratch.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def r(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(shape, device=dev) * scale).to(dtype)


out = {}
for name, (N, K) in {"wgrad_qkv": (3 * H, H), "wgrad_out": (H, H), "wgrad_ffn1": (F, H), "wgrad_ffn2": (H, F)}.items():
    dy, x = r(M, N), r(M, K)
    dw = torch.zeros((N, K), dtype=torch.float32, device=dev)
    ms = ev(lambda: ops.wgrad(dy, x, dw))
    out[name] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
    del dy, x, dw
# dgrad GEMMs
u = r(M, F)
dyH = r(M, H)
w2t, w1t, wqkvt, wot = r(F, H, scale=0.03), r(H, F, scale=0.03), r(H, 3 * H, scale=0.03), r(H, H, scale=0.03)
ms = ev(lambda: ops.linear(dyH, w2t, None, act=2, aux=u))
out["dgrad_ffn2_gelu"] = {"ms": ms, "tflops": 2.0 * M * H * F / ms / 1e9}
du = r(M, F)
ms = ev(lambda: ops.linear(du, w1t, None, want_f32=True, want_bf16=False))
out["dgrad_ffn1_f32"] = {"ms": ms, "tflops": 2.0 * M * H * F / ms / 1e9}
dqkv = r(M, 3 * H)
ms = ev(lambda: ops.linear(dqkv, wqkvt, None, want_f32=True, want_bf16=False))
out["dgrad_qkv_f32"] = {"ms": ms, "tflops": 2.0 * M * H * 3 * H / ms / 1e9}
ms = ev(lambda: ops.linear(dyH, wot, None))
out["dgrad_out_bf16"] = {"ms": ms, "tflops": 2.0 * M * H * H / ms / 1e9}
del u, du, w2t, w1t
# attention
qkv = r(M, 3 * H)
qkv[:, :H] *= 0.25
g = torch.Generator().manual_seed(0)
klen = torch.randint(200, T + 1, (B,), generator=g).to(torch.int32).to(dev)
klen[0] = T
lse = torch.empty((B, heads, T), dtype=torch.float32, device=dev)
ctx = ops.attention(qkv, klen, B, T, heads, lse=lse)
dctx = r(M, H)
fl = sum(4.0 * T * int(k) * 64 * heads for k in klen.tolist())
ms = ev(lambda: ops.attention(qkv, klen, B, T, heads, lse=lse))
out["attention_fwd"] = {"ms": ms, "tflops": fl / ms / 1e9}
ms = ev(lambda: ops.attention_bwd(qkv, ctx, dctx, lse, klen, B, T, heads, 0.125))
out["attention_bwd_total(dot+memset+kernel+cast)"] = {"ms": ms, "tflops": 2.5 * fl / ms / 1e9}
# LayerNorm backward, colsum
dy32, x32 = r(M, H, dtype=torch.float32), r(M, H, dtype=torch.float32)
gam = torch.ones(H, device=dev)
dg, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
ms = ev(lambda: ops.layernorm_bwd(dy32, x32, gam, 1e-5, dres=dy32, dgamma=dg, dbeta=db, want_bf16=True))
out["layernorm_bwd"] = {"ms": ms, "gbs": M * H * (4 * 3 + 4 + 2) / ms / 1e6}
for N in (H, 3 * H, F):
    xx = r(M, N)
    o = torch.zeros(N, device=dev)
    ms = ev(lambda: ops.colsum(xx, o))
    out[f"colsum_{N}"] = {"ms": ms, "gbs": M * N * 2 / ms / 1e6}
print(json.dumps({k: {a: round(b, 4) for a, b in v.items()} for k, v in out.items()}))
