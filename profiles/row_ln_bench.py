"""Fused row LayerNorm behind the in-place residual update (ops.linear(row_ln=...)) against the two-launch form, on the
out-proj (K = 1024) and FFN2 (K = 4096) shapes.  Buffers rotate so that nothing is L2-resident from the previous call.
Usage: python profiles/row_ln_bench.py [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aptai_b200 import ops


def timeit(fn, n=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
    dev = torch.device("cuda:0")
    N = 1024
    g = torch.ones((N,), device=dev); e = torch.zeros((N,), device=dev); b = torch.zeros((N,), device=dev)
    for name, K in (("out-proj", 1024), ("FFN2", 4096)):
        NB = 3
        a = [torch.randn((M, K), device=dev).bfloat16() for _ in range(NB)]
        h = [torch.randn((M, N), device=dev) for _ in range(NB)]
        x = [torch.empty((M, N), dtype=torch.bfloat16, device=dev) for _ in range(NB)]
        w = (torch.randn((N, K), device=dev) * 0.02).bfloat16()
        k = [0]

        def nxt():
            k[0] = (k[0] + 1) % NB
            return k[0]

        def plain():
            i = nxt(); ops.linear(a[i], w, b, residual=h[i], out_f32=h[i], want_bf16=False)

        def two():
            i = nxt(); ops.linear(a[i], w, b, residual=h[i], out_f32=h[i], want_bf16=False)
            ops.layernorm(h[i], g, e, 1e-5, out_bf16=x[i])

        def fused(eps=1e-5):
            i = nxt(); ops.linear(a[i], w, b, residual=h[i], out_f32=h[i], want_bf16=False, row_ln=(g, e, eps, x[i]))

        def ln_only():
            i = nxt(); ops.layernorm(h[i], g, e, 1e-5, out_bf16=x[i])

        t = {"gemm": timeit(plain), "gemm + LN launch": timeit(two), "LN launch": timeit(ln_only),
             "fused": timeit(fused), "fused, arrivals only": timeit(lambda: fused(-1.0))}
        print(f"{name} M={M} K={K}: " + "  ".join(f"{k_} {v:.1f} us" for k_, v in t.items()), flush=True)
        del a, h, x


if __name__ == "__main__":
    main()
