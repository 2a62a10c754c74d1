#!/usr/bin/env python
"""BiLSTM recurrence kernels (csrc/lstm.cu) at Force_APTAI's BASELINE config-3 shape: 64 utterances x 399 frames.
Prints microseconds per launch and per recurrence step for the forward (inference and training variants) and the
backward through time, CUDA events, 20 launches after 3 warm-ups."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aptai_b200 import ops  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    B, T = 64, 399
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(256, 256, bidirectional=True, num_layers=1, batch_first=True).to(dev)
    x = torch.randn((B, T, 256), device=dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    fwd = timed(lambda: ops.bilstm_256(x, lstm, lens))
    fwd_t = timed(lambda: ops.bilstm_256(x, lstm, lens, save=True))
    _, sv = ops.bilstm_256(x, lstm, lens, save=True)
    dh = torch.randn((B, T, 512), device=dev)
    grads = {n: torch.zeros_like(p) for n, p in lstm.named_parameters()}
    bwd = timed(lambda: ops.bilstm_256_bwd(sv, dh, grads))
    print(json.dumps({"shape": f"B={B} T={T} (16 clusters of 8 CTAs)", "fwd_us": fwd, "fwd_train_us": fwd_t,
                      "bwd_us": bwd, "fwd_us_per_step": fwd / T, "bwd_us_per_step": bwd / T,
                      "note": "whole op incl. the input-projection GEMM (fwd) / the gradient GEMMs (bwd)"}))


if __name__ == "__main__":
    main()
