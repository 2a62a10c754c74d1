#!/usr/bin/env python
"""conv layer 0 + LayerNorm / GroupNorm + GELU (csrc/frontend.cu) at B=120 x 8 s: microseconds per launch and the
fraction of the HBM floor (waveform read once, fp16 output written once)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aptai_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, L = 120, 128000
g = torch.Generator().manual_seed(0)
wav = torch.empty((B, L)).normal_(0.0, 0.1, generator=g).to(dev)
w = (torch.randn((512, 10), generator=g) * 0.3).to(dev)
bias, gamma, beta = (torch.randn((512,), generator=g).to(dev) * 0.1 for _ in range(3))
gamma = gamma + 1
out = {}
for norm, name in ((1, "layer"), (2, "group")):
    fn = lambda: ops.conv0(wav, w, bias, gamma, beta, norm, out_dtype=torch.float16)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    T0 = (L - 10) // 5 + 1
    bytes_ = B * (L * 4 + T0 * 512 * 2)
    out[name] = {"us": us, "GBps": bytes_ / us / 1e3, "frac_of_6451": bytes_ / us / 1e3 / 6451.5}
print(json.dumps(out))
