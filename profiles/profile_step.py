"""One APTAI.predict pass (B=16 x 8 s, 24x1024 'layer' backbone: BASELINE config-2 shape) for ncu captures.
Usage on the GPU box:  python profiles/profile_step.py   (plain)   |   ncu ... python profiles/profile_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aptai_b200 import lib  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402

B = int(os.environ.get("PROFILE_B", 16))
L = int(os.environ.get("PROFILE_L", 128000))
dev = torch.device("cuda:0")
cfg = W2V2Config.large(**bench.NO_REG)
model = bench.make_model(cfg, dev)
g = torch.Generator().manual_seed(0)
wav = torch.empty((B, L)).normal_(0.0, 0.1, generator=g).to(dev)
lens = torch.full((B,), L, dtype=torch.int64, device=dev)
lens[1::2] -= 16000
tg = (torch.arange(59, dtype=torch.int32)[None] % 45 + 1).repeat(B, 1).to(dev)
tl = torch.full((B,), 40, dtype=torch.int32, device=dev)
n0 = lib.launch_count()
r = model.predict(wav, lens, phn_targets=tg, phn_target_lens=tl)
torch.cuda.synchronize()
print("launches", lib.launch_count() - n0, "status", r["align_status"].tolist()[:4])
