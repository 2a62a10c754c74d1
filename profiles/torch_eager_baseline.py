#!/usr/bin/env python
"""The "existing Blackwell library path" bar: the reference's own model code — transformers'
Wav2Vec2Model + the APTAI heads as models/aptai.py wires them — run by stock PyTorch on the same B200 (cuBLAS / cuDNN /
SDPA kernels, bf16 autocast, fp32 master weights), next to aptai_b200 on identical shapes:

  inference   B x 8 s utterances, 24x1024 'layer' backbone, heads + low-pass + argmax       (audio-s/s)
  training    BASELINE config 4: B=32 x <= 8 s, frozen conv encoder, Adam, no regularisers    (ms / step)

This is a reported baseline (not the product path, nothing here is imported by aptai_b200)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import Wav2Vec2Config, Wav2Vec2Model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=120)
ap.add_argument("--train-batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda:0")
kw = dict(vocab_size=46, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
          feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True, hidden_dropout=0.0,
          activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0,
          apply_spec_augment=False)
cfg = Wav2Vec2Config(**kw)


class RefAPTAI(nn.Module):                      # the wiring of models/aptai.py:33-55,75-106 on transformers
    def __init__(self):
        super().__init__()
        self.wav2vec2 = Wav2Vec2Model(cfg)
        self.wav2vec2.freeze_feature_encoder()
        self.tv_head = nn.Sequential(nn.Dropout(0.0), nn.Tanh(), nn.Linear(1024, 9))
        self.phn_head = nn.Sequential(nn.Dropout(0.0), nn.LeakyReLU(), nn.Linear(1024, 46))
        n = np.arange(51)
        h = np.sinc(2 * 10 / 49 * (n - 25)) * (0.5 * (1 - np.cos(2 * np.pi * n / 50)))
        self.register_buffer("taps", torch.tensor(h / h.sum()).view(1, 1, 51))

    def forward(self, wav, lens, phn=None, tvt=None):
        mask = (torch.arange(wav.shape[1], device=wav.device)[None, :] < lens[:, None]).long()
        h = self.wav2vec2(wav, attention_mask=mask).last_hidden_state
        tv = self.tv_head(h).float()
        B, T, C = tv.shape
        tv = F.conv1d(tv.permute(0, 2, 1).reshape(B * C, 1, T).double(), self.taps, padding="same").float()
        tv = tv.view(B, C, T).permute(0, 2, 1)
        logits = self.phn_head(h).float()
        if phn is None:
            return tv, logits.argmax(-1)
        m = tvt != -100.0
        mse = F.mse_loss(tv[m], tvt[m])
        pm = (phn != 0).flatten()
        ce = F.cross_entropy(logits.view(-1, 46)[pm], phn.flatten()[pm], ignore_index=0)
        return 0.5 * mse + 0.5 * ce


def ev(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
model = RefAPTAI().to(dev)
out = {}
# ---- inference
B, L = args.batch, 128000
wav = (0.1 * torch.randn(B, L)).to(dev)
lens = torch.full((B,), L, device=dev)
lens[1::2] -= 16000
model.eval()


def infer():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(wav, lens)


ms = ev(infer, args.iters)
out["torch_eager_bf16_inference"] = {"batch": B, "ms": ms, "audio_s_per_s": float(lens.sum()) / 16000 / (ms * 1e-3)}
try:
    import bench
    from aptai_b200.config import W2V2Config
    ours = bench.make_model(W2V2Config.large(**bench.NO_REG), dev)
    ms2 = ev(lambda: ours.predict(wav, lens), args.iters)
    out["aptai_b200_inference"] = {"batch": B, "ms": ms2, "audio_s_per_s": float(lens.sum()) / 16000 / (ms2 * 1e-3)}
    del ours
except Exception as e:                                                      # noqa: BLE001
    out["aptai_b200_inference"] = {"error": str(e)[:200]}
# ---- training step (config 4 shape)
B = args.train_batch
rng = np.random.Generator(np.random.PCG64(21))
lens_t = torch.from_numpy(rng.integers(64000, 128001, size=B)).to(dev)
lens_t[0] = L
wav_t = (0.1 * torch.randn(B, L)).to(dev)
T = 399
flen = ((((((lens_t - 10) // 5 + 1 - 3) // 2 + 1 - 3) // 2 + 1 - 3) // 2 + 1 - 3) // 2 + 1 - 2) // 2 + 1
flen = (flen - 2) // 2 + 1
phn = torch.zeros((B, T), dtype=torch.long, device=dev)
tvt = torch.full((B, T, 9), -100.0, device=dev)
for b in range(B):
    n = int(flen[b])
    phn[b, :n] = torch.randint(1, 46, (n,), device=dev)
    tvt[b, :n] = torch.randn((n, 9), device=dev)
model.train()
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5)


def train_step():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(wav_t, lens_t, phn, tvt)
    loss.backward()
    opt.step()


torch.cuda.reset_peak_memory_stats()
ms = ev(train_step, args.iters)
out["torch_eager_bf16_train_config4"] = {"batch": B, "ms_per_step": ms,
                                         "audio_s_per_s": float(lens_t.sum()) / 16000 / (ms * 1e-3),
                                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
print(json.dumps(out))
