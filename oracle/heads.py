"""CPU restatement of the APTAI / Wav2Vec2_PR / Force_APTAI module arithmetic above the backbone —
TEST INFRASTRUCTURE ONLY.  fp32 torch-CPU for the floating-point stages (fp64 where the reference uses fp64).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def lowpass_taps(cutoff=10, sampling_rate=49):
    """models/modules.py:27-44: 51-tap Hann-windowed sinc, normalised to sum 1, float64."""
    fc = cutoff / sampling_rate
    if fc > 0.5:
        raise Exception("Cutoff frequency must be at least twice the sampling rate.")
    b = 0.08
    N = int(np.ceil(4 / b))
    if not N % 2:
        N += 1
    n = np.arange(N)
    h = np.sinc(fc * 2 * (n - (N - 1) / 2))
    w = 0.5 * (1 - np.cos(n * 2 * math.pi / (N - 1)))
    h = h * w
    h = h / np.sum(h)
    return torch.tensor(h)


def lowpass(y, taps):
    """models/modules.py:46-61: per channel conv1d(padding='same') in float64, result cast to float32."""
    B, L, C = y.shape
    yd = y.double().permute(0, 2, 1).reshape(B * C, 1, L)
    out = F.conv1d(yd, taps.view(1, 1, -1).double(), padding="same")
    return out.view(B, C, L).permute(0, 2, 1).float().contiguous()


def aptai_heads(h, tv_w, tv_b, phn_w, phn_b, taps):
    """models/aptai.py:43-55,83-86 (eval: dropout is identity)."""
    tv_raw = F.linear(torch.tanh(h), tv_w, tv_b)
    tv = lowpass(tv_raw, taps)
    logits = F.linear(F.leaky_relu(h), phn_w, phn_b)
    return tv_raw, tv, logits


def aptai_losses(tv, logits, phn_frames, tv_targets):
    """models/aptai.py:67-102."""
    tv_mask = tv_targets != -100.0
    phn_mask = phn_frames != 0
    mse = F.mse_loss(tv[tv_mask], tv_targets[tv_mask], reduction="mean")
    ce = F.cross_entropy(logits.view(-1, logits.size(2))[phn_mask.flatten()], phn_frames.flatten()[phn_mask.flatten()],
                         ignore_index=0, reduction="mean")
    return 0.5 * mse + 0.5 * ce, mse, ce


def positional_encoding(d_model=128, max_len=60):
    """models/modules.py:217-227."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, 1, d_model)
    pe[:, 0, 0::2] = torch.sin(position * div_term)
    pe[:, 0, 1::2] = torch.cos(position * div_term)
    return pe


def cross_attention(frame_hidden, phn_hidden, mask, qw, qb, kw, kb, lnw, lnb):
    """models/modules.py:139-153: unscaled energy, -1000 padding mask, values = keys, LayerNorm(cat)."""
    q = F.linear(frame_hidden, qw, qb)
    k = F.linear(phn_hidden, kw, kb)
    energy = torch.bmm(q, k.transpose(2, 1))
    energy = energy + ((1 - mask) * -1000.0).unsqueeze(1)
    att = torch.softmax(energy, dim=-1)
    out = torch.cat([torch.bmm(att, k), q], dim=-1)
    out = F.layer_norm(out, (out.shape[-1],), lnw, lnb, 1e-5)
    return out, energy
