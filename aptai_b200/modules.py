"""Drop-in counterparts of the reference's models/modules.py building blocks, on the aptai_b200 kernels.

Same class names, constructor signatures, parameter/buffer names and shapes (so reference checkpoints load with
strict=True), same forward semantics — minus the reference's defects listed in SURVEY.md Appendix B.
`ConvBank` (models/modules.py:156-187) is dead code in the reference and is not provided.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops


class LowPassFilterLayer(nn.Module):
    """models/modules.py:13-61.  51-tap Hann-windowed sinc FIR, fc = cutoff/sampling_rate, taps normalised to
    sum 1, applied to every channel independently with zero 'same' padding, fp64 accumulation, fp32 result.
    The reference loops over channels with nine conv1d launches and stages the result in a CPU tensor
    (modules.py:55-61); here it is one kernel and stays on the device."""

    def __init__(self, device, cutoff, sampling_rate, out_dim=9):
        super().__init__()
        self.device = device
        self.out_dim = out_dim
        w = self._get_filter_weights(cutoff, sampling_rate).view(1, 1, -1)
        # parameter container with the reference's quirky shape: Conv1d(1, out_dim) whose weight is replaced by
        # a (1,1,N) float64 tensor (modules.py:24-25) -> state_dict key 'lowpass.weight' (1,1,51) float64
        self.lowpass = nn.Conv1d(1, self.out_dim, self.N, stride=1, padding="same", bias=False)
        self.lowpass.weight = nn.Parameter(w, requires_grad=False)

    def _get_filter_weights(self, cutoff, sampling_rate):
        fc = cutoff / sampling_rate
        if fc > 0.5:
            raise Exception("Cutoff frequency must be at least twice the sampling rate.")
        b = 0.08
        N = int(np.ceil(4 / b))
        if not N % 2:
            N += 1
        self.N = N
        n = np.arange(N)
        h = np.sinc(fc * 2 * (n - (N - 1) / 2))
        w = 0.5 * (1 - np.cos(n * 2 * math.pi / (N - 1)))
        h = h * w
        h = h / np.sum(h)
        return torch.tensor(h, device=self.device)

    def forward(self, y: Tensor) -> Tensor:
        taps = self.lowpass.weight.detach().reshape(-1).to(device=y.device, dtype=torch.float64).contiguous()
        return ops.lowpass(y.detach().float().contiguous(), taps)


class ForwardSumLoss(nn.Module):
    """models/modules.py:65-117.  Per utterance: prepend a blank column of log-prob -1, keep the first
    text_len+1 columns and mel_len rows, log_softmax, CTC (zero_infinity, mean over the target length) against
    the identity target 1..text_len; average over the batch.  The reference runs B separate CTC launches in a
    Python loop; here it is one launch of the fused log-softmax+CTC kernel."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.blank_logprob = blank_logprob

    def forward(self, attn_logprob, text_lens, mel_lens):
        B, one, T, N = attn_logprob.shape
        dev = attn_logprob.device
        text = torch.as_tensor(text_lens, dtype=torch.int32, device=dev).reshape(B).contiguous()
        mel = torch.as_tensor(mel_lens, dtype=torch.int32, device=dev).reshape(B).contiguous()
        tg = torch.arange(1, N + 1, dtype=torch.int32, device=dev)[None].expand(B, N).contiguous()
        scale = (1.0 / (text.clamp(min=1).float() * B)).contiguous()
        r = ops.logsoftmax_ctc(attn_logprob.detach().float().reshape(B, T, N).contiguous(), tg, mel, text, blank=0,
                               zero_infinity=True, scale=scale, want_log_probs=False, prepend_blank=True,
                               blank_value=float(self.blank_logprob), vocab_len=(text + 1).contiguous())
        return r["loss_sum"][0]


class CrossAttention(nn.Module):
    """models/modules.py:129-153: q = W_q frame, k = W_k phn, energy = q k^T (unscaled) - 1000*pad,
    out = LayerNorm(cat[softmax(energy) k, q]).  One fused kernel (csrc/xattn.cu) for the 128-wide block with up to
    60 phoneme slots that Force_APTAI instantiates; returns (att_out, energy) like the reference."""

    def __init__(self, frame_dim, phn_dim, att_dim):
        super().__init__()
        self.q = nn.Linear(frame_dim, att_dim)
        self.k = nn.Linear(phn_dim, att_dim)
        self.layer_norm = nn.LayerNorm(att_dim * 2)

    def forward(self, frame_hidden, phn_hidden, labels_att_mask):
        if not (frame_hidden.shape[-1] == phn_hidden.shape[-1] == self.q.out_features == 128
                and phn_hidden.shape[1] == 60):
            raise NotImplementedError("aptai_b200 CrossAttention kernel is built for the 128-dim / 60-slot block of "
                                      "Force_APTAI (models/force_aptai.py:28-41)")
        att_out, energy, _ = ops.cross_attention(
            frame_hidden.detach().float().contiguous(), labels_att_mask.to(torch.int32).contiguous(), None, None,
            self.q.weight, self.q.bias, self.k.weight, self.k.bias, self.layer_norm.weight, self.layer_norm.bias,
            self.layer_norm.eps, phn_hidden=phn_hidden.detach().float().contiguous())
        return att_out, energy


class RNN(nn.Module):
    """models/modules.py:190-214: BiLSTM(256) -> Linear -> Dropout -> Tanh -> Linear(9).  The reference's packed
    path for batch > 1 raises NameError (`packed_putput`, modules.py:207); the intended `packed_output`
    semantics are implemented (per-utterance lengths, zeros beyond them; for batch 1 the reference runs the
    whole padded row, which is the same thing with len = T).
    Kernels: input projection = fp32-accurate tcgen05 GEMM, recurrence = persistent cluster kernel
    (csrc/lstm.cu), Linear(512,256) = tcgen05 GEMM, Tanh + Linear(256,9) = the heads kernel."""

    def __init__(self, hidden_dim, out_dim, drop=0.1):
        super().__init__()
        self.lstm = nn.LSTM(hidden_dim, hidden_dim, bidirectional=True, num_layers=1, batch_first=True)
        self.linear = nn.Sequential(nn.Linear(2 * hidden_dim, hidden_dim), nn.Dropout(drop), nn.Tanh(),
                                    nn.Linear(hidden_dim, out_dim))

    @torch.no_grad()
    def forward(self, embeddings, lens):
        if self.training and self.linear[1].p > 0:
            raise NotImplementedError("aptai_b200: as a standalone module RNN runs forward-only (no autograd graph); its "
                                      "training path (dropout, BiLSTM backward through time) is driven by "
                                      "Force_APTAI.forward in train mode — call .eval() for inference")
        x = embeddings.detach().float().contiguous()
        B, T, D = x.shape
        if B > 1:
            ln = torch.as_tensor(lens, dtype=torch.int32).reshape(B).to(x.device).contiguous()
        else:
            ln = torch.full((1,), T, dtype=torch.int32, device=x.device)
        hidden = ops.bilstm_256(x, self.lstm, ln)
        l0, l3 = self.linear[0], self.linear[3]
        f = lambda p: p.detach().float().contiguous()
        t = ops.linear_f32x3(hidden.view(B * T, -1), l0.weight, f(l0.bias))
        out, _, _ = ops.heads(t, f(l3.weight), f(l3.bias), ops.ACT_TANH, None, None, 0, want_argmax=False)
        return out.view(B, T, -1), hidden


class PositionalEncoding(nn.Module):
    """models/modules.py:217-235."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 60):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x: Tensor) -> Tensor:
        x = x + self.pe[: x.size(0)]
        return self.dropout(x)
