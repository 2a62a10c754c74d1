"""PyTorch-CPU fp32 restatement of transformers.Wav2Vec2Model.forward in eval mode — TEST INFRASTRUCTURE ONLY.

The reference's acoustic encoder lives in the un-vendored dependency `transformers` (no version pinned by the
reference; 5.5.0 is what the image ships, "HF:n" = models/wav2vec2/modeling_wav2vec2.py line n).  Reference call
sites: models/aptai.py:75-81,135-141; models/w2v2_pr.py:47-52,93-101,132-140,179-186.

`forward(sd, cfg, wav, lengths)` takes a state dict with HF key names and returns the tuple of N+1 hidden states
exactly as `Wav2Vec2Model(..., attention_mask=lengths[:, None], output_hidden_states=True)` does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def conv_out_length(n, cfg):
    """HF:1005-1024 `_get_feat_extract_output_lengths`."""
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        n = (n - k) // s + 1
    return n


def feature_encoder(sd, cfg, wav):
    """HF:409-419 with layer types HF:254-323.  wav [B,L] -> [B,512,T]."""
    x = wav[:, None]
    for i, (k, s) in enumerate(zip(cfg.conv_kernel, cfg.conv_stride)):
        p = f"feature_extractor.conv_layers.{i}."
        x = F.conv1d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), stride=s)
        if cfg.feat_extract_norm == "layer":
            x = F.layer_norm(x.transpose(-2, -1), (x.shape[1],), sd[p + "layer_norm.weight"],
                             sd[p + "layer_norm.bias"], 1e-5).transpose(-2, -1)          # HF:291-299
        elif i == 0:
            x = F.group_norm(x, x.shape[1], sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)  # HF:308-323
        x = F.gelu(x)
    return x


def pos_conv_weight(sd):
    """torch.nn.utils.parametrizations.weight_norm(dim=2): w = g * v / ||v||_(0,1)  (HF:336-352)."""
    g = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return g * v / torch.linalg.vector_norm(v, dim=(0, 1), keepdim=True)


def attention(sd, p, cfg, x, key_mask, prob_mask=None):
    """HF:500-549 with SDPA semantics: softmax(q k^T / sqrt(d) + key-padding mask) v, fp32; `prob_mask` [B,nh,T,T]
    (already scaled by 1/(1-p)) replays dropout on the attention probabilities (HF:461)."""
    B, T, H = x.shape
    nh = cfg.num_attention_heads
    d = H // nh
    q = F.linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    k = F.linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    v = F.linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"]).view(B, T, nh, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    if prob_mask is not None:
        a = a * prob_mask
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, H)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def feed_forward(sd, p, x, act_mask=None):
    """HF:566-573 (intermediate_dropout after the activation, HF:570)."""
    x = F.gelu(F.linear(x, sd[p + "intermediate_dense.weight"], sd[p + "intermediate_dense.bias"]))
    if act_mask is not None:
        x = x * act_mask
    return F.linear(x, sd[p + "output_dense.weight"], sd[p + "output_dense.bias"])


def forward(sd, cfg, wav, lengths, return_features=False, reg=None):
    """HF:1327-1383.  wav fp32 [B,L], lengths int [B] (samples).  Returns tuple of N+1 hidden states [B,T,H].

    `reg` (training mode): explicit regulariser draws so that a stochastic step can be replayed — dict with optional
    keys 'proj' (HF:434), 'enc' (HF:694/766), ('attn', l) (HF:603/647), ('act', l) (HF:570), ('ffn', l) (HF:573),
    ('attp', l) (HF:461, attention probabilities [B,nh,T,T]):
    multiplicative masks already scaled by 1/(1-p), broadcastable to the activation; 'skip': set of layers dropped by
    LayerDrop (HF:701-706/773-778); 'spec': bool [B,T] SpecAugment mask (HF:1303, rows replaced by masked_spec_embed);
    'spec_feat': bool [B,H] feature-axis mask (HF:1314-1322, channels zeroed for the whole utterance)."""
    reg = reg or {}
    m = lambda key, x: x if reg.get(key) is None else x * reg[key].view(x.shape)
    eps = cfg.layer_norm_eps
    feats = feature_encoder(sd, cfg, wav).transpose(1, 2)                       # HF:1348-1349  [B,T,512]
    B, T, _ = feats.shape
    flen = torch.as_tensor([conv_out_length(int(n), cfg) for n in lengths])     # HF:1026-1044
    mask = torch.arange(T)[None, :] < flen[:, None]
    x = F.layer_norm(feats, (feats.shape[-1],), sd["feature_projection.layer_norm.weight"],
                     sd["feature_projection.layer_norm.bias"], eps)             # HF:431
    x = F.linear(x, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])
    x = m("proj", x)
    if reg.get("spec") is not None:
        x = torch.where(reg["spec"][:, :, None], sd["masked_spec_embed"][None, None, :], x)
    if reg.get("spec_feat") is not None:
        x = x * (~reg["spec_feat"])[:, None, :].to(x.dtype)
    x = x.clone()
    x[~mask] = 0.0                                                              # HF:679-682 / 753-756
    key_mask = None if bool(mask.all()) else mask
    K = cfg.num_conv_pos_embeddings
    pos = F.conv1d(x.transpose(1, 2), pos_conv_weight(sd), sd["encoder.pos_conv_embed.conv.bias"], padding=K // 2,
                   groups=cfg.num_conv_pos_embedding_groups)
    if K % 2 == 0:
        pos = pos[:, :, :-1]                                                    # HF:371-379
    x = x + F.gelu(pos).transpose(1, 2)                                         # HF:689-690 / 763-764
    H = x.shape[-1]
    hidden = []
    if not cfg.do_stable_layer_norm:
        x = F.layer_norm(x, (H,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], eps)   # HF:692
        x = m("enc", x)
        for l in range(cfg.num_hidden_layers):
            hidden.append(x)
            if l in reg.get("skip", ()):
                continue
            p = f"encoder.layers.{l}."
            x = x + m(("attn", l), attention(sd, p + "attention.", cfg, x, key_mask, reg.get(("attp", l))))   # HF:592-609 (post-LN)
            x = F.layer_norm(x, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps)
            x = x + m(("ffn", l), feed_forward(sd, p + "feed_forward.", x, reg.get(("act", l))))
            x = F.layer_norm(x, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
        hidden.append(x)
    else:
        x = m("enc", x)
        for l in range(cfg.num_hidden_layers):
            hidden.append(x)
            if l in reg.get("skip", ()):
                continue
            p = f"encoder.layers.{l}."
            y = F.layer_norm(x, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], eps)
            x = x + m(("attn", l), attention(sd, p + "attention.", cfg, y, key_mask, reg.get(("attp", l))))   # HF:632-655 (pre-LN)
            y = F.layer_norm(x, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
            x = x + m(("ffn", l), feed_forward(sd, p + "feed_forward.", y, reg.get(("act", l))))
        x = F.layer_norm(x, (H,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], eps)   # HF:792
        hidden.append(x)
    if return_features:
        return tuple(hidden), feats, flen
    return tuple(hidden)


def flops_forward(cfg, L: int, heads_out: int = 55) -> float:
    """Algorithmic forward FLOPs of one utterance of L samples (SURVEY.md §8d closed form)."""
    H, Fi, N = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    t = L
    f = 0.0
    cin = 1
    for k, s, c in zip(cfg.conv_kernel, cfg.conv_stride, cfg.conv_dim):
        t = (t - k) // s + 1
        f += 2.0 * c * cin * k * t
        cin = c
    T = t
    f += 2.0 * T * cin * H
    f += 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    f += N * T * (8.0 * H * H + 4.0 * H * Fi)
    f += N * 4.0 * T * T * H
    f += 2.0 * T * H * heads_out + 2.0 * T * 9 * 51
    return f
