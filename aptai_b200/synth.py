"""Deterministic synthetic weights and inputs for benchmarks, smoke runs and tests.

NumPy's PCG64 stream is stable across platforms and versions, so the build container (where the golden fixtures
are produced with the reference's own classes) and the GPU box regenerate bit-identical tensors from a seed.
Distributions follow HF `_init_weights` (HF:967-1003): Linear N(0, 0.02), conv kaiming-normal, pos-conv
N(0, 2*sqrt(1/(k*C_in))), projection U(+-sqrt(1/512)), norms (1, 0) — with small random perturbations on the
norm affine parameters and biases so that every parameter participates in parity checks.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _normal(rng, shape, std):
    return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))


def _uniform(rng, shape, bound):
    return torch.from_numpy(((rng.random(shape, dtype=np.float32) * 2 - 1) * np.float32(bound)))


def backbone_state_dict(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """State dict with transformers.Wav2Vec2Model key names, all fp32."""
    rng = _rng(seed)
    H, F = cfg.hidden_size, cfg.intermediate_size
    sd = OrderedDict()
    sd["masked_spec_embed"] = _uniform(rng, (H,), 1.0)
    cin = 1
    for i, (cout, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        p = f"feature_extractor.conv_layers.{i}."
        sd[p + "conv.weight"] = _normal(rng, (cout, cin, k), math.sqrt(2.0 / (cin * k)))
        if cfg.conv_bias:
            sd[p + "conv.bias"] = _uniform(rng, (cout,), math.sqrt(1.0 / (cin * k)))
        if cfg.feat_extract_norm == "layer" or i == 0:
            sd[p + "layer_norm.weight"] = 1.0 + _normal(rng, (cout,), 0.05)
            sd[p + "layer_norm.bias"] = _normal(rng, (cout,), 0.05)
        cin = cout
    sd["feature_projection.layer_norm.weight"] = 1.0 + _normal(rng, (cin,), 0.05)
    sd["feature_projection.layer_norm.bias"] = _normal(rng, (cin,), 0.05)
    kk = math.sqrt(1.0 / cin)
    sd["feature_projection.projection.weight"] = _uniform(rng, (H, cin), kk)
    sd["feature_projection.projection.bias"] = _uniform(rng, (H,), kk)
    K, G = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    sd["encoder.pos_conv_embed.conv.bias"] = _normal(rng, (H,), 0.02)
    v = _normal(rng, (H, H // G, K), 2 * math.sqrt(1.0 / (K * H)))
    # weight_norm(dim=2): g initialised to the per-tap norm of v, then perturbed so that g != ||v||
    g = v.double().pow(2).sum(dim=(0, 1), keepdim=True).sqrt().float() * (1.0 + _normal(rng, (1, 1, K), 0.05))
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = g
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = v
    sd["encoder.layer_norm.weight"] = 1.0 + _normal(rng, (H,), 0.05)
    sd["encoder.layer_norm.bias"] = _normal(rng, (H,), 0.05)
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            sd[p + f"attention.{nm}.weight"] = _normal(rng, (H, H), 0.02)
            sd[p + f"attention.{nm}.bias"] = _normal(rng, (H,), 0.02)
        sd[p + "layer_norm.weight"] = 1.0 + _normal(rng, (H,), 0.05)
        sd[p + "layer_norm.bias"] = _normal(rng, (H,), 0.05)
        sd[p + "feed_forward.intermediate_dense.weight"] = _normal(rng, (F, H), 0.02)
        sd[p + "feed_forward.intermediate_dense.bias"] = _normal(rng, (F,), 0.02)
        sd[p + "feed_forward.output_dense.weight"] = _normal(rng, (H, F), 0.02)
        sd[p + "feed_forward.output_dense.bias"] = _normal(rng, (H,), 0.02)
        sd[p + "final_layer_norm.weight"] = 1.0 + _normal(rng, (H,), 0.05)
        sd[p + "final_layer_norm.bias"] = _normal(rng, (H,), 0.05)
    return sd


def linear_params(rng_seed: int, out_f: int, in_f: int):
    """torch nn.Linear default-like init: U(+-1/sqrt(in))."""
    rng = _rng(rng_seed)
    b = 1.0 / math.sqrt(in_f)
    return _uniform(rng, (out_f, in_f), b), _uniform(rng, (out_f,), b)


def waveforms(B: int, L: int, lengths=None, seed: int = 1234) -> torch.Tensor:
    """0.1*N(0,1) fp32 [B, L], zeroed beyond lengths[b]."""
    x = torch.empty((B, L), dtype=torch.float32)
    for b in range(B):
        x[b] = _normal(_rng(seed + b), (L,), 0.1)
        if lengths is not None:
            x[b, int(lengths[b]):] = 0.0
    return x


def phoneme_sequences(B: int, lo: int, hi: int, vocab_lo: int, vocab_hi: int, seed: int, pad: int = -100):
    """int32 [B, hi] label matrix padded with `pad`, lengths U{lo..hi}, ids U{vocab_lo..vocab_hi}."""
    rng = _rng(seed)
    lens = rng.integers(lo, hi + 1, size=B)
    out = np.full((B, hi), pad, dtype=np.int32)
    for b in range(B):
        out[b, : lens[b]] = rng.integers(vocab_lo, vocab_hi + 1, size=lens[b])
    return torch.from_numpy(out), torch.from_numpy(lens.astype(np.int32))
