"""Backbone geometry for the wav2vec2-style acoustic encoder.

The reference passes a `transformers.Wav2Vec2Config` as `pretrain_cfg`
(models/aptai.py:20,35; models/w2v2_pr.py:22,30).  The kernels only need the
geometry, so `W2V2Config` is a light mirror of the fields the hot path reads;
`W2V2Config.from_any` accepts an HF config object, a dict, or a `W2V2Config`.
Field names and defaults follow configuration_wav2vec2.py (transformers 5.5.0),
so the *base* variant is the default and *large* / XLS-R is `W2V2Config.large()`.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Any, Sequence, Tuple


@dataclass
class W2V2Config:
    vocab_size: int = 46
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-5
    feat_extract_norm: str = "group"          # "group" (base) | "layer" (large)
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_bias: bool = False
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    do_stable_layer_norm: bool = False
    # regularisers (training only)
    hidden_dropout: float = 0.1
    activation_dropout: float = 0.1
    attention_dropout: float = 0.1
    feat_proj_dropout: float = 0.0
    final_dropout: float = 0.1
    layerdrop: float = 0.1
    apply_spec_augment: bool = True
    mask_time_prob: float = 0.05
    mask_time_length: int = 10
    mask_time_min_masks: int = 2
    mask_feature_prob: float = 0.0            # feature-axis SpecAugment (HF:1314-1322); 0 in every config the reference uses
    mask_feature_length: int = 10
    mask_feature_min_masks: int = 0
    # CTC (train/train_phoneme_recognizer.py:336-347)
    blank: int = 0
    ctc_loss_reduction: str = "mean"
    ctc_zero_infinity: bool = True

    @staticmethod
    def base(**kw) -> "W2V2Config":
        return W2V2Config(**kw)

    @staticmethod
    def large(**kw) -> "W2V2Config":
        d = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                 intermediate_size=4096, feat_extract_norm="layer", conv_bias=True,
                 do_stable_layer_norm=True)
        d.update(kw)
        return W2V2Config(**d)

    @staticmethod
    def from_any(cfg: Any) -> "W2V2Config":
        if isinstance(cfg, W2V2Config):
            return cfg
        if isinstance(cfg, dict):
            src = cfg
        elif hasattr(cfg, "to_dict"):
            src = cfg.to_dict()
            # attributes set after construction (e.g. cfg.blank = 0) may not be in to_dict()
            for k in W2V2Config.__dataclass_fields__:
                if k not in src and hasattr(cfg, k):
                    src[k] = getattr(cfg, k)
        else:
            src = {k: getattr(cfg, k) for k in W2V2Config.__dataclass_fields__ if hasattr(cfg, k)}
        kw = {}
        for k in W2V2Config.__dataclass_fields__:
            if k in src and src[k] is not None:
                v = src[k]
                if k in ("conv_dim", "conv_stride", "conv_kernel"):
                    v = tuple(int(x) for x in v)
                kw[k] = v
        return W2V2Config(**kw)

    def to_dict(self) -> dict:
        d = asdict(self)
        for k in ("conv_dim", "conv_stride", "conv_kernel"):
            d[k] = list(d[k])
        return d

    # ---- geometry helpers -------------------------------------------------
    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def conv_out_length(self, n: int, upto: int | None = None) -> int:
        """HF:1005-1024 `_get_feat_extract_output_lengths`: floor((n-k)/s)+1 per layer."""
        ks, ss = self.conv_kernel, self.conv_stride
        upto = len(ks) if upto is None else upto
        for k, s in zip(ks[:upto], ss[:upto]):
            n = (n - k) // s + 1
        return n

    def validate_for_kernels(self) -> None:
        if self.head_dim != 64:
            raise ValueError("aptai_b200 kernels require head_dim == 64")
        if any(c != 512 for c in self.conv_dim):
            raise ValueError("aptai_b200 kernels require conv_dim == 512 for every layer")
        if self.feat_extract_norm not in ("group", "layer"):
            raise ValueError("feat_extract_norm must be 'group' or 'layer'")
        if self.hidden_size % 64 or self.intermediate_size % 64:
            raise ValueError("hidden/intermediate size must be multiples of 64")
        if self.hidden_size % self.num_conv_pos_embedding_groups:
            raise ValueError("hidden_size must divide into pos-conv groups")


def frame_lengths(cfg: W2V2Config, sample_lengths: Sequence[int]) -> list[int]:
    return [cfg.conv_out_length(int(n)) for n in sample_lengths]
