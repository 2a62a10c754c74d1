"""SpecAugment time masks for the training path: restatement of `_compute_mask_indices` of the un-vendored
transformers dependency (HF:101-217, called from HF:1305-1312 with the frame-level attention mask).

It consumes NumPy's GLOBAL generator exactly like the original (one `rand` for the probabilistic rounding, one
`choice` without replacement per utterance), so that `np.random.seed(s)` before a step selects the same frames here
as in the reference — pinned by tests/test_host_cpu.py against the installed transformers."""
from __future__ import annotations

import numpy as np


def compute_mask_indices(shape, mask_prob: float, mask_length: int, frame_lens=None, min_masks: int = 0) -> np.ndarray:
    """bool [B, T]: True where the hidden state is replaced by `masked_spec_embed`."""
    B, T = shape
    if mask_length < 1:
        raise ValueError("`mask_length` has to be bigger than 0.")
    if mask_length > T:
        raise ValueError(f"`mask_length` has to be smaller than `sequence_length`, but got `mask_length`: {mask_length}"
                         f" and `sequence_length`: {T}`")
    eps = np.random.rand(1).item()                      # probabilistic rounding of the span count

    def n_spans(n):
        k = max(int(mask_prob * n / mask_length + eps), min_masks)
        if k * mask_length > T:
            k = T // mask_length
        if n - (mask_length - 1) < k:
            k = max(n - (mask_length - 1), 0)
        return k

    lens = [T] * B if frame_lens is None else [int(n) for n in frame_lens]
    mask = np.zeros((B, T), dtype=bool)
    k_max = n_spans(T)
    if k_max == 0:
        return mask
    starts = []
    for n in lens:
        k = n_spans(n)
        idx = np.random.choice(np.arange(n - (mask_length - 1)), k, replace=False)
        pad = T - 1 if len(idx) == 0 else idx[0]      # pad with a start that is masked anyway (or a padding frame)
        starts.append(np.concatenate([idx, np.ones(k_max - k, dtype=np.int32) * pad]))
    starts = np.array(starts)
    spans = (starts[:, :, None] + np.arange(mask_length)[None, None, :]).reshape(B, k_max * mask_length)
    spans = np.minimum(spans, T - 1)
    np.put_along_axis(mask, spans, True, -1)
    return mask
