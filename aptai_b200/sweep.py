"""Utterance bucketing and sharding for batch inference (BASELINE config 5).

The path shards by utterance with no collective: sort by frame count, cut into buckets whose frame counts differ
by at most `bucket_width` (padding stays below a few percent), split buckets into batches of at most `max_rows`
frames, and give batches to ranks by longest-processing-time-first on the forward FLOP estimate.
Pure host logic (no torch device code) so that it is testable on CPU.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List, Sequence

from .config import W2V2Config


@dataclass
class Batch:
    indices: List[int]        # utterance ids
    samples: int              # padded length L of the batch (samples)
    frames: int               # padded T
    flops: float              # estimated forward FLOPs (valid frames)


def synth_durations(n: int, lo_s: float = 2.0, hi_s: float = 20.0, seed: int = 0, sr: int = 16000) -> List[int]:
    """BASELINE config 5: n utterance lengths in samples, duration U[lo_s, hi_s] from random.Random(seed)."""
    rng = random.Random(seed)
    return [int(rng.uniform(lo_s, hi_s) * sr) for _ in range(n)]


def flops_utt(cfg: W2V2Config, L: int) -> float:
    """SURVEY.md §8d closed form (forward, one utterance of L samples)."""
    H, F, N = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    t, cin, f = L, 1, 0.0
    for k, s, c in zip(cfg.conv_kernel, cfg.conv_stride, cfg.conv_dim):
        t = (t - k) // s + 1
        f += 2.0 * c * cin * k * t
        cin = c
    T = t
    f += 2.0 * T * cin * H + 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    f += N * T * (8.0 * H * H + 4.0 * H * F) + N * 4.0 * T * T * H
    f += 2.0 * T * H * 55 + 2.0 * T * 9 * 51
    return f


def gemm_flops_utt(cfg: W2V2Config, L: int) -> float:
    """FLOPs of the tcgen05 GEMM family only (conv1..6, projection, pos-conv, QKV/out/FFN) for one utterance."""
    H, F, N = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    t, cin, f = L, 1, 0.0
    for i, (k, s, c) in enumerate(zip(cfg.conv_kernel, cfg.conv_stride, cfg.conv_dim)):
        t = (t - k) // s + 1
        if i > 0:
            f += 2.0 * c * cin * k * t
        cin = c
    T = t
    f += 2.0 * T * cin * H + 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    f += N * T * (8.0 * H * H + 4.0 * H * F)
    return f


def make_batches(cfg: W2V2Config, lengths: Sequence[int], bucket_width: int = 32, max_rows: int = 49152) -> List[Batch]:
    order = sorted(range(len(lengths)), key=lambda i: lengths[i])
    batches: List[Batch] = []
    cur: List[int] = []
    t_lo = None
    for i in order:
        T = cfg.conv_out_length(lengths[i])
        if cur and (T - t_lo > bucket_width or (len(cur) + 1) * T > max_rows):
            batches.append(_close(cfg, cur, lengths))
            cur = []
        if not cur:
            t_lo = T
        cur.append(i)
    if cur:
        batches.append(_close(cfg, cur, lengths))
    return batches


def _close(cfg, idx, lengths) -> Batch:
    L = max(lengths[i] for i in idx)
    return Batch(list(idx), L, cfg.conv_out_length(L), sum(flops_utt(cfg, lengths[i]) for i in idx))


def shard_lpt(batches: Sequence[Batch], world: int, cfg: W2V2Config = None, lengths: Sequence[int] = None,
              tol: float = 0.003) -> List[List[Batch]]:
    """Longest-processing-time-first assignment of batches to ranks.

    With `cfg` and `lengths` the assignment is refined: whole batches leave an imbalance of a few percent once there
    are only 4-5 of them per rank (36 batches over 8 ranks: 1.024 max / mean), so the most loaded rank hands the tail
    of one of its batches — utterances of the same length bucket: no extra padding — to the least loaded rank as a
    batch of its own, until max / mean <= 1 + tol.  Every utterance stays in exactly one batch of exactly one rank."""
    loads = [0.0] * world
    out: List[List[Batch]] = [[] for _ in range(world)]
    for b in sorted(batches, key=lambda b: -b.flops):
        r = min(range(world), key=lambda r: loads[r])
        out[r].append(b)
        loads[r] += b.flops
    if cfg is None or lengths is None or world < 2:
        return out
    for _ in range(4 * world):
        mean = sum(loads) / world
        hi = max(range(world), key=lambda r: loads[r])
        lo = min(range(world), key=lambda r: loads[r])
        if mean <= 0 or loads[hi] <= (1.0 + tol) * mean:
            break
        want = 0.5 * (loads[hi] - loads[lo])
        # the batch of the loaded rank whose utterances are cheapest: the finest granularity for the hand-over
        src = min((b for b in out[hi] if len(b.indices) > 1), key=lambda b: b.flops / len(b.indices), default=None)
        if src is None:
            break
        moved, got = [], 0.0
        while len(src.indices) - len(moved) > 1:
            f = flops_utt(cfg, lengths[src.indices[-1 - len(moved)]])
            if got + f > want and moved:
                break
            moved.append(src.indices[-1 - len(moved)])
            got += f
            if got >= want:
                break
        if not moved:
            break
        keep = src.indices[: len(src.indices) - len(moved)]
        out[hi][out[hi].index(src)] = _close(cfg, keep, lengths)
        out[lo].append(_close(cfg, list(reversed(moved)), lengths))
        loads[hi] -= got
        loads[lo] += got
    return out
