"""TEST INFRASTRUCTURE ONLY (never imported by aptai_b200/): CPU restatement of the reference's CTC decode stage.

The reference decodes phoneme logits with `torchaudio.models.decoder.ctc_decoder(lexicon=None, lm=None, nbest=1,
beam_size=10, beam_size_token=None, beam_threshold=50, blank_token='(blank)', sil_token='(...)')`
(models/w2v2_pr.py:143-159,209-229,257-272; utility.py:448-471).  That is two layers:

  1. flashlight-text's `LexiconFreeDecoder` (C++; third-party, NOT vendored in /root/reference and NOT installable in
     this image — no version is pinned by the reference; torchaudio 2.11 builds against flashlight-text 0.0.x).
     `flashlight_lexfree_decode` restates its published algorithm (LexiconFreeDecoder.cpp: decodeBegin / decodeStep /
     decodeEnd / getAllFinalHypothesis with `ZeroLM`, CTC criterion, logAdd = false).  PARITY UNPINNED for this
     layer: there is no flashlight binary here to run it against.
  2. torchaudio's Python wrapper (`_ctc_decoder.py:248-262`, `_get_tokens` / `_get_timesteps`), restated line by
     line in `torchaudio_tokens` / `torchaudio_timesteps` (the module itself cannot be imported: its first import is
     flashlight).

Consequences the drop-in reproduces (pinned by tests/test_oracle_cpu.py against this restatement):
  * the raw token path flashlight returns has T+2 entries: a leading `sil` (decodeBegin's root hypothesis), the
    per-frame labels, a trailing `sil` (decodeEnd) — so the collapsed token list starts and ends with the silence
    id unless the path itself starts / ends in silence, and `timesteps` are frame index + 1;
  * with no LM, max-merge and every token a candidate at every frame, the best hypothesis is the frame-wise argmax
    path (the best single alignment is never pruned: it is the top-scoring candidate at every frame), so the beam
    search and the greedy path agree wherever the per-frame argmax is unique;
  * the reference passes no lengths: padded frames are decoded too.
"""
from __future__ import annotations

import itertools as it

import numpy as np


def torchaudio_tokens(raw, blank: int) -> np.ndarray:
    """_ctc_decoder.py:248-251: groupby, then drop blanks."""
    idxs = (g[0] for g in it.groupby(list(raw)))
    return np.asarray([x for x in idxs if x != blank], dtype=np.int64)


def torchaudio_timesteps(raw, blank: int) -> np.ndarray:
    """_ctc_decoder.py:253-262: index (in the raw T+2 path) of the first entry of every non-blank run."""
    raw = list(raw)
    ts = []
    for i, idx in enumerate(raw):
        if idx == blank:
            continue
        if i == 0 or idx != raw[i - 1]:
            ts.append(i)
    return np.asarray(ts, dtype=np.int32)


class _Hyp:
    __slots__ = ("score", "lm", "parent", "token", "prev_blank")

    def __init__(self, score, lm, parent, token, prev_blank):
        self.score, self.lm, self.parent, self.token, self.prev_blank = score, lm, parent, token, prev_blank


def _store(cands, beam_size, threshold_score):
    """candidatesStore: threshold prune, merge hypotheses equal in (LM state, token, prevBlank) keeping the max
    (logAdd = false), keep the `beam_size` best."""
    merged = {}
    for c in cands:
        if c.score < threshold_score:
            continue
        k = (c.lm, c.token, c.prev_blank)
        o = merged.get(k)
        if o is None or c.score > o.score:
            merged[k] = c
    out = sorted(merged.values(), key=lambda h: -h.score)
    return out[:beam_size]


def flashlight_lexfree_decode(emissions: np.ndarray, blank: int, sil: int, beam_size: int = 10,
                              beam_threshold: float = 50.0, sil_score: float = 0.0) -> np.ndarray:
    """Raw token path (length T+2) of the best hypothesis.  emissions fp32 [T, N].  ZeroLM: the LM state is the
    token history (a trie node: `state->child(token)`), its score 0; represented here as a tuple."""
    T, N = emissions.shape
    hyps = [_Hyp(0.0, (), None, sil, False)]
    for t in range(T):
        cands, best = [], -np.inf
        for ph in hyps:
            for n in range(N):                                   # beam_size_token=None -> every token
                score = ph.score + float(emissions[t, n])
                if n == sil:
                    score += sil_score
                if n != blank and (n != ph.token or ph.prev_blank):
                    h = _Hyp(score, ph.lm + (n,), ph, n, False)
                elif n == blank:
                    h = _Hyp(score, ph.lm, ph, n, True)
                else:
                    h = _Hyp(score, ph.lm, ph, n, False)
                if score >= best - beam_threshold:               # candidatesAdd
                    cands.append(h)
                    best = max(best, score)
        hyps = _store(cands, beam_size, best - beam_threshold)
    cands = [_Hyp(ph.score, ph.lm, ph, sil, False) for ph in hyps]     # decodeEnd (ZeroLM.finish scores 0)
    hyps = _store(cands, beam_size, max(c.score for c in cands) - beam_threshold)
    h, raw = hyps[0], []
    while h is not None:
        raw.append(h.token)
        h = h.parent
    return np.asarray(raw[::-1], dtype=np.int64)


def greedy_raw_path(emissions: np.ndarray, sil: int) -> np.ndarray:
    """The same raw path without the search: [sil] + frame-wise argmax + [sil]."""
    return np.concatenate([[sil], emissions.argmax(-1), [sil]]).astype(np.int64)


def reference_decode(emissions: np.ndarray, blank: int = 0, sil: int = 1):
    """(tokens, timesteps) as `decoder(emissions)[0][0].tokens / .timesteps` of the reference's decoder."""
    raw = greedy_raw_path(emissions, sil)
    return torchaudio_tokens(raw, blank), torchaudio_timesteps(raw, blank)


class Hypothesis:
    """Stand-in for torchaudio's CTCHypothesis (`.tokens`, `.timesteps`) used when the reference's classes are run
    with this module as their decoder (tests/golden/make_golden_v2.py)."""

    def __init__(self, tokens, timesteps):
        import torch
        self.tokens = torch.as_tensor(np.asarray(tokens), dtype=torch.long)
        self.timesteps = torch.as_tensor(np.asarray(timesteps), dtype=torch.int32)
        self.words, self.score = [], 0.0


def ctc_decoder(lexicon=None, tokens=None, lm=None, nbest=1, beam_size=10, beam_size_token=None, beam_threshold=50,
                blank_token="-", sil_token="|", **_):
    """Signature-compatible replacement for `torchaudio.models.decoder.ctc_decoder` (lexicon-free, no LM)."""
    assert lexicon is None and lm is None and nbest == 1
    blank, sil = list(tokens).index(blank_token), list(tokens).index(sil_token)

    def decode(emissions, lengths=None):
        em = emissions.detach().cpu().numpy() if hasattr(emissions, "detach") else np.asarray(emissions)
        out = []
        for b in range(em.shape[0]):
            Tb = em.shape[1] if lengths is None else int(lengths[b])
            tk, ts = reference_decode(em[b, :Tb], blank, sil)
            out.append([Hypothesis(tk, ts)])
        return out

    return decode
