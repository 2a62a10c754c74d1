"""NumPy restatement of the alignment stage — TEST INFRASTRUCTURE ONLY.

* `ctc_loss_grad`: log-softmax + CTC negative log-likelihood and its gradient w.r.t. the logits, float64,
  following ATen's LossCTC (the library behind F.ctc_loss / nn.CTCLoss at models/w2v2_pr.py:73-81 and
  models/modules.py:75,110): alpha/beta recursions in log space, zero_infinity, 'mean' = mean_b(nll_b /
  clamp_min(target_len_b, 1)).
* `forward_sum_loss`: models/modules.py:93-116.
* `viterbi_align`: CTC Viterbi forced alignment with the exact tie-breaking and fp32 arithmetic of
  `torchaudio.functional.forced_align` — bit-exact oracle for the CUDA kernel.
* `greedy_decode`: argmax -> merge repeats -> drop blank.
"""
from __future__ import annotations

import numpy as np

NEG = -np.inf


def _lse(*xs):
    m = np.max(np.stack(xs), axis=0)
    m0 = np.where(np.isneginf(m), 0.0, m)
    with np.errstate(divide="ignore"):
        return np.log(sum(np.exp(x - m0) for x in xs)) + m0


def log_softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    return x - m - np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True))


def ctc_single(lp, target, blank=0):
    """lp float64 [T, V] log-probs of ONE utterance (already cut to its length); target int [S].
    Returns (nll, dnll/dlogits [T, V])."""
    T, V = lp.shape
    S = len(target)
    NS = 2 * S + 1
    lab = np.full(NS, blank, dtype=np.int64)
    lab[1::2] = target
    skip = np.zeros(NS, dtype=bool)
    skip[3::2] = lab[3::2] != lab[1:-2:2]
    alpha = np.full((T, NS), NEG)
    beta = np.full((T, NS), NEG)
    if T == 0:
        return np.inf, np.zeros((T, V))
    alpha[0, 0] = lp[0, blank]
    if NS > 1:
        alpha[0, 1] = lp[0, lab[1]]
    for t in range(1, T):
        a0 = alpha[t - 1]
        a1 = np.concatenate(([NEG], a0[:-1]))
        a2 = np.concatenate(([NEG, NEG], a0[:-2]))
        a2 = np.where(skip, a2, NEG)
        alpha[t] = _lse(a0, a1, a2) + lp[t, lab]
    nll = -_lse(alpha[T - 1, NS - 1], alpha[T - 1, NS - 2] if NS > 1 else np.float64(NEG))
    beta[T - 1, NS - 1] = lp[T - 1, blank]
    if NS > 1:
        beta[T - 1, NS - 2] = lp[T - 1, lab[NS - 2]]
    skipb = np.zeros(NS, dtype=bool)
    skipb[1:-2:2] = lab[1:-2:2] != lab[3::2]
    for t in range(T - 2, -1, -1):
        b0 = beta[t + 1]
        b1 = np.concatenate((b0[1:], [NEG]))
        b2 = np.concatenate((b0[2:], [NEG, NEG]))
        b2 = np.where(skipb, b2, NEG)
        beta[t] = _lse(b0, b1, b2) + lp[t, lab]
    grad = np.exp(lp)
    if np.isfinite(nll):
        ab = alpha + beta
        for c in np.unique(lab):
            idx = np.nonzero(lab == c)[0]
            occ = _lse(*[ab[:, i] for i in idx])
            with np.errstate(over="ignore"):
                grad[:, c] -= np.exp(occ + nll - lp[:, c])
    return float(nll), grad


def ctc_loss_grad(logits, targets, input_len, target_len, blank=0, zero_infinity=True, reduction="mean"):
    """logits [B,T,V]; targets int [B,Smax]; returns dict(loss, nll [B], grad [B,T,V] of `loss`, log_probs [T,B,V])."""
    logits = np.asarray(logits, dtype=np.float64)
    B, T, V = logits.shape
    lp_all = log_softmax(logits)
    nll = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        Tb, Sb = int(input_len[b]), int(target_len[b])
        n, g = ctc_single(lp_all[b, :Tb], np.asarray(targets[b, :Sb], dtype=np.int64), blank)
        if not np.isfinite(n) and zero_infinity:
            n, g = 0.0, np.zeros_like(g)
        nll[b] = n
        grad[b, :Tb] = g
    if reduction == "mean":
        scale = 1.0 / (np.maximum(np.asarray(target_len, dtype=np.float64), 1.0) * B)
    elif reduction == "sum":
        scale = np.ones(B)
    else:
        raise ValueError(reduction)
    return {"loss": float(np.sum(nll * scale)), "nll": nll, "grad": grad * scale[:, None, None],
            "log_probs": np.transpose(lp_all, (1, 0, 2)), "scale": scale}


def forward_sum_loss(attn_logprob, text_lens, mel_lens, blank_logprob=-1.0):
    """models/modules.py:77-117.  attn_logprob [B,1,T,N]; returns (loss, nll per utterance)."""
    a = np.asarray(attn_logprob, dtype=np.float64)
    B = a.shape[0]
    total = 0.0
    nlls = []
    for b in range(B):
        N, Tb = int(text_lens[b]), int(mel_lens[b])
        cur = np.concatenate((np.full((Tb, 1), blank_logprob), a[b, 0, :Tb, :N]), axis=1)   # blank column first
        lp = log_softmax(cur)
        n, _ = ctc_single(lp, np.arange(1, N + 1), blank=0)
        if not np.isfinite(n):
            n = 0.0                                   # zero_infinity=True
        nlls.append(n)
        total += n / max(N, 1)                        # nn.CTCLoss reduction='mean' on a batch of one
    return total / B, np.asarray(nlls)


def viterbi_align(lp, target, blank=0):
    """SURVEY.md Appendix E.  lp float32 [T, C]; target int [L].  Returns (path int32 [T], scores float32 [T]).
    Raises ValueError when T < L + R (R = adjacent repeats), as torchaudio does."""
    lp = np.asarray(lp, dtype=np.float32)
    T, C = lp.shape
    L = len(target)
    S = 2 * L + 1
    R = sum(1 for i in range(1, L) if target[i] == target[i - 1])
    if T < L + R:
        raise ValueError("targets length is too long for CTC")
    lab = np.full(S, blank, dtype=np.int64)
    lab[1::2] = target
    skip = np.zeros(S, dtype=bool)
    for s in range(3, S, 2):
        skip[s] = target[s // 2] != target[s // 2 - 1]
    alpha = np.full(S, -np.inf, dtype=np.float32)
    start = 0 if T - (L + R) > 0 else 1
    end = 1 if S == 1 else 2
    alpha[start:end] = lp[0, lab[start:end]]
    bp = np.zeros((T, S), dtype=np.int8)
    ninf = np.float32(-np.inf)
    for t in range(1, T):
        x0 = alpha
        x1 = np.concatenate(([ninf], alpha[:-1])).astype(np.float32)
        x2 = np.concatenate(([ninf, ninf], alpha[:-2])).astype(np.float32)
        x2 = np.where(skip, x2, ninf)
        c2 = (x2 > x1) & (x2 > x0)
        c1 = ~c2 & (x1 > x0) & (x1 > x2)
        best = np.where(c2, x2, np.where(c1, x1, x0)).astype(np.float32)
        bp[t] = np.where(c2, 2, np.where(c1, 1, 0))
        with np.errstate(invalid="ignore"):
            alpha = np.where(np.isneginf(best), ninf, (best + lp[t, lab]).astype(np.float32)).astype(np.float32)
    s = 0 if S == 1 else (S - 1 if alpha[S - 1] > alpha[S - 2] else S - 2)
    path = np.zeros(T, dtype=np.int32)
    for t in range(T - 1, -1, -1):
        path[t] = lab[s]
        s -= int(bp[t, s])
    scores = lp[np.arange(T), path]
    return path, scores


def greedy_decode(logits, blank=0):
    """argmax (first maximum) -> collapse repeats -> drop blank.  logits [T,V] -> (tokens, frame index of each)."""
    am = np.argmax(np.asarray(logits), axis=-1)
    toks, frames = [], []
    prev = -1
    for t, a in enumerate(am):
        if a != prev and a != blank:
            toks.append(int(a))
            frames.append(t)
        prev = a
    return np.asarray(toks, dtype=np.int32), np.asarray(frames, dtype=np.int32)
