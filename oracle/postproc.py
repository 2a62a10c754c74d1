"""NumPy restatement of the reference's caller-side data formats — TEST INFRASTRUCTURE
ONLY.  Pinned against outputs of the reference's own functions (tests/golden/golden_post_v1.npz, made by
tests/golden/make_golden_post.py)."""
from __future__ import annotations

import math
from itertools import groupby

import numpy as np


def phn_frames2dur(phns, resolution=0.02):
    """utility.py:539-558."""
    counter, out = 0, []
    for p, grp in groupby(phns):
        n = len(list(grp))
        out.append((round(counter * resolution, 2), round((counter + n) * resolution, 2), p))
        counter += n
    return out


def phn_frame_id2phn(frame_id_seq):
    """utility.py:561-566."""
    return [p for p, _ in groupby(frame_id_seq)]


def tvs_metric_rmse(tvs_gt, tvs_pred):
    """utility.py:393-418: per channel sqrt(sum(se) / len(se)) with a sequential Python sum of float64 values."""
    out = []
    for c in range(tvs_gt.shape[1]):
        se = np.square(np.subtract(tvs_gt[:, c].tolist(), tvs_pred[:, c].tolist()))
        out.append(math.sqrt(sum(se) / len(se)))
    return np.asarray(out)


def tvs_metric_pcc(tvs_gt, tvs_pred):
    """utility.py:422-444 (scipy.stats.pearsonr statistic): r = <xm/|xm|, ym/|ym|>."""
    out = []
    for c in range(tvs_gt.shape[1]):
        x = np.asarray(tvs_gt[:, c].tolist(), dtype=np.float64)
        y = np.asarray(tvs_pred[:, c].tolist(), dtype=np.float64)
        xm, ym = x - x.mean(), y - y.mean()
        out.append(float(np.clip(np.dot(xm / np.linalg.norm(xm), ym / np.linalg.norm(ym)), -1.0, 1.0)))
    return np.asarray(out)


def get_metrics(precision_counter, recall_counter, pred_counter, gt_counter):
    """utility.py:572-586."""
    EPS, eps = 1e-7, 1e-5
    precision = precision_counter / (pred_counter + eps)
    recall = recall_counter / (gt_counter + eps)
    f1 = 2 * (precision * recall) / (precision + recall + eps)
    os_ = recall / (precision + EPS) - 1
    r1 = np.sqrt((1 - recall) ** 2 + os_ ** 2)
    r2 = (-os_ + recall - 1) / (np.sqrt(2))
    return precision, recall, f1, 1 - (np.abs(r1) + np.abs(r2)) / 2


def boundary_counters(y, yhat, tolerance=0.02):
    """utility.py:589-603: (precision_counter, recall_counter, len(yhat), len(y))."""
    y, yhat = np.asarray(y), np.asarray(yhat)
    pc = sum(int(np.abs(y - h).min() <= tolerance) for h in yhat)
    rc = sum(int(np.abs(yhat - v).min() <= tolerance) for v in y)
    return pc, rc, len(yhat), len(y)


def evaluate_overlap(gt_f, p_f):
    """utility.py:614-622."""
    hits = sum(int((np.asarray(a) == np.asarray(b)).sum()) for a, b in zip(gt_f, p_f))
    return hits / sum(len(a) for a in gt_f)


def interpolate_signal(org_sig, tar_len):
    """data/dataset_hprc.py:2307-2313 = scipy interp1d(kind='linear', axis=0) on arange(n) evaluated at
    linspace(0, n-1, tar_len).  The reference calls it per trajectory (1-D, dataset_hprc.py:2370,2411), where scipy
    dispatches to numpy.interp: slope * (x - x_lo) + y_lo with exact hits returned as is; for N-D input scipy's
    `_call_linear` computes (x_new-x_lo)/(x_hi-x_lo) * y_hi + (x_hi-x_new)/(x_hi-x_lo) * y_lo."""
    sig = np.asarray(org_sig, dtype=np.float64)
    n = sig.shape[0]
    x_new = np.linspace(0, n - 1, tar_len)
    if sig.ndim == 1:
        out = np.empty(tar_len)
        for i, x in enumerate(x_new):
            j = min(int(np.floor(x)), n - 2)
            if x == n - 1:
                out[i] = sig[n - 1]
            elif x == j:
                out[i] = sig[j]
            else:
                out[i] = (sig[j + 1] - sig[j]) / 1.0 * (x - j) + sig[j]
        return out
    hi = np.clip(np.searchsorted(np.arange(n), x_new), 1, n - 1)
    lo = hi - 1
    d = (hi - lo).astype(np.float64)
    bshape = (tar_len,) + (1,) * (sig.ndim - 1)
    return ((x_new - lo) / d).reshape(bshape) * sig[hi] + ((hi - x_new) / d).reshape(bshape) * sig[lo]


def pad_sequence(seqs, pad_value, dtype):
    """torch.nn.utils.rnn.pad_sequence(batch_first=True) as used by train/train_aptai.py:268-332."""
    L = max(len(s) for s in seqs)
    out = np.full((len(seqs), L), pad_value, dtype=dtype)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = np.asarray(s, dtype=dtype)
    return out


def sinc_resample(x, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio.functional.resample (sinc_interp_hann; torchaudio `_get_sinc_resample_kernel` +
    `_apply_sinc_resample_kernel`, the call data/dataset_hprc.py:68-72 makes), float64 accumulation."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float32)[:, None] / np.float32(new)).astype(np.float64) + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = (k * (window * (base / orig))).astype(np.float32).astype(np.float64)
    L = len(x)
    xp = np.concatenate([np.zeros(width), np.asarray(x, dtype=np.float64), np.zeros(width + orig)])
    nblk = (len(xp) - k.shape[1]) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, k.shape[1])[::orig][:nblk]      # [nblk, klen]
    y = (win @ k.T).reshape(-1)
    return y[: int(math.ceil(new * L / orig))].astype(np.float32)
