"""Device-side counterparts of what the reference's callers do right before and right after the hot path
(SURVEY.md §8f rows 3 and 4), with the reference's function names and return formats:

  input side   `collate_fn` (train/train_aptai.py:268-332 `_collate_fn`), `resample` (torchaudio.functional.resample
               as called by data/dataset_hprc.py:68-72), `interpolate_signal` (data/dataset_hprc.py:2307-2313)
  output side  `phn_frames2dur`, `phn_frame_id2phn` (utility.py:539-566), `tvs_metric_rmse`, `tvs_metric_ppc`
               (utility.py:393-444), `get_stats` / `get_metrics` (utility.py:572-611), `evaluate_overlap`
               (utility.py:614-622)

The single-utterance functions keep the reference's signatures; the `*_batch` variants take padded device tensors
+ lengths and do the whole batch in one launch.  No CPU fallback: inputs are moved to the CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from .lib import check

TV_NAMES = ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")
F32, F64, I32, I64 = torch.float32, torch.float64, torch.int32, torch.int64


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("aptai_b200.postproc: no CUDA device (there is no CPU path)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ================================================================================================ input side
def pad_ragged(seqs: Sequence, dtype, pad_value, device=None) -> torch.Tensor:
    """pad_sequence(batch_first=True, padding_value=pad_value) in one launch: the sequences are concatenated on the
    host, copied once, and scattered into the padded [B, Lmax] tensor on the device."""
    dev = _dev(device)
    ts = [torch.as_tensor(np.asarray(s) if not torch.is_tensor(s) else s).to(dtype).reshape(-1) for s in seqs]
    lens = [int(t.numel()) for t in ts]
    B, Lmax = len(ts), max(1, max(lens))
    offs = torch.tensor([0] + list(np.cumsum(lens)), dtype=I64)
    flat = (torch.cat(ts) if sum(lens) else torch.zeros(1, dtype=dtype)).to(dev, non_blocking=True)
    out = torch.empty((B, Lmax), dtype=dtype, device=dev)
    nbytes = flat.element_size()
    pad = (np.asarray([pad_value], dtype={F32: np.float32, I64: np.int64, F64: np.float64, I32: np.int32}[dtype])).tobytes()
    offs_d = offs.to(dev)
    check(_lib.load().aptai_collate_pad(flat.data_ptr(), nbytes, offs_d.data_ptr(), B, Lmax,
                                        C.c_char_p(pad), out.data_ptr(), _stream()), "collate_pad")
    return out


def collate_fn(batch: List[dict], device=None) -> Dict[str, torch.Tensor]:
    """train/train_aptai.py:268-332 `_collate_fn` (49 Hz normalised-TV variant): same keys, dtypes and padding values
    (audio 0.0, phoneme frames 0, TV targets -100.0), built on the device."""
    dev = _dev(device)
    out = {
        "audio_inputs": pad_ragged([x["audio"] for x in batch], F32, 0.0, dev),
        "audio_lengths": torch.tensor([x["audio_len"] for x in batch], dtype=I64, device=dev),
        "phn_frames_49hz": pad_ragged([x["phn_frames_49hz"] for x in batch], I64, 0, dev),
    }
    for k in TV_NAMES:
        src = [x["tvs_norm_49hz"][k] for x in batch]
        dt = F64 if np.asarray(src[0]).dtype == np.float64 else F32
        out[k] = pad_ragged(src, dt, -100.0, dev)
    return out


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio `_get_sinc_resample_kernel` (sinc_interp_hann, dtype=None: float64 construction, float32 result).
    Returns (kernel fp32 [new][2*width+orig], width, orig/gcd, new/gcd)."""
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
    k = k * (window * (base / orig))
    return torch.from_numpy(k.astype(np.float32)), width, orig, new


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int, lengths: Optional[torch.Tensor] = None):
    """torchaudio.functional.resample(waveform, orig_freq, new_freq) for [L] or [B, L] fp32 input.  With `lengths`
    (valid samples per row) every row is resampled for its own length and zero-filled beyond; returns the tensor
    (and the new lengths when `lengths` was given)."""
    dev = _dev(waveform.device if waveform.is_cuda else None)
    if int(orig_freq) == int(new_freq):
        return waveform if lengths is None else (waveform, lengths)
    x = waveform.to(device=dev, dtype=F32).contiguous()
    single = x.dim() == 1
    x2 = x.view(1, -1) if single else x
    B, L = x2.shape
    kern, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    ln = (torch.full((B,), L, dtype=I64) if lengths is None else lengths.to(I64).cpu()).to(dev)
    out_ld = int(math.ceil(new * L / orig))
    y = torch.empty((B, out_ld), dtype=F32, device=dev)
    kern_d = kern.to(dev)
    check(_lib.load().aptai_resample_fir(x2.data_ptr(), ln.data_ptr(), B, L, kern_d.data_ptr(), orig, new, width,
                                         y.data_ptr(), out_ld, _stream()), "resample_fir")
    y = y[0] if single else y
    if lengths is None:
        return y
    return y, torch.ceil(new * ln.double() / orig).long()


def interpolate_signal(org_sig, tar_len: int, per_channel: Optional[bool] = None):
    """data/dataset_hprc.py:2307-2313 (scipy interp1d, linear, axis 0) in fp64 on the device, bit-exact; returns a
    numpy array like the reference.  1-D input follows scipy's 1-D path (numpy.interp), which is how the reference
    calls it (one trajectory at a time, dataset_hprc.py:2370); `per_channel=True` applies that path to every column
    of a [n, C] array in one launch (all nine trajectories at once)."""
    dev = _dev()
    a = np.asarray(org_sig, dtype=np.float64)
    if per_channel is None:
        per_channel = a.ndim == 1
    shp = a.shape
    s = torch.from_numpy(a.reshape(shp[0], -1)).to(dev).contiguous()
    n, Cc = s.shape
    out = torch.empty((int(tar_len), Cc), dtype=F64, device=dev)
    check(_lib.load().aptai_interp_linear_f64(s.data_ptr(), n, Cc, int(tar_len), int(bool(per_channel)),
                                              out.data_ptr(), _stream()), "interp_linear")
    return out.cpu().numpy().reshape((int(tar_len),) + shp[1:])


# ================================================================================================ output side
def frames_to_segments_batch(frames: torch.Tensor, lens: torch.Tensor, max_seg: Optional[int] = None):
    """frames int64 [B,T] (device), lens int32 [B] -> (start int32 [B,S], end int32 [B,S], phn int64 [B,S], nseg [B])."""
    dev = _dev(frames.device if frames.is_cuda else None)
    f = frames.to(device=dev, dtype=I64).contiguous()
    B, T = f.shape
    ln = lens.to(device=dev, dtype=I32).contiguous()
    S = max_seg or T
    st = torch.zeros((B, S), dtype=I32, device=dev)
    en = torch.zeros((B, S), dtype=I32, device=dev)
    ph = torch.zeros((B, S), dtype=I64, device=dev)
    ns = torch.empty((B,), dtype=I32, device=dev)
    check(_lib.load().aptai_frames_to_segments(f.data_ptr(), ln.data_ptr(), B, T, st.data_ptr(), en.data_ptr(),
                                               ph.data_ptr(), ns.data_ptr(), S, _stream()), "frames_to_segments")
    return st, en, ph, ns


def phn_frames2dur(phns, resolution=0.02):
    """utility.py:539-558: list of (start s, end s, phoneme) with the reference's `round(..., 2)`."""
    n = len(phns)
    if n == 0:
        return []
    st, en, ph, ns = frames_to_segments_batch(torch.as_tensor(np.asarray(phns, dtype=np.int64)).view(1, -1),
                                              torch.tensor([n], dtype=I32))
    k = int(ns[0])
    st, en, ph = st[0, :k].tolist(), en[0, :k].tolist(), ph[0, :k].tolist()
    kind = type(phns[0]) if not isinstance(phns[0], (np.integer,)) else int
    return [(round(a * resolution, 2), round(b * resolution, 2), kind(p)) for a, b, p in zip(st, en, ph)]


def phn_frame_id2phn(frame_id_seq):
    """utility.py:561-566."""
    n = len(frame_id_seq)
    if n == 0:
        return []
    _, _, ph, ns = frames_to_segments_batch(torch.as_tensor(np.asarray(frame_id_seq, dtype=np.int64)).view(1, -1),
                                            torch.tensor([n], dtype=I32))
    return ph[0, : int(ns[0])].tolist()


def tv_metrics_batch(tvs_gt: torch.Tensor, tvs_pred: torch.Tensor, lens: torch.Tensor):
    """gt, pred fp32 [B,T,C] (device), lens [B] -> (rmse fp64 [B,C], pcc fp64 [B,C])."""
    dev = _dev(tvs_pred.device if tvs_pred.is_cuda else None)
    g = tvs_gt.to(device=dev, dtype=F32).contiguous()
    p = tvs_pred.to(device=dev, dtype=F32).contiguous()
    B, T, Cc = g.shape
    ln = lens.to(device=dev, dtype=I32).contiguous()
    rm = torch.empty((B, Cc), dtype=F64, device=dev)
    pc = torch.empty((B, Cc), dtype=F64, device=dev)
    check(_lib.load().aptai_tv_metrics(g.data_ptr(), p.data_ptr(), ln.data_ptr(), B, T, Cc, rm.data_ptr(),
                                       pc.data_ptr(), _stream()), "tv_metrics")
    return rm, pc


def tvs_metric_rmse(tvs_gt, tvs_pred):
    """utility.py:393-418: dict channel -> RMSE for one utterance ([T,9] arrays)."""
    g = torch.as_tensor(np.asarray(tvs_gt, dtype=np.float32))[None]
    p = torch.as_tensor(np.asarray(tvs_pred, dtype=np.float32))[None]
    rm, _ = tv_metrics_batch(g, p, torch.tensor([g.shape[1]], dtype=I32))
    return {k: float(v) for k, v in zip(TV_NAMES, rm[0].tolist())}


class PearsonRResult(tuple):
    """What `scipy.stats.pearsonr` returns, as far as the reference's callers use it: `.statistic`, `.pvalue`
    (train/train_aptai.py:582-584,767-768) and tuple unpacking `(r, p)`."""

    def __new__(cls, statistic: float, pvalue: float):
        return super().__new__(cls, (statistic, pvalue))

    statistic = property(lambda self: self[0])
    pvalue = property(lambda self: self[1])
    correlation = statistic


def _pearson_pvalue(r: float, n: int) -> float:
    """Two-sided p-value of scipy.stats.pearsonr: r ~ Beta(n/2-1, n/2-1) on [-1, 1] under the null hypothesis
    (host arithmetic on one scalar per channel; the correlation itself comes from the device kernel)."""
    if n < 3 or not math.isfinite(r):
        return 1.0 if n == 2 else float("nan")
    from scipy.special import betainc
    a = n / 2.0 - 1.0
    return float(min(1.0, 2.0 * betainc(a, a, (1.0 - min(1.0, abs(r))) / 2.0)))


def tvs_metric_ppc(tvs_gt, tvs_pred):
    """utility.py:422-444: dict channel -> scipy-style `pearsonr` result (`.statistic`, `.pvalue`, unpacks as a
    tuple).  r is computed by the device kernel, the p-value on the host from (r, n)."""
    g = torch.as_tensor(np.asarray(tvs_gt, dtype=np.float32))[None]
    p = torch.as_tensor(np.asarray(tvs_pred, dtype=np.float32))[None]
    n = int(g.shape[1])
    _, pc = tv_metrics_batch(g, p, torch.tensor([n], dtype=I32))
    return {k: PearsonRResult(float(v), _pearson_pvalue(float(v), n)) for k, v in zip(TV_NAMES, pc[0].tolist())}


def get_metrics(precision_counter, recall_counter, pred_counter, gt_counter):
    """utility.py:572-586 (host arithmetic on four counters)."""
    EPS, eps = 1e-7, 1e-5
    precision = precision_counter / (pred_counter + eps)
    recall = recall_counter / (gt_counter + eps)
    f1 = 2 * (precision * recall) / (precision + recall + eps)
    os_ = recall / (precision + EPS) - 1
    r1 = np.sqrt((1 - recall) ** 2 + os_ ** 2)
    r2 = (-os_ + recall - 1) / (np.sqrt(2))
    rval = 1 - (np.abs(r1) + np.abs(r2)) / 2
    return precision, recall, f1, rval


def boundary_counters_batch(ys: Sequence[np.ndarray], yhats: Sequence[np.ndarray], tolerance=0.02) -> np.ndarray:
    """int32 [B,4] = {precision_counter, recall_counter, len(yhat), len(y)} per utterance, one launch."""
    dev = _dev()
    B = len(ys)
    maxn = max(1, max(max(len(a) for a in ys), max(len(a) for a in yhats)))
    Y = np.zeros((B, maxn), dtype=np.float64)
    Hh = np.zeros((B, maxn), dtype=np.float64)
    for b in range(B):
        Y[b, : len(ys[b])] = np.asarray(ys[b], dtype=np.float64)
        Hh[b, : len(yhats[b])] = np.asarray(yhats[b], dtype=np.float64)
    ny = torch.tensor([len(a) for a in ys], dtype=I32, device=dev)
    nh = torch.tensor([len(a) for a in yhats], dtype=I32, device=dev)
    cnt = torch.empty((B, 4), dtype=I32, device=dev)
    Yd, Hd = torch.from_numpy(Y).to(dev), torch.from_numpy(Hh).to(dev)
    check(_lib.load().aptai_boundary_stats(Yd.data_ptr(), ny.data_ptr(), Hd.data_ptr(), nh.data_ptr(), B, maxn,
                                           float(tolerance), cnt.data_ptr(), _stream()), "boundary_stats")
    return cnt.cpu().numpy()


def get_stats(y, yhat, tolerance=0.02):
    """utility.py:589-611: boundary precision, recall, F1, R-value of one utterance."""
    c = boundary_counters_batch([np.asarray(y)], [np.asarray(yhat)], tolerance)[0]
    return get_metrics(int(c[0]), int(c[1]), int(c[2]), int(c[3]))


def evaluate_overlap(gt_f, p_f):
    """utility.py:614-622: fraction of frames whose labels agree, over a list of utterances."""
    dev = _dev()
    for a, b in zip(gt_f, p_f):
        assert len(a) == len(b)
    A = pad_ragged([np.asarray(a, dtype=np.int64) for a in gt_f], I64, -1, dev)
    Bm = pad_ragged([np.asarray(a, dtype=np.int64) for a in p_f], I64, -2, dev)
    lens = torch.tensor([len(a) for a in gt_f], dtype=I32, device=dev)
    hc = torch.empty((2,), dtype=torch.int64, device=dev)
    check(_lib.load().aptai_frame_overlap(A.data_ptr(), Bm.data_ptr(), lens.data_ptr(), A.shape[0], A.shape[1],
                                          hc.data_ptr(), _stream()), "frame_overlap")
    h, c = hc.tolist()
    return h / c
