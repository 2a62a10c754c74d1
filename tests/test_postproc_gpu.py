"""Device-side caller formats (aptai_b200.postproc, through the C ABI) against the reference's own outputs
(tests/golden/golden_post_v1.npz) and the oracle (oracle/postproc.py).  Integer / index / fp64-sequential results
are bit-exact; the fp32 resampler is within 5e-6 of the reference (fp32 accumulation of ~460 taps on |x| <= 0.45)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import ROOT
from aptai_b200 import postproc as pp
from oracle import postproc as op

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_post_v1.npz"))
NAMES = pp.TV_NAMES


def test_segments_bit_exact(cuda):
    frames = G["p_frames"].tolist()
    dur = pp.phn_frames2dur(frames)
    assert [d[0] for d in dur] == G["p_dur_start"].tolist()
    assert [d[1] for d in dur] == G["p_dur_end"].tolist()
    assert [d[2] for d in dur] == G["p_dur_phn"].tolist()
    assert pp.phn_frame_id2phn(frames) == G["p_id2phn"].tolist()
    assert pp.phn_frames2dur([]) == [] and pp.phn_frame_id2phn([]) == []
    # batched, ragged, at the maximum utterance length (999 frames)
    rng = np.random.Generator(np.random.PCG64(3))
    B, T = 64, 999
    fr = np.repeat(rng.integers(0, 46, size=(B, 400)), 3, axis=1)[:, :T].astype(np.int64)
    lens = rng.integers(1, T + 1, size=B).astype(np.int32)
    lens[0], lens[1] = T, 1
    st, en, ph, ns = pp.frames_to_segments_batch(torch.from_numpy(fr).to(cuda), torch.from_numpy(lens).to(cuda))
    st, en, ph, ns = st.cpu().numpy(), en.cpu().numpy(), ph.cpu().numpy(), ns.cpu().numpy()
    for b in range(B):
        ref = op.phn_frames2dur(fr[b, : lens[b]].tolist(), resolution=1)
        assert ns[b] == len(ref)
        assert st[b, : ns[b]].tolist() == [int(r[0]) for r in ref]
        assert en[b, : ns[b]].tolist() == [int(r[1]) for r in ref]
        assert ph[b, : ns[b]].tolist() == [r[2] for r in ref]


def test_tv_metrics(cuda):
    rm = pp.tvs_metric_rmse(G["m_gt"], G["m_pred"])
    pc = pp.tvs_metric_ppc(G["m_gt"], G["m_pred"])
    assert np.array_equal(np.asarray([rm[k] for k in NAMES]), G["m_rmse"])               # bit-exact
    np.testing.assert_allclose(np.asarray([pc[k].statistic for k in NAMES]), G["m_pcc"], rtol=1e-12)
    # scipy-style result object, as the reference's validate/test loops read it (train/train_aptai.py:582-584)
    from scipy.stats import pearsonr
    for i, k in enumerate(NAMES):
        r, pv = pc[k]
        ref = pearsonr(G["m_gt"][:, i].astype(np.float64), G["m_pred"][:, i].astype(np.float64))
        assert r == pc[k].statistic and pv == pc[k].pvalue
        np.testing.assert_allclose(pv, ref.pvalue, rtol=1e-6, atol=1e-300)
    # batched with lengths
    rng = np.random.Generator(np.random.PCG64(4))
    gt = rng.standard_normal((5, 300, 9)).astype(np.float32)
    pr = (gt + rng.standard_normal((5, 300, 9)) * 0.5).astype(np.float32)
    lens = np.asarray([300, 17, 2, 299, 128], dtype=np.int32)
    r, p = pp.tv_metrics_batch(torch.from_numpy(gt).to(cuda), torch.from_numpy(pr).to(cuda), torch.from_numpy(lens))
    for b in range(5):
        assert np.array_equal(r[b].cpu().numpy(), op.tvs_metric_rmse(gt[b, : lens[b]], pr[b, : lens[b]]))
        np.testing.assert_allclose(p[b].cpu().numpy(), op.tvs_metric_pcc(gt[b, : lens[b]], pr[b, : lens[b]]), rtol=1e-10)


def test_boundary_stats_and_overlap_bit_exact(cuda):
    got = np.asarray(pp.get_stats(G["b_y"], G["b_yhat"], tolerance=0.02), dtype=np.float64)
    assert np.array_equal(got, G["b_stats"])
    lens = G["o_lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    a = [G["o_a"][off[i]: off[i + 1]].tolist() for i in range(len(lens))]
    b = [G["o_b"][off[i]: off[i + 1]].tolist() for i in range(len(lens))]
    assert pp.evaluate_overlap(a, b) == float(G["o_overlap"][0])
    rng = np.random.Generator(np.random.PCG64(5))
    ys = [np.sort(np.round(rng.uniform(0, 20, size=n), 2)) for n in (1, 60, 200, 33)]
    hs = [np.sort(np.round(rng.uniform(0, 20, size=n), 2)) for n in (5, 61, 180, 1)]
    cnt = pp.boundary_counters_batch(ys, hs, 0.02)
    for i in range(4):
        assert tuple(cnt[i]) == op.boundary_counters(ys[i], hs[i], 0.02)


def test_interpolate_and_collate_bit_exact(cuda):
    assert np.array_equal(pp.interpolate_signal(G["i_sig"], 97), G["i_out_97"])
    assert np.array_equal(pp.interpolate_signal(G["i_sig"], 400), G["i_out_400"])
    for m in (50, 97, 211, 399):          # the reference's call pattern: one trajectory (1-D) at a time
        assert np.array_equal(pp.interpolate_signal(G["i_sig"][:, 0], m), op.interpolate_signal(G["i_sig"][:, 0], m))
    all9 = pp.interpolate_signal(G["i_sig"], 97, per_channel=True)
    for c in range(9):
        assert np.array_equal(all9[:, c], op.interpolate_signal(G["i_sig"][:, c], 97))
    batch = []
    for i in range(3):
        tv = G[f"c_tv{i}"]
        batch.append({"audio": torch.from_numpy(G[f"c_audio{i}"]), "audio_len": len(G[f"c_audio{i}"]),
                      "phn_frames_49hz": G[f"c_phn{i}"].tolist(),
                      "tvs_norm_49hz": {k: tv[:, j] for j, k in enumerate(NAMES)}})
    col = pp.collate_fn(batch)
    assert col["audio_inputs"].is_cuda and col["audio_inputs"].dtype == torch.float32
    assert np.array_equal(col["audio_inputs"].cpu().numpy(), G["c_audio_inputs"])
    assert np.array_equal(col["audio_lengths"].cpu().numpy(), G["c_audio_lengths"]) and col["audio_lengths"].dtype == torch.int64
    assert np.array_equal(col["phn_frames_49hz"].cpu().numpy(), G["c_phn_frames"]) and col["phn_frames_49hz"].dtype == torch.int64
    for j, k in enumerate(NAMES):
        assert np.array_equal(col[k].cpu().numpy(), G["c_tvs"][:, :, j])


def test_resample(cuda):
    for key, fs, n in (("r_44100", 44100, 22050), ("r_22050", 22050, 9999), ("r_8000", 8000, 4000)):
        y = pp.resample(torch.from_numpy(G["r_wav"][:n]), fs, 16000).cpu().numpy()
        assert y.shape == G[key].shape
        np.testing.assert_allclose(y, G[key], atol=5e-6, rtol=0)
        np.testing.assert_allclose(y, op.sinc_resample(G["r_wav"][:n], fs, 16000), atol=2e-6, rtol=0)
    # batched with per-row lengths: every row equals its own single-utterance result, zero beyond
    x = torch.from_numpy(np.stack([G["r_wav"][:20000], np.concatenate([G["r_wav"][:12345], np.zeros(7655, np.float32)])]))
    y, nl = pp.resample(x, 44100, 16000, lengths=torch.tensor([20000, 12345]))
    y1 = pp.resample(torch.from_numpy(G["r_wav"][:12345]), 44100, 16000)
    assert nl.tolist() == [7257, y1.numel()]
    assert torch.equal(y[1, : y1.numel()], y1) and float(y[1, y1.numel():].abs().max()) == 0.0
