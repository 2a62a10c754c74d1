"""Oracle (oracle/postproc.py) vs the reference's own caller-side functions (tests/golden/golden_post_v1.npz)."""
import os

import numpy as np

from helpers import ROOT
from oracle import postproc as op

G = np.load(os.path.join(ROOT, "tests", "golden", "golden_post_v1.npz"))


def test_segments():
    dur = op.phn_frames2dur(G["p_frames"].tolist())
    assert [d[0] for d in dur] == G["p_dur_start"].tolist()
    assert [d[1] for d in dur] == G["p_dur_end"].tolist()
    assert [d[2] for d in dur] == G["p_dur_phn"].tolist()
    assert op.phn_frame_id2phn(G["p_frames"].tolist()) == G["p_id2phn"].tolist()
    assert op.phn_frames2dur([]) == []


def test_tv_metrics():
    assert np.array_equal(op.tvs_metric_rmse(G["m_gt"], G["m_pred"]), G["m_rmse"])          # bit-exact
    np.testing.assert_allclose(op.tvs_metric_pcc(G["m_gt"], G["m_pred"]), G["m_pcc"], rtol=1e-13)


def test_boundary_stats_and_overlap():
    c = op.boundary_counters(G["b_y"], G["b_yhat"], 0.02)
    np.testing.assert_allclose(np.asarray(op.get_metrics(*c)), G["b_stats"], rtol=0, atol=0)
    lens = G["o_lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    a = [G["o_a"][off[i]: off[i + 1]] for i in range(len(lens))]
    b = [G["o_b"][off[i]: off[i + 1]] for i in range(len(lens))]
    assert op.evaluate_overlap(a, b) == float(G["o_overlap"][0])


def test_interpolate_signal():
    assert np.array_equal(op.interpolate_signal(G["i_sig"], 97), G["i_out_97"])
    assert np.array_equal(op.interpolate_signal(G["i_sig"], 400), G["i_out_400"])


def test_collate():
    aud = [G[f"c_audio{i}"] for i in range(3)]
    assert np.array_equal(op.pad_sequence(aud, 0.0, np.float32), G["c_audio_inputs"])
    assert np.array_equal(op.pad_sequence([G[f"c_phn{i}"] for i in range(3)], 0, np.int64), G["c_phn_frames"])
    for c in range(9):
        assert np.array_equal(op.pad_sequence([G[f"c_tv{i}"][:, c] for i in range(3)], -100.0, np.float32),
                              G["c_tvs"][:, :, c])


def test_resample():
    for key, fs, n in (("r_44100", 44100, 22050), ("r_22050", 22050, 9999), ("r_8000", 8000, 4000)):
        y = op.sinc_resample(G["r_wav"][:n], fs, 16000)
        assert y.shape == G[key].shape
        # the reference accumulates ~460 taps in fp32 (torch conv1d); the oracle in fp64: 5e-6 on |x| <= 0.45
        np.testing.assert_allclose(y, G[key], atol=5e-6, rtol=0)
