"""CPU suite: the oracle restatement against the golden fixtures produced by the reference's own classes
(tests/golden/make_golden.py), known-answer vectors of SURVEY.md §8c, and host-side logic."""
import numpy as np
import pytest
import torch

from helpers import VOCAB, cfg_base, cfg_large, golden, backbone_sd
from oracle import ctc as octc
from oracle import heads as oh
from oracle import w2v2 as ow
from oracle import weights as W
from aptai_b200.config import W2V2Config, frame_lengths


def test_frame_count_known_answers():
    cfg = W2V2Config.base()
    assert frame_lengths(cfg, [16000, 32000, 64000, 128000, 320000]) == [49, 99, 199, 399, 999]
    assert frame_lengths(cfg, [30000, 25000]) == [93, 77]
    assert ow.conv_out_length(128000, cfg) == 399


def test_lowpass_taps_known_answer():
    g = golden()
    taps = oh.lowpass_taps()
    assert taps.numel() == 51 and taps.dtype == torch.float64
    assert abs(float(taps.sum()) - 1.0) < 1e-15
    np.testing.assert_array_equal(taps.numpy(), g["g5_taps"])
    y = oh.lowpass(torch.from_numpy(g["g5_lp_in"]), taps)
    np.testing.assert_allclose(y.numpy(), g["g5_lp_out"], atol=1e-6)
    np.testing.assert_allclose(oh.positional_encoding(128, 60).numpy(), g["g5_pe"], atol=1e-7)


def test_forward_sum_oracle_matches_reference():
    g = golden()
    loss, _ = octc.forward_sum_loss(g["g5_fs_in"], g["g5_fs_text"], g["g5_fs_mel"], -1.0)
    assert abs(loss - float(g["g5_fs_loss"][0])) < 1e-4 * abs(loss)


def test_viterbi_oracle_matches_torchaudio_golden():
    g = golden()
    for i in range(int(g["g6_n"][0])):
        lp = torch.log_softmax(torch.from_numpy(g[f"g6_x{i}"]), -1).numpy()
        p, s = octc.viterbi_align(lp, g[f"g6_t{i}"], blank=0)
        assert np.array_equal(p, g[f"g6_p{i}"]), i
        assert np.array_equal(s, g[f"g6_s{i}"]), i
    # SURVEY.md §4 known answer
    p, _ = octc.viterbi_align(torch.log_softmax(torch.zeros(6, 3), -1).numpy(), [1, 2])
    assert p.tolist() == [1, 2, 2, 2, 2, 2]
    with pytest.raises(ValueError):
        octc.viterbi_align(np.zeros((3, 4), np.float32), [1, 1, 1])


def test_oracle_pr_base_matches_reference_golden():
    """Wav2Vec2_PR.forward of the reference (12x768 'group' backbone): logits, log-probs, CTC loss, d loss/d logits."""
    g = golden()
    cfg = cfg_base()
    sd = backbone_sd(cfg, 1)
    lens = [32000, 27000, 16000]
    wav = W.waveforms(3, 32000, lens, seed=3234)
    hs = ow.forward(sd, cfg, wav, lens)
    hw, hb = W.linear_params(104, 46, 768)
    logits = torch.nn.functional.linear(hs[-1], hw, hb).numpy()
    np.testing.assert_allclose(logits, g["g3_logits"], atol=2e-4, rtol=1e-3)
    np.testing.assert_allclose(hs[-1].numpy()[:, ::8, ::16], g["g3_hidden"], atol=2e-4, rtol=1e-3)
    labels = g["g3_labels"]
    il = np.asarray(frame_lengths(cfg, lens))
    tl = (labels >= 0).sum(-1)
    o = octc.ctc_loss_grad(g["g3_logits"], labels, il, tl, blank=0, zero_infinity=True, reduction="mean")
    assert abs(o["loss"] - float(g["g3_loss"][0])) < 1e-4 * abs(o["loss"])
    np.testing.assert_allclose(o["log_probs"], g["g3_log_probs"], atol=1e-5)
    np.testing.assert_allclose(o["grad"], g["g3_grad_logits"], atol=2e-6, rtol=1e-3)


@pytest.mark.timeout(600)
def test_oracle_aptai_large_matches_reference_golden():
    """APTAI.get_aptai_output / APTAI.forward of the reference on the 24x1024 'layer' backbone."""
    g = golden()
    cfg = cfg_large()
    sd = backbone_sd(cfg, 0)
    tvw, tvb = W.linear_params(101, 9, 1024)
    pw, pb = W.linear_params(102, 46, 1024)
    taps = oh.lowpass_taps()
    wav1 = W.waveforms(1, 32000, None, seed=1234)
    h = ow.forward(sd, cfg, wav1, [32000])[-1]
    _, tv, logits = oh.aptai_heads(h, tvw, tvb, pw, pb, taps)
    np.testing.assert_allclose(logits[0].numpy(), g["g1_logits"], atol=2e-4, rtol=1e-3)
    np.testing.assert_allclose(tv[0].numpy(), g["g1_tvs"], atol=1e-4)
    assert (logits[0].argmax(-1).numpy() == g["g1_pred"]).mean() == 1.0
    lens2 = [32000, 24000]
    wav2 = W.waveforms(2, 32000, lens2, seed=2234)
    h2 = ow.forward(sd, cfg, wav2, lens2)[-1]
    _, tv2, logits2 = oh.aptai_heads(h2, tvw, tvb, pw, pb, taps)
    np.testing.assert_allclose(tv2.numpy(), g["g2_tvs"], atol=1e-4)
    loss, mse, ce = oh.aptai_losses(tv2, logits2, torch.from_numpy(g["g2_phn"]), torch.from_numpy(g["g2_tvt"]))
    np.testing.assert_allclose([float(loss), float(mse), float(ce)], g["g2_losses"], rtol=1e-4)
    assert (logits2.argmax(-1).numpy() == g["g2_pred"]).mean() == 1.0


def test_ctc_oracle_vs_torch():
    rng = np.random.default_rng(0)
    B, T, V, S = 5, 40, 10, 8
    logits = torch.randn(B, T, V, dtype=torch.float64, requires_grad=True)
    tl = np.array([8, 5, 1, 7, 3])
    il = np.array([40, 30, 20, 9, 40])
    tg = np.full((B, S), -100, dtype=np.int32)
    for b in range(B):
        tg[b, : tl[b]] = rng.integers(1, V, size=tl[b])
    tg[3, :7] = 2                                   # needs 13 frames, has 9 -> infeasible -> zero_infinity
    lp = torch.log_softmax(logits, -1).transpose(0, 1)
    loss = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl),
                                        reduction="mean", zero_infinity=True, blank=0)
    loss.backward()
    o = octc.ctc_loss_grad(logits.detach().numpy(), tg, il, tl)
    assert abs(o["loss"] - loss.item()) < 1e-9
    np.testing.assert_allclose(o["grad"], logits.grad.numpy(), atol=1e-12)
    assert o["nll"][3] == 0.0


def test_force_tail_forward_sum_agrees_with_numpy_oracle_and_reference():
    """oracle/force_tail.py's differentiable forward-sum loss (the checker of the Force_APTAI training tests) against
    the NumPy restatement and against the reference's own ForwardSumLoss output stored in golden_v1 (G5)."""
    import torch
    from oracle.force_tail import forward_sum_loss
    g = golden()
    att = torch.from_numpy(np.asarray(g["g5_fs_in"], dtype=np.float32))
    text, mel = [int(x) for x in g["g5_fs_text"]], [int(x) for x in g["g5_fs_mel"]]
    loss_t = float(forward_sum_loss(att, text, mel, -1.0))
    loss_n, _ = octc.forward_sum_loss(g["g5_fs_in"], g["g5_fs_text"], g["g5_fs_mel"], -1.0)
    assert abs(loss_t - loss_n) < 1e-4 * abs(loss_n)
    assert abs(loss_t - float(g["g5_fs_loss"][0])) < 1e-4 * abs(float(g["g5_fs_loss"][0]))


def test_ctc_decode_restatement_beam_equals_greedy_wrap():
    """oracle/ctc_decode.py: the restated flashlight lexicon-free beam search (beam 10, threshold 50, ZeroLM,
    max-merge) returns the [sil] + frame-wise argmax + [sil] raw path, and torchaudio's token / time-stamp collapse of
    that path gives leading / trailing silence ids and frame + 1 time stamps."""
    from oracle import ctc_decode as D
    g = np.random.default_rng(3)
    for trial in range(40):
        T, N = int(g.integers(1, 30)), int(g.integers(3, 9))
        em = (g.standard_normal((T, N)) * (1.0 if trial % 2 else 0.05)).astype(np.float32)
        raw = D.flashlight_lexfree_decode(em, blank=0, sil=1)
        assert raw.shape == (T + 2,) and raw[0] == 1 and raw[-1] == 1
        assert np.array_equal(raw, D.greedy_raw_path(em, 1)), trial
    # known answers for the torchaudio layer (_ctc_decoder.py:248-262)
    raw = [1, 0, 0, 5, 5, 0, 7, 1, 1]          # sil | . . 5 5 . 7 sil | sil   (path ends in silence: merged)
    assert D.torchaudio_tokens(raw, 0).tolist() == [1, 5, 7, 1]
    assert D.torchaudio_timesteps(raw, 0).tolist() == [0, 3, 6, 7]
    raw = [1, 1, 4, 4, 0, 4, 1]                # path starts in silence: merged with the root entry
    assert D.torchaudio_tokens(raw, 0).tolist() == [1, 4, 4, 1]
    assert D.torchaudio_timesteps(raw, 0).tolist() == [0, 2, 5, 6]
    tk, ts = D.reference_decode(np.eye(4, dtype=np.float32)[[2, 2, 0, 3]], blank=0, sil=1)
    assert tk.tolist() == [1, 2, 3, 1] and ts.tolist() == [0, 1, 4, 5]
