"""The oracle's BACKWARD is pinned too: torch-CPU autograd through oracle/w2v2.py + the reference's loss recipe
must reproduce the gradients the reference's own classes produced (tests/golden/golden_train_v1.npz)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from helpers import ROOT, backbone_sd, cfg_base
from oracle import w2v2 as ow
from oracle import weights as W

GOLD = os.path.join(ROOT, "tests", "golden", "golden_train_v1.npz")


def test_oracle_pr_gradients_match_reference():
    g = np.load(GOLD)
    cfg = cfg_base(vocab_size=46)
    sd = {k: v.clone() for k, v in backbone_sd(cfg, 1).items()}
    trainable = [k for k in sd if not k.startswith("feature_extractor.") and k != "masked_spec_embed"]
    for k in trainable:
        sd[k].requires_grad_(True)
    hw, hb = W.linear_params(104, 46, 768)
    hw.requires_grad_(True); hb.requires_grad_(True)
    lens3 = [32000, 27000, 16000]
    wav3 = W.waveforms(3, 32000, lens3, seed=3234)
    labels, _ = W.phoneme_sequences(3, 10, 40, 2, 45, seed=7, pad=-100)
    labels[2, 5:] = -100
    torch.set_num_threads(os.cpu_count())
    hidden, _, flen = ow.forward(sd, cfg, wav3, lens3, return_features=True)
    logits = F.linear(hidden[-1], hw, hb)
    # models/w2v2_pr.py:59-81
    lp = F.log_softmax(logits, dim=-1, dtype=torch.float32).transpose(0, 1)
    tl = (labels >= 0).sum(-1)
    flat = labels[labels >= 0]
    loss = F.ctc_loss(lp, flat, flen, tl, blank=0, reduction="mean", zero_infinity=True)
    assert abs(float(loss) - float(g["t3_loss"][0])) < 1e-4 * float(g["t3_loss"][0])
    loss.backward()
    ref = dict(zip([str(n) for n in g["t3_grad_names"]], g["t3_grad_norms"]))
    assert len(ref) == len(trainable) + 2
    floor = 1e-6 * max(ref.values())     # k_proj.bias gradients are analytically zero (softmax shift invariance)
    for k in trainable:
        ours = float(sd[k].grad.double().norm())
        assert abs(ours - ref["wav2vec2." + k]) <= 2e-4 * ref["wav2vec2." + k] + floor, k
    assert abs(float(hw.grad.double().norm()) - ref["pr_head.weight"]) <= 2e-4 * ref["pr_head.weight"]
    for key in g.files:
        if key.startswith("t3_grad::wav2vec2."):
            k = key[len("t3_grad::wav2vec2."):]
            np.testing.assert_allclose(sd[k].grad.reshape(-1)[:256].numpy(), g[key], rtol=2e-3, atol=1e-7)


def test_oracle_regulariser_placement_matches_transformers():
    """The oracle's training-mode forward (explicit dropout masks, LayerDrop set, SpecAugment mask) against the
    installed transformers' Wav2Vec2Model in train mode with its nn.Dropout modules replaced by the same masks —
    pins WHERE every regulariser acts (HF:434,570,573,603,647,694,766,701-706,1303) for both layer wirings."""
    import torch.nn as nn
    from transformers import Wav2Vec2Config, Wav2Vec2Model
    from aptai_b200.config import W2V2Config
    from aptai_b200.specaug import compute_mask_indices

    class MaskDrop(nn.Module):
        def __init__(self, mask):
            super().__init__()
            self.mask = mask

        def forward(self, x):
            return x * self.mask.view(x.shape)

    for stable in (False, True):
        kw = dict(vocab_size=46, num_hidden_layers=3, hidden_dropout=0.1, activation_dropout=0.1, feat_proj_dropout=0.1,
                  attention_dropout=0.1, attn_implementation="eager", final_dropout=0.0, layerdrop=0.4, apply_spec_augment=True, mask_time_prob=0.3,
                  mask_time_length=10, mask_time_min_masks=2, mask_feature_prob=0.2, mask_feature_length=8,
                  mask_feature_min_masks=1)
        if stable:
            kw.update(feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)
        hf_cfg = Wav2Vec2Config(**kw)
        cfg = W2V2Config.from_any(hf_cfg)
        sd = backbone_sd(cfg, 3)
        model = Wav2Vec2Model(hf_cfg)
        model.load_state_dict(sd, strict=True)
        model.train()
        lens = [16000, 12000]
        wav = W.waveforms(2, 16000, lens, seed=99)
        B, T, H, Fi = 2, 49, cfg.hidden_size, cfg.intermediate_size
        g = torch.Generator().manual_seed(1)
        mk = lambda shape, p: (torch.rand(shape, generator=g) >= p).float() / (1 - p)
        reg = {"proj": mk((B, T, H), 0.1), "enc": mk((B, T, H), 0.1)}
        model.feature_projection.dropout = MaskDrop(reg["proj"])
        model.encoder.dropout = MaskDrop(reg["enc"])
        for l, layer in enumerate(model.encoder.layers):
            reg[("attn", l)], reg[("act", l)], reg[("ffn", l)] = mk((B, T, H), 0.1), mk((B, T, Fi), 0.1), mk((B, T, H), 0.1)
            layer.dropout = MaskDrop(reg[("attn", l)])
            layer.feed_forward.intermediate_dropout = MaskDrop(reg[("act", l)])
            layer.feed_forward.output_dropout = MaskDrop(reg[("ffn", l)])
        am = torch.zeros((B, 16000), dtype=torch.long)
        for b, n in enumerate(lens):
            am[b, :n] = 1
        torch.manual_seed(11)
        reg["skip"] = {l for l in range(cfg.num_hidden_layers) if bool(torch.rand([]) < cfg.layerdrop)}
        assert 0 < len(reg["skip"]) < cfg.num_hidden_layers
        live = [l for l in range(cfg.num_hidden_layers) if l not in reg["skip"]]
        for l in live:
            reg[("attp", l)] = mk((B, cfg.num_attention_heads, T, T), 0.1)
        # dropout on the attention probabilities is a functional call (HF:461): patch it to the same masks
        orig_dropout, seen = F.dropout, []

        def fake_dropout(x, p=0.5, training=True, inplace=False):
            if x.dim() == 4 and training and p > 0:
                seen.append(1)
                return x * reg[("attp", live[len(seen) - 1])]
            return orig_dropout(x, p, training, inplace)

        torch.manual_seed(11)
        np.random.seed(7)
        torch.nn.functional.dropout = fake_dropout
        try:
            with torch.no_grad():
                ref = model(wav, attention_mask=am).last_hidden_state
        finally:
            torch.nn.functional.dropout = orig_dropout
        assert len(seen) == len(live)
        np.random.seed(7)
        flen = [ow.conv_out_length(n, cfg) for n in lens]
        reg["spec"] = torch.from_numpy(compute_mask_indices((B, T), 0.3, 10, frame_lens=flen, min_masks=2))
        reg["spec_feat"] = torch.from_numpy(compute_mask_indices((B, H), 0.2, 8, min_masks=1))
        assert reg["spec"].any() and reg["spec_feat"].any()
        with torch.no_grad():
            got = ow.forward(sd, cfg, wav, lens, reg=reg)[-1]
        torch.testing.assert_close(got, ref, atol=2e-4, rtol=1e-4)


def test_oracle_force_tail_gradients_match_reference():
    """oracle/force_tail.py (the checker of the Force_APTAI training tests) against the reference's own Force_APTAI
    class in train mode (tests/golden/golden_force_train_v1.npz: batch 1, dropouts at p = 0, 24x1024 recogniser)."""
    from helpers import VOCAB, cfg_large, force_tail_state
    from oracle.force_tail import ForceTail
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_force_train_v1.npz"))
    cfg = cfg_large(vocab_size=46)
    wav = W.waveforms(1, 32000, None, seed=5151)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        h = ow.forward(backbone_sd(cfg, 0), cfg, wav, [32000])[-1]
    ref = ForceTail(1024, len(VOCAB))
    ref.load_state_dict(force_tail_state(ref.state_dict()), strict=True)
    known = g["known"]
    ids = torch.zeros((1, 60), dtype=torch.int64)
    ids[0, : len(known)] = torch.from_numpy(known)
    out = ref(h, ids, [99], [len(known)], torch.from_numpy(g["tvt"]))
    out["loss"].backward()
    np.testing.assert_allclose([float(out["loss"]), float(out["tv_loss"]), float(out["align_loss"])], g["losses"], rtol=2e-5)
    norms = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    params = dict(ref.named_parameters())
    assert set(norms) == set(params)
    for k, n_ref in norms.items():
        got = params[k].grad
        assert abs(float(got.double().norm()) - n_ref) <= 2e-4 * n_ref + 1e-9, k
        sl = g[f"grad::{k}"]
        np.testing.assert_allclose(got.reshape(-1)[:256].numpy(), sl, atol=2e-4 * float(np.abs(sl).max()) + 1e-9, rtol=2e-3)
