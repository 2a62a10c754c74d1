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
