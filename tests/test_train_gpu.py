"""Training-step parity of the drop-in modules against gradients produced by the REFERENCE's own classes
(tests/golden/golden_train_v1.npz, made by tests/golden/make_golden_train.py: torch CPU fp32 autograd through
models/aptai.py / models/w2v2_pr.py + transformers).  Our path: bf16 tensor-core forward and backward kernels
through the C ABI, fp32 master weights and gradient accumulation.

Tolerances (bf16 operands, fp32 accumulation; gradients of a 24-layer pre-LN stack on random weights):
  * loss: north-star tolerance (CTC 1e-3 relative; APTAI losses 2e-3 relative)
  * every parameter's gradient L2 norm within 5 % of the reference's
  * stored gradient slices: cosine similarity >= 0.98
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import ROOT, VOCAB, backbone_sd, cfg_base, cfg_large
from aptai_b200 import APTAI, Wav2Vec2_PR
from aptai_b200.backbone import register_in_memory_checkpoint
from aptai_b200.train import FusedAdam
from oracle import weights as W

GOLD = os.path.join(ROOT, "tests", "golden", "golden_train_v1.npz")


def _check_grads(model, g, tag, norm_tol=0.05, cos_min=0.98):
    names = [str(n) for n in g[f"{tag}_grad_names"]]
    norms = g[f"{tag}_grad_norms"]
    params = dict(model.named_parameters())
    worst = ("", 0.0)
    floor = 1e-4 * float(np.median(norms))   # k_proj.bias gradients are analytically zero (softmax shift invariance)
    for n, ref in zip(names, norms):
        p = params[n]
        assert p.grad is not None, n
        ours = float(p.grad.double().norm())
        if ref < floor:          # analytically zero: only bf16 rounding noise of the column sums is left
            assert ours < 100 * floor, (n, ours)
            continue
        rel = abs(ours - ref) / ref
        if rel > worst[1]:
            worst = (n, rel)
    assert worst[1] < norm_tol, f"gradient norm of {worst[0]} off by {worst[1]:.3%}"
    # parameters the reference leaves without gradient (frozen conv encoder, masked_spec_embed) stay zero / absent
    for n, p in params.items():
        if n not in names and p.grad is not None:
            assert float(p.grad.abs().max()) == 0.0, n
    low = ("", 1.0)
    for key in g.files:
        if not key.startswith(f"{tag}_grad::"):
            continue
        n = key.split("::", 1)[1]
        ref = torch.from_numpy(g[key]).double()
        ours = params[n].grad.reshape(-1)[: ref.numel()].double().cpu()
        cos = float((ours @ ref) / (ours.norm() * ref.norm()).clamp_min(1e-30))
        if cos < low[1]:
            low = (n, cos)
    assert low[1] > cos_min, f"gradient slice of {low[0]}: cosine {low[1]:.4f}"
    return worst, low


def _g2_inputs():
    lens2 = [32000, 24000]
    wav2 = W.waveforms(2, 32000, lens2, seed=2234)
    T = 99
    rng = np.random.Generator(np.random.PCG64(21))
    flen = [99, 74]
    phn = np.zeros((2, T), dtype=np.int64)
    tvt = np.full((2, T, 9), -100.0, dtype=np.float32)
    for b in range(2):
        phn[b, : flen[b]] = rng.integers(1, 46, size=flen[b])
        tvt[b, : flen[b]] = rng.standard_normal((flen[b], 9), dtype=np.float32)
    return wav2, torch.tensor(lens2), torch.from_numpy(phn), [torch.from_numpy(tvt[:, :, i]) for i in range(9)]


def test_aptai_training_step_vs_reference(cuda):
    g = np.load(GOLD)
    cfg = cfg_large()
    name = register_in_memory_checkpoint("mem://large-seed0-train", backbone_sd(cfg, 0))
    m = APTAI(cuda, VOCAB, name, cfg, None, phn_drop=0.0, tv_drop=0.0)
    tvw, tvb = W.linear_params(101, 9, 1024)
    pw, pb = W.linear_params(102, 46, 1024)
    with torch.no_grad():
        m.tv_head[2].weight.copy_(tvw); m.tv_head[2].bias.copy_(tvb)
        m.phn_head[2].weight.copy_(pw); m.phn_head[2].bias.copy_(pb)
    m = m.to(cuda).train()
    wav, lens, phn, tvs = _g2_inputs()
    args = (0, wav.to(cuda), lens.to(cuda), phn.to(cuda), *[t.to(cuda) for t in tvs])
    opt = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    opt.zero_grad()
    out = m(*args)
    losses = np.asarray([float(out["loss"].detach()), float(out["mse_loss"]), float(out["ce_loss"])])
    np.testing.assert_allclose(losses, g["t2_losses"], rtol=2e-3)
    assert out["loss"].requires_grad
    out["loss"].backward()
    worst, low = _check_grads(m, g, "t2")
    print("APTAI large: worst grad-norm deviation", worst, "lowest slice cosine", low)
    # gradient accumulation: a second backward doubles the buffer
    n0 = float(m.grad_buffer().flat.double().norm())
    m(*args)["loss"].backward()
    assert abs(float(m.grad_buffer().flat.double().norm()) / n0 - 2.0) < 1e-3
    # one optimizer step (reference: torch.optim.Adam(lr=1e-4)) from the single-batch gradient
    opt.zero_grad()
    m(*args)["loss"].backward()
    opt.step()
    with torch.no_grad():
        out2 = m(*args)
    after = np.asarray([float(out2["loss"]), float(out2["mse_loss"]), float(out2["ce_loss"])])
    print("after one Adam step:", after, "reference:", g["t2_losses_after_step"])
    # Adam's first step moves every weight by ~lr * sign(g): sign flips of near-zero gradients make this the
    # loosest check of the file
    np.testing.assert_allclose(after, g["t2_losses_after_step"], rtol=0.15)


def test_pr_training_step_vs_reference(cuda):
    g = np.load(GOLD)
    cfg = cfg_base(vocab_size=46)
    name = register_in_memory_checkpoint("mem://base-seed1-train", backbone_sd(cfg, 1))
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(104, 46, 768)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.wav2vec2.freeze_feature_encoder()
    pr = pr.to(cuda).train()
    lens3 = [32000, 27000, 16000]
    wav3 = W.waveforms(3, 32000, lens3, seed=3234)
    labels, _ = W.phoneme_sequences(3, 10, 40, 2, 45, seed=7, pad=-100)
    labels[2, 5:] = -100
    r = pr(wav3.to(cuda), torch.tensor(lens3, device=cuda), labels.to(cuda))
    assert abs(float(r["loss"].detach()) - float(g["t3_loss"][0])) / float(g["t3_loss"][0]) < 1e-3
    r["loss"].backward()
    worst, low = _check_grads(pr, g, "t3")
    print("PR base: worst grad-norm deviation", worst, "lowest slice cosine", low)


def test_dropout_kernel(cuda):
    from aptai_b200 import ops
    n = 1 << 22
    x = torch.randn(n, device=cuda)
    res = torch.randn(n, device=cuda)
    for p in (0.1, 0.5):
        y, yb = ops.dropout(x, p, 1234, want_f32=True, want_bf16=True)
        keep = y != 0
        assert abs(float(keep.float().mean()) - (1 - p)) < 3e-3
        torch.testing.assert_close(y[keep], x[keep] / (1 - p), rtol=1e-6, atol=0)
        assert torch.equal(yb, y.bfloat16())
        y2, _ = ops.dropout(x, p, 1234, want_f32=True)                       # deterministic in (seed, index)
        assert torch.equal(y, y2)
        y3, _ = ops.dropout(x, p, 1235, want_f32=True)
        assert 0.05 < float(((y3 != 0) != keep).float().mean()) < 2 * p      # another seed: another mask
        yr, _ = ops.dropout(x, p, 1234, residual=res, want_f32=True)
        torch.testing.assert_close(yr, y + res, rtol=1e-6, atol=1e-6)
        xb = x.bfloat16()
        _, zb = ops.dropout(xb, p, 1234, want_bf16=True)                     # bf16 input: same mask
        assert torch.equal((zb != 0) | (xb == 0), keep | (xb == 0))
    y0, _ = ops.dropout(x, 0.0, 7, residual=res, want_f32=True)               # p = 0: a plain add
    torch.testing.assert_close(y0, x + res)


def test_training_step_with_regularisers_replayed_by_the_oracle(cuda):
    """Dropout (feat_proj / hidden / activation / final), LayerDrop and SpecAugment (time and feature axis) on: the masks the step used are
    materialised from the same counter-based generator and replayed through the oracle (whose regulariser placement
    is pinned against transformers in tests/test_oracle_train_cpu.py); loss and every gradient must agree."""
    import torch.nn.functional as F
    from aptai_b200 import ops
    from oracle import w2v2 as ow
    cfg = cfg_base(vocab_size=46, hidden_dropout=0.1, activation_dropout=0.1, feat_proj_dropout=0.1, final_dropout=0.1,
                   attention_dropout=0.1, layerdrop=0.25, apply_spec_augment=True, mask_time_prob=0.2,
                   mask_time_length=10, mask_time_min_masks=2, mask_feature_prob=0.1, mask_feature_length=16,
                   mask_feature_min_masks=1)
    sd0 = backbone_sd(cfg_base(vocab_size=46), 1)
    name = register_in_memory_checkpoint("mem://base-seed1-reg", sd0)
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(104, 46, 768)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.wav2vec2.freeze_feature_encoder()
    pr = pr.to(cuda).train()
    lens3 = [32000, 27000, 16000]
    wav3 = W.waveforms(3, 32000, lens3, seed=3234)
    labels, _ = W.phoneme_sequences(3, 10, 40, 2, 45, seed=7, pad=-100)
    labels[2, 5:] = -100
    torch.manual_seed(1234)
    np.random.seed(5)
    r = pr(wav3.to(cuda), torch.tensor(lens3, device=cuda), labels.to(cuda))
    r["loss"].backward()
    w2v = pr.wav2vec2
    info = w2v._last_regularisers
    step, B, T, H, Fi = info["step"], 3, 99, 768, 3072
    assert 0 < len(info["skipped"]) < 12 and info["spec_rows"] is not None

    def mask(layer, site, shape, p):
        return ops.dropout(torch.ones(shape, device=cuda), p, w2v.drop_seed(step, layer, site), want_f32=True)[0].cpu()

    reg = {"proj": mask(-1, w2v.SITE_PROJ, (B, T, H), 0.1), "enc": mask(-1, w2v.SITE_ENC, (B, T, H), 0.1),
           "skip": set(info["skipped"])}
    for l in range(12):
        reg[("attn", l)] = mask(l, w2v.SITE_ATTN, (B, T, H), 0.1)
        reg[("act", l)] = mask(l, w2v.SITE_ACT, (B, T, Fi), 0.1)
        reg[("ffn", l)] = mask(l, w2v.SITE_FFN, (B, T, H), 0.1)
        reg[("attp", l)] = ops.attention_dropout_mask(B, T, 12, 0.1, w2v.drop_seed(step, l, w2v.SITE_ATTN_P), cuda).cpu()
    spec = torch.zeros(B * T, dtype=torch.bool)
    spec[info["spec_rows"].cpu()] = True
    reg["spec"] = spec.view(B, T)
    reg["spec_feat"] = info["spec_keep"].view(B, H).cpu() == 0
    assert reg["spec_feat"].any() and not reg["spec_feat"].all()
    fin = mask(-1, w2v.SITE_HEAD_A, (B, T, H), 0.1)
    # oracle replay (torch CPU fp32 autograd)
    sd = {k: v.clone() for k, v in sd0.items()}
    trainable = [k for k in sd if not k.startswith("feature_extractor.")]
    for k in trainable:
        sd[k].requires_grad_(True)
    hw_, hb_ = hw.clone().requires_grad_(True), hb.clone().requires_grad_(True)
    torch.set_num_threads(os.cpu_count())
    hidden, _, flen = ow.forward(sd, cfg, wav3, lens3, return_features=True, reg=reg)
    logits = F.linear(hidden[-1] * fin, hw_, hb_)
    lp = F.log_softmax(logits, dim=-1, dtype=torch.float32).transpose(0, 1)
    loss = F.ctc_loss(lp, labels[labels >= 0], flen, (labels >= 0).sum(-1), blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    assert abs(float(r["loss"].detach()) - float(loss)) / float(loss) < 2e-3, (float(r["loss"].detach()), float(loss))
    params = dict(pr.named_parameters())
    norms = {k: float(sd[k].grad.double().norm()) for k in trainable if sd[k].grad is not None}
    floor = 1e-4 * float(np.median(list(norms.values())))
    worst, low = ("", 0.0), ("", 1.0)
    for k, ref in norms.items():
        ours = params["wav2vec2." + k].grad.double().cpu()
        if ref < floor:
            assert float(ours.norm()) < 100 * floor, k
            continue
        rel = abs(float(ours.norm()) - ref) / ref
        cos = float((ours.flatten() @ sd[k].grad.double().flatten()) / (ours.norm() * ref))
        worst = max(worst, (k, rel), key=lambda t: t[1])
        low = min(low, (k, cos), key=lambda t: t[1])
    print("regularised step: worst grad-norm deviation", worst, "lowest cosine", low)
    assert worst[1] < 0.05 and low[1] > 0.98
    for l in info["skipped"]:                       # LayerDrop: a dropped layer receives no gradient
        assert float(params[f"wav2vec2.encoder.layers.{l}.feed_forward.output_dense.weight"].grad.abs().max()) == 0.0
    assert norms["masked_spec_embed"] > floor      # SpecAugment rows feed masked_spec_embed


def test_aptai_head_dropouts_replayed_by_the_oracle(cuda):
    """APTAI with tv_drop / phn_drop (models/aptai.py:43-55): the two head masks are materialised and replayed."""
    import torch.nn.functional as F
    from aptai_b200 import ops
    from oracle import heads as oh
    from oracle import w2v2 as ow
    cfg = cfg_base(vocab_size=46)
    sd0 = backbone_sd(cfg, 1)
    name = register_in_memory_checkpoint("mem://base-seed1-heads", sd0)
    m = APTAI(cuda, VOCAB, name, cfg, None, phn_drop=0.2, tv_drop=0.1)
    tvw, tvb = W.linear_params(111, 9, 768)
    pw, pb = W.linear_params(112, 46, 768)
    with torch.no_grad():
        m.tv_head[2].weight.copy_(tvw); m.tv_head[2].bias.copy_(tvb)
        m.phn_head[2].weight.copy_(pw); m.phn_head[2].bias.copy_(pb)
    m = m.to(cuda).train()
    wav, lens, phn, tvs = _g2_inputs()
    out = m(0, wav.to(cuda), lens.to(cuda), phn.to(cuda), *[t.to(cuda) for t in tvs])
    out["loss"].backward()
    w2v = m.wav2vec2
    step, B, T, H = w2v._last_regularisers["step"], 2, 99, 768
    mk = lambda site, p: ops.dropout(torch.ones((B, T, H), device=cuda), p, w2v.drop_seed(step, -1, site),
                                     want_f32=True)[0].cpu()
    m_tv, m_phn = mk(w2v.SITE_HEAD_A, 0.1), mk(w2v.SITE_HEAD_B, 0.2)
    sd = {k: v.clone() for k, v in sd0.items()}
    trainable = [k for k in sd if not k.startswith("feature_extractor.") and k != "masked_spec_embed"]
    for k in trainable:
        sd[k].requires_grad_(True)
    ws = [t.clone().requires_grad_(True) for t in (tvw, tvb, pw, pb)]
    torch.set_num_threads(os.cpu_count())
    h = ow.forward(sd, cfg, wav, lens.tolist())[-1]
    tv = oh.lowpass(F.linear(torch.tanh(h * m_tv), ws[0], ws[1]), oh.lowpass_taps())
    logits = F.linear(F.leaky_relu(h * m_phn), ws[2], ws[3])
    loss, _, _ = oh.aptai_losses(tv, logits, phn, torch.stack(tvs, -1).float())
    loss.backward()
    assert abs(float(out["loss"].detach()) - float(loss.detach())) / float(loss.detach()) < 2e-3
    params = dict(m.named_parameters())
    for name_, ref in (("tv_head.2.weight", ws[0]), ("tv_head.2.bias", ws[1]), ("phn_head.2.weight", ws[2]),
                       ("phn_head.2.bias", ws[3]), ("wav2vec2.encoder.layers.11.feed_forward.output_dense.weight",
                                                    sd["encoder.layers.11.feed_forward.output_dense.weight"]),
                       ("wav2vec2.encoder.layers.0.attention.v_proj.weight", sd["encoder.layers.0.attention.v_proj.weight"])):
        ours, r = params[name_].grad.double().cpu().flatten(), ref.grad.double().flatten()
        cos = float(ours @ r / (ours.norm() * r.norm()))
        assert abs(float(ours.norm() / r.norm()) - 1) < 0.03 and cos > 0.995, (name_, float(ours.norm() / r.norm()), cos)


@pytest.mark.parametrize("variant", ["layer", "group"])
def test_unfrozen_conv_encoder_training_vs_oracle(cuda, variant):
    """Recogniser training with the conv feature encoder UNFROZEN (the reference's default,
    train/train_phoneme_recognizer.py:170) on a 'layer'-norm backbone: gradients of all seven conv layers (weights,
    biases, LayerNorms) and of the rest of the model against the oracle's autograd."""
    import torch.nn.functional as F
    from oracle import w2v2 as ow
    cfg = cfg_large(vocab_size=46, num_hidden_layers=2) if variant == "layer" else cfg_base(vocab_size=46,
                                                                                             num_hidden_layers=2)
    sd0 = backbone_sd(cfg, 5)
    name = register_in_memory_checkpoint(f"mem://{variant}2-seed5-conv", sd0)
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(105, 46, cfg.hidden_size)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr = pr.to(cuda).train()                                   # feature encoder NOT frozen
    lens = [24000, 17000]
    wav = W.waveforms(2, 24000, lens, seed=4321)
    labels, _ = W.phoneme_sequences(2, 8, 20, 2, 45, seed=9, pad=-100)
    r = pr(wav.to(cuda), torch.tensor(lens, device=cuda), labels.to(cuda))
    r["loss"].backward()
    sd = {k: v.clone() for k, v in sd0.items()}
    trainable = [k for k in sd if k != "masked_spec_embed"]
    for k in trainable:
        sd[k].requires_grad_(True)
    hw_, hb_ = hw.clone().requires_grad_(True), hb.clone().requires_grad_(True)
    torch.set_num_threads(os.cpu_count())
    hidden, _, flen = ow.forward(sd, cfg, wav, lens, return_features=True)
    lp = F.log_softmax(F.linear(hidden[-1], hw_, hb_), dim=-1, dtype=torch.float32).transpose(0, 1)
    loss = F.ctc_loss(lp, labels[labels >= 0], flen, (labels >= 0).sum(-1), blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    assert abs(float(r["loss"].detach()) - float(loss.detach())) / float(loss.detach()) < 2e-3
    params = dict(pr.named_parameters())
    norms = {k: float(sd[k].grad.double().norm()) for k in trainable}
    floor = 1e-4 * float(np.median(list(norms.values())))
    report = {}
    for k, ref in norms.items():
        ours = params["wav2vec2." + k].grad.double().cpu().flatten()
        if ref < floor:
            continue
        g = sd[k].grad.double().flatten()
        report[k] = (abs(float(ours.norm()) - ref) / ref, float(ours @ g / (ours.norm() * g.norm())))
    conv = {k: v for k, v in report.items() if k.startswith("feature_extractor.")}
    assert len(conv) >= (7 * 3 if variant == "layer" else 7 + 2)     # conv weights (+ norms) of all seven layers
    worst = max(report.items(), key=lambda t: t[1][0])
    low = min(report.items(), key=lambda t: t[1][1])
    print("unfrozen conv: worst grad-norm deviation", worst, "lowest cosine", low)
    assert worst[1][0] < 0.05 and low[1][1] > 0.98


@pytest.mark.parametrize("variant,lens", [("layer", [16000]), ("group", [320000]), ("layer", [100000, 6000, 52000])])
def test_training_edge_shapes_vs_oracle(cuda, variant, lens):
    """Edge shapes of the training step against the oracle's autograd (2-layer backbones, conv encoder frozen):
    a single 1 s utterance (49 frames: fewer rows than one GEMM / wgrad tile), the 20 s maximum (999 frames: eight
    attention tiles in both directions), and a ragged batch whose shortest utterance is 18 frames long."""
    import torch.nn.functional as F
    from oracle import w2v2 as ow
    mk = cfg_large if variant == "layer" else cfg_base
    cfg = mk(vocab_size=46, num_hidden_layers=2)
    sd0 = backbone_sd(cfg, 6)
    name = register_in_memory_checkpoint(f"mem://{variant}2-seed6-edge", sd0)
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(106, 46, cfg.hidden_size)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.wav2vec2.freeze_feature_encoder()
    pr = pr.to(cuda).train()
    B, L = len(lens), max(lens)
    wav = W.waveforms(B, L, lens, seed=777)
    labels, _ = W.phoneme_sequences(B, 3, 8, 2, 45, seed=13, pad=-100)
    r = pr(wav.to(cuda), torch.tensor(lens, device=cuda), labels.to(cuda))
    r["loss"].backward()
    sd = {k: v.clone() for k, v in sd0.items()}
    trainable = [k for k in sd if not k.startswith("feature_extractor.") and k != "masked_spec_embed"]
    for k in trainable:
        sd[k].requires_grad_(True)
    torch.set_num_threads(os.cpu_count())
    hidden, _, flen = ow.forward(sd, cfg, wav, lens, return_features=True)
    lp = F.log_softmax(F.linear(hidden[-1], hw, hb), dim=-1, dtype=torch.float32).transpose(0, 1)
    loss = F.ctc_loss(lp, labels[labels >= 0], flen, (labels >= 0).sum(-1), blank=0, reduction="mean", zero_infinity=True)
    loss.backward()
    assert abs(float(r["loss"].detach()) - float(loss.detach())) / float(loss.detach()) < 2e-3
    params = dict(pr.named_parameters())
    norms = {k: float(sd[k].grad.double().norm()) for k in trainable}
    floor = 1e-4 * float(np.median(list(norms.values())))
    worst, low = ("", 0.0), ("", 1.0)
    for k, ref in norms.items():
        if ref < floor:
            continue
        ours = params["wav2vec2." + k].grad.double().cpu().flatten()
        g = sd[k].grad.double().flatten()
        worst = max(worst, (k, abs(float(ours.norm()) - ref) / ref), key=lambda t: t[1])
        low = min(low, (k, float(ours @ g / (ours.norm() * g.norm()))), key=lambda t: t[1])
    print(f"edge {variant} {lens}: worst grad-norm deviation", worst, "lowest cosine", low)
    assert worst[1] < 0.05 and low[1] > 0.98


@pytest.mark.parametrize("B,drop", [(1, False), (3, True)])
def test_force_aptai_training_step_vs_oracle(cuda, B, drop):
    """Force_APTAI in train mode (train/train_force_aptai.py): loss.backward() runs the hand-written backward of the
    tail (low-pass adjoint, head MLP, BiLSTM through time, cross-attention, embedding, frame projection); the frozen
    recogniser gets no gradient.  Oracle: oracle/force_tail.py (torch autograd on the CPU) fed with the recogniser's
    hidden states, the same phoneme sequences and — for the stochastic case — the three dropout masks the step used,
    materialised from the counter-based generator.  Followed by one fused Adam step."""
    from helpers import force_tail_state
    from aptai_b200 import Force_APTAI, ops
    from oracle.force_tail import ForceTail
    cfg = cfg_base(vocab_size=46)
    name = register_in_memory_checkpoint("mem://base-seed1-force", backbone_sd(cfg, 1))
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(103, 46, cfg.hidden_size)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    fa = Force_APTAI("unused", cuda, VOCAB, w2v2_pr=pr)
    tail = force_tail_state(fa.state_dict())
    fa.load_state_dict(tail, strict=False)
    fa = fa.to(cuda).train()
    if not drop:
        fa.frame_drop.p = fa.pe_phn.dropout.p = fa.rnn.linear[1].p = 0.0
    lens = [32000, 27000, 16000][:B]
    wav = W.waveforms(B, 32000, lens, seed=4321)
    seqs, _ = W.phoneme_sequences(B, 8, 30, 1, 45, seed=17, pad=0)
    seqs = [np.asarray(s[s != 0], dtype=np.int64) for s in seqs.numpy()]
    T = 99
    g = torch.Generator().manual_seed(9)
    tvt = torch.randn((B, T, 9), generator=g)
    flen = [cfg.conv_out_length(n) for n in lens]
    for b in range(B):
        tvt[b, flen[b]:] = -100.0
    opt = FusedAdam([p for p in fa.parameters() if p.requires_grad], lr=1e-3)
    opt.zero_grad()
    torch.manual_seed(77)
    tv_cols = [tvt[:, :, i].contiguous().to(cuda) for i in range(9)]
    res = fa(0, wav.to(cuda), torch.tensor(lens, device=cuda), None, None, *tv_cols, phn_seqs=seqs)
    res["loss"].backward()
    torch.cuda.synchronize()
    assert all(p.grad is None for p in fa.w2v2_pr.parameters())
    # ---- oracle replay
    with torch.no_grad():
        _, h, _ = pr._logits(wav.to(cuda), torch.tensor(lens, device=cuda))
    ids = torch.zeros((B, 60), dtype=torch.int64)
    for b, s in enumerate(seqs):
        ids[b, : len(s)] = torch.from_numpy(s)
    reg = {}
    if drop:
        s_f, s_pe, s_r = fa._last_train["seeds"]
        mk = lambda shape, p, seed: ops.dropout(torch.ones(shape, device=cuda), p, seed, want_f32=True)[0].cpu()
        reg = {"frame": mk((B, T, 128), 0.2, s_f), "pe": mk((B, 60, 128), 0.2, s_pe), "rnn": mk((B, T, 256), 0.1, s_r)}
    ref = ForceTail(cfg.hidden_size, len(VOCAB))
    ref.load_state_dict(tail, strict=True)
    out = ref(h.cpu(), ids, flen, [len(s) for s in seqs], tvt, reg)
    out["loss"].backward()
    for k in ("loss", "tv_loss", "align_loss"):
        assert abs(float(res[k].detach()) - float(out[k].detach())) / abs(float(out[k].detach())) < 2e-3, (k, float(res[k].detach()), float(out[k].detach()))
    torch.testing.assert_close(res["tvs_pred"].cpu(), out["tvs"].detach(), atol=2e-3, rtol=1e-3)
    params = dict(fa.named_parameters())
    worst, low = ("", 0.0), ("", 1.0)
    for k, p_ref in ref.named_parameters():
        got, want = params[k].grad.double().cpu().flatten(), p_ref.grad.double().flatten()
        rel = abs(float(got.norm()) - float(want.norm())) / float(want.norm())
        cos = float(got @ want / (got.norm() * want.norm()))
        worst = max(worst, (k, rel), key=lambda t: t[1])
        low = min(low, (k, cos), key=lambda t: t[1])
    print(f"Force_APTAI training B={B} dropout={drop}: worst grad-norm deviation", worst, "lowest cosine", low)
    assert worst[1] < 0.03 and low[1] > 0.995
    assert float(params["phn_emb_layer.weight"].grad[0].abs().max()) == 0.0          # padding_idx row
    before = params["rnn.lstm.weight_hh_l0"].detach().clone()
    opt.step()
    assert not torch.equal(before, params["rnn.lstm.weight_hh_l0"].detach())


def test_force_aptai_training_step_vs_reference(cuda):
    """Force_APTAI training step against gradients of the reference's own class (golden_force_train_v1.npz: batch 1,
    24x1024 recogniser, dropouts at p = 0)."""
    from helpers import force_tail_state
    from aptai_b200 import Force_APTAI
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_force_train_v1.npz"))
    cfg = cfg_large(vocab_size=46)
    name = register_in_memory_checkpoint("mem://large-seed0", backbone_sd(cfg, 0))
    pr = Wav2Vec2_PR(cfg, None, name, VOCAB)
    hw, hb = W.linear_params(103, 46, 1024)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    fa = Force_APTAI("unused", cuda, VOCAB, w2v2_pr=pr)
    fa.load_state_dict(force_tail_state(fa.state_dict()), strict=False)
    fa = fa.to(cuda).train()
    fa.frame_drop.p = fa.pe_phn.dropout.p = fa.rnn.linear[1].p = 0.0
    wav = W.waveforms(1, 32000, None, seed=5151)
    tvt = torch.from_numpy(g["tvt"]).to(cuda)
    res = fa(0, wav.to(cuda), torch.tensor([32000], device=cuda), None, None,
             *[tvt[:, :, i].contiguous() for i in range(9)], phn_seqs=[g["known"]])
    res["loss"].backward()
    got = [float(res[k].detach()) for k in ("loss", "tv_loss", "align_loss")]
    np.testing.assert_allclose(got, g["losses"], rtol=5e-3)
    norms = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    params = dict(fa.named_parameters())
    worst, low = ("", 0.0), ("", 1.0)
    for k, n_ref in norms.items():
        gr = params[k].grad.double().cpu().flatten()
        worst = max(worst, (k, abs(float(gr.norm()) - n_ref) / n_ref), key=lambda t: t[1])
        sl = torch.from_numpy(g[f"grad::{k}"]).double()
        low = min(low, (k, float(gr[:256] @ sl / (gr[:256].norm() * sl.norm()))), key=lambda t: t[1])
    print("Force_APTAI training vs reference: worst grad-norm deviation", worst, "lowest slice cosine", low)
    assert worst[1] < 0.05 and low[1] > 0.98
