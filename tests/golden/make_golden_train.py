"""Golden vectors for the TRAINING step: gradients produced by the REFERENCE's own classes (read-only import from
/root/reference, transformers 5.5.0 / torch CPU fp32, train mode, all stochastic regularisers off) on the same
deterministic weights and inputs as golden_v1 (G2: APTAI 24x1024, B=2; G3: Wav2Vec2_PR 12x768, B=3).

Run once in the build container:   python tests/golden/make_golden_train.py
Stored per case: the loss, the L2 norm of every parameter gradient (by state_dict name) and the first 256 entries
of a few representative gradients.  The GPU box regenerates weights/inputs from aptai_b200.synth.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import transformers  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for n in ["editdistance", "librosa", "librosa.filters", "librosa.sequence"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["librosa.filters"].mel = None
sys.modules["librosa.sequence"].dtw = None
REF = os.environ.get("APTAI_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "models"), REF]

import aptai as ref_aptai  # noqa: E402
import w2v2_pr as ref_pr  # noqa: E402

from make_golden import VOCAB, hf_config, save_backbone  # noqa: E402
from oracle import weights as W  # noqa: E402

SLICES = ("encoder.layers.0.attention.q_proj.weight", "encoder.layers.5.feed_forward.intermediate_dense.weight",
          "encoder.layers.11.feed_forward.output_dense.bias", "encoder.layers.3.layer_norm.weight",
          "encoder.pos_conv_embed.conv.parametrizations.weight.original1",
          "encoder.pos_conv_embed.conv.parametrizations.weight.original0",
          "feature_projection.projection.weight", "feature_projection.layer_norm.weight", "encoder.layer_norm.bias")


def collect(model, out, tag):
    names, norms = [], []
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(n)
        norms.append(float(p.grad.double().norm()))
        short = n[len("wav2vec2."):] if n.startswith("wav2vec2.") else n
        if short in SLICES or not n.startswith("wav2vec2."):
            out[f"{tag}_grad::{n}"] = p.grad.reshape(-1)[:256].numpy().copy()
    out[f"{tag}_grad_names"] = np.asarray(names)
    out[f"{tag}_grad_norms"] = np.asarray(norms, dtype=np.float64)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    tmp = tempfile.mkdtemp(prefix="aptai_golden_train_")
    out = {}

    # ---- G2 (training): APTAI.forward + backward, 24x1024 'layer' backbone
    cfg_l = hf_config("large")
    dir_l = os.path.join(tmp, "large")
    save_backbone(cfg_l, 0, dir_l)
    m = ref_aptai.APTAI(torch.device("cpu"), VOCAB, dir_l, cfg_l, None, phn_drop=0.0, tv_drop=0.0)
    tvw, tvb = W.linear_params(101, 9, 1024)
    pw, pb = W.linear_params(102, 46, 1024)
    with torch.no_grad():
        m.tv_head[2].weight.copy_(tvw); m.tv_head[2].bias.copy_(tvb)
        m.phn_head[2].weight.copy_(pw); m.phn_head[2].bias.copy_(pb)
    m.train()
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
    r2 = m(0, wav2, torch.tensor(lens2), torch.from_numpy(phn), *[torch.from_numpy(tvt[:, :, i]) for i in range(9)])
    r2["loss"].backward()
    out["t2_losses"] = np.asarray([float(r2["loss"]), float(r2["mse_loss"]), float(r2["ce_loss"])])
    collect(m, out, "t2")
    print("T2 loss", out["t2_losses"], "params with grad", len(out["t2_grad_names"]))
    # one Adam step of the reference's optimizer, then the loss again (train/train_aptai.py:350-356 defaults)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    opt.step()
    with torch.no_grad():
        r2b = m(0, wav2, torch.tensor(lens2), torch.from_numpy(phn), *[torch.from_numpy(tvt[:, :, i]) for i in range(9)])
    out["t2_losses_after_step"] = np.asarray([float(r2b["loss"]), float(r2b["mse_loss"]), float(r2b["ce_loss"])])
    print("T2 after one Adam step", out["t2_losses_after_step"])
    del m

    # ---- G3 (training): Wav2Vec2_PR.forward + backward, 12x768 'group' backbone
    cfg_b = hf_config("base")
    dir_b = os.path.join(tmp, "base")
    save_backbone(cfg_b, 1, dir_b)
    pr = ref_pr.Wav2Vec2_PR(cfg_b, None, dir_b, VOCAB)
    hw, hb = W.linear_params(104, 46, 768)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.wav2vec2.freeze_feature_encoder()
    pr.train()
    lens3 = [32000, 27000, 16000]
    wav3 = W.waveforms(3, 32000, lens3, seed=3234)
    labels, _ = W.phoneme_sequences(3, 10, 40, 2, 45, seed=7, pad=-100)
    labels[2, 5:] = -100
    r3 = pr(wav3, torch.tensor(lens3), labels)
    r3["loss"].backward()
    out["t3_loss"] = np.asarray([float(r3["loss"])])
    collect(pr, out, "t3")
    print("T3 loss", out["t3_loss"], "params with grad", len(out["t3_grad_names"]))

    path = os.path.join(HERE, "golden_train_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
