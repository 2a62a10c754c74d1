"""Golden vectors for the Force_APTAI TRAINING step: loss and gradients of the trainable tail produced by the
REFERENCE's own class (read-only import from /root/reference, transformers 5.5.0 / torch CPU fp32) in train mode with
its three dropouts set to p = 0 (torch's RNG stream cannot be replayed elsewhere), batch 1 (the reference's RNN
raises NameError for batch > 1, models/modules.py:207), 24x1024 'layer' recogniser (the reference hard-codes frame_lin = Linear(1024, 128)), 2 s utterance, a known 22-phoneme
sequence injected through a stub of the absent flashlight decoder.

Run once in the build container:   python tests/golden/make_golden_force_train.py
Stored: the three losses, the L2 norm of every tail gradient (by state_dict name) and its first 256 entries.  The
GPU box regenerates weights/inputs from aptai_b200.synth / tests/helpers.force_tail_state.
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
import torchaudio
import transformers  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for n in ["editdistance", "librosa", "librosa.filters", "librosa.sequence"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["librosa.filters"].mel = None
sys.modules["librosa.sequence"].dtw = None
REF = os.environ.get("APTAI_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "models"), REF]

import force_aptai as ref_force  # noqa: E402
import w2v2_pr as ref_pr  # noqa: E402

from helpers import force_tail_state  # noqa: E402
from make_golden import VOCAB, hf_config, save_backbone  # noqa: E402
from oracle import weights as W  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    tmp = tempfile.mkdtemp(prefix="aptai_golden_force_train_")
    cfg = hf_config("large")
    d = os.path.join(tmp, "large")
    save_backbone(cfg, 0, d)
    pr = ref_pr.Wav2Vec2_PR(cfg, None, d, VOCAB)
    hw, hb = W.linear_params(103, 46, 1024)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    ck = os.path.join(tmp, "pr_large", "best-model-ckpt")
    os.makedirs(ck)
    torch.save(pr.state_dict(), os.path.join(ck, "pytorch_model.bin"))
    pickle.dump(pr.get_config(), open(os.path.join(ck, "model_cfg.pkl"), "wb"))
    phn_seq, _ = W.phoneme_sequences(1, 22, 22, 1, 45, seed=19, pad=0)
    known = phn_seq[0].numpy().astype(np.int64)

    class _Hyp:
        def __init__(self, toks):
            self.tokens = torch.as_tensor(toks)
            self.timesteps = torch.arange(len(toks))

    torchaudio.models.decoder.ctc_decoder = lambda **kw: (lambda em: [[_Hyp(known)] for _ in range(em.shape[0])])
    fa = ref_force.Force_APTAI(os.path.join(tmp, "pr_large"), "cpu", VOCAB)
    fa.load_state_dict(force_tail_state(fa.state_dict()), strict=False)
    fa.train()
    fa.frame_drop.p = 0.0
    fa.pe_phn.dropout.p = 0.0
    fa.rnn.linear[1].p = 0.0
    wav = W.waveforms(1, 32000, None, seed=5151)
    T = 99
    tvt = np.random.Generator(np.random.PCG64(23)).standard_normal((1, T, 9), dtype=np.float32)
    res = fa(0, wav, torch.tensor([32000]), None, None, *[torch.from_numpy(tvt[:, :, i]) for i in range(9)])
    res["loss"].backward()
    out = {"known": known, "tvt": tvt,
           "losses": np.asarray([float(res["loss"]), float(res["tv_loss"]), float(res["align_loss"])])}
    names, norms = [], []
    for n, p in fa.named_parameters():
        if p.grad is None:
            continue
        assert not n.startswith("w2v2_pr."), n
        names.append(n)
        norms.append(float(p.grad.double().norm()))
        out[f"grad::{n}"] = p.grad.reshape(-1)[:256].numpy().copy()
    out["grad_names"] = np.asarray(names)
    out["grad_norms"] = np.asarray(norms, dtype=np.float64)
    path = os.path.join(HERE, "golden_force_train_v1.npz")
    np.savez_compressed(path, **out)
    print("losses", out["losses"], "|", len(names), "gradients ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
