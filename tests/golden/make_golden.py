"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own classes (read-only import from
/root/reference, transformers 5.5.0 / torch CPU fp32) on deterministic synthetic weights and inputs.

Run once in the build container:   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests only read the committed .npz files and regenerate the same weights and
inputs from oracle/weights.py (NumPy PCG64 streams are platform-stable).

Recipe follows SURVEY.md Appendix C: import transformers first, stub editdistance/librosa so `utility` imports,
use a local save_pretrained directory as `huggingface_model_id`, inject a phoneme decoder (flashlight-text is not
installable here).
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch
import transformers  # noqa: F401  (must precede the stubs)
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
for n in ["editdistance", "librosa", "librosa.filters", "librosa.sequence"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["librosa.filters"].mel = None
sys.modules["librosa.sequence"].dtw = None
REF = os.environ.get("APTAI_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "models"), REF]

from transformers import Wav2Vec2Config, Wav2Vec2Model  # noqa: E402

import aptai as ref_aptai  # noqa: E402
import force_aptai as ref_force  # noqa: E402
import modules as ref_modules  # noqa: E402
import w2v2_pr as ref_pr  # noqa: E402

from aptai_b200.config import W2V2Config  # noqa: E402
from oracle import weights as W  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(os.cpu_count())
VOCAB = {"(blank)": 0, "(...)": 1, **{f"p{i}": i for i in range(2, 46)}}
NO_REG = dict(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
              final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)


def hf_config(variant):
    kw = dict(vocab_size=46, **NO_REG)
    if variant == "large":
        kw.update(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)
    d = Wav2Vec2Config(**kw).to_dict()
    d.update(blank=0, ctc_zero_infinity=True, ctc_loss_reduction="mean")
    return Wav2Vec2Config.from_dict(d)


def save_backbone(hf_cfg, seed, path):
    sd = W.backbone_state_dict(W2V2Config.from_any(hf_cfg), seed)
    m = Wav2Vec2Model(hf_cfg)
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    m.save_pretrained(path)
    return sd


def main():
    tmp = tempfile.mkdtemp(prefix="aptai_golden_")
    out = {}

    # ------------------------------------------------------------------ G1/G2: APTAI on the 24x1024 'layer' backbone
    cfg_l = hf_config("large")
    dir_l = os.path.join(tmp, "large")
    save_backbone(cfg_l, 0, dir_l)
    m = ref_aptai.APTAI(torch.device("cpu"), VOCAB, dir_l, cfg_l, None, phn_drop=0.0, tv_drop=0.0)
    tvw, tvb = W.linear_params(101, 9, 1024)
    pw, pb = W.linear_params(102, 46, 1024)
    with torch.no_grad():
        m.tv_head[2].weight.copy_(tvw); m.tv_head[2].bias.copy_(tvb)
        m.phn_head[2].weight.copy_(pw); m.phn_head[2].bias.copy_(pb)
    m.eval()
    wav1 = W.waveforms(1, 32000, None, seed=1234)
    r = m.get_aptai_output(wav1[0].numpy())
    out["g1_logits"] = r["phn_fc_logits"]
    out["g1_pred"] = r["phn_fc_pred"]
    out["g1_probs"] = r["phn_fc_probs"]
    out["g1_tvs"] = np.stack([np.asarray(r["tvs_pred"][k], dtype=np.float32) for k in
                              ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")], axis=-1)
    print("G1", out["g1_logits"].shape, out["g1_probs"].shape, out["g1_tvs"].shape)

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
    out["g2_phn"], out["g2_tvt"] = phn, tvt
    with torch.no_grad():
        r2 = m(0, wav2, torch.tensor(lens2), torch.from_numpy(phn), *[torch.from_numpy(tvt[:, :, i]) for i in range(9)])
    out["g2_losses"] = np.asarray([float(r2["loss"]), float(r2["mse_loss"]), float(r2["ce_loss"])], dtype=np.float64)
    out["g2_tvs"] = r2["tvs_pred"].numpy()
    out["g2_pred"] = r2["phn_fc_pred"].numpy()
    print("G2 losses", out["g2_losses"])

    # ------------------------------------------------------------------ G4: Force_APTAI on a PR model sharing that backbone
    pr_l = ref_pr.Wav2Vec2_PR(cfg_l, None, dir_l, VOCAB)
    hw, hb = W.linear_params(103, 46, 1024)
    with torch.no_grad():
        pr_l.pr_head.weight.copy_(hw); pr_l.pr_head.bias.copy_(hb)
    ck = os.path.join(tmp, "pr_large", "best-model-ckpt")
    os.makedirs(ck)
    torch.save(pr_l.state_dict(), os.path.join(ck, "pytorch_model.bin"))
    pickle.dump(pr_l.get_config(), open(os.path.join(ck, "model_cfg.pkl"), "wb"))
    phn_seq, _ = W.phoneme_sequences(1, 30, 30, 1, 45, seed=11, pad=0)
    known = phn_seq[0].numpy().astype(np.int64)

    class _Hyp:
        def __init__(self, toks):
            self.tokens = torch.as_tensor(toks)
            self.timesteps = torch.arange(len(toks))

    def fake_decoder(**kw):
        return lambda emissions: [[_Hyp(known)] for _ in range(emissions.shape[0])]

    torchaudio.models.decoder.ctc_decoder = fake_decoder
    fa = ref_force.Force_APTAI(os.path.join(tmp, "pr_large"), "cpu", VOCAB)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import force_tail_state
    tail = force_tail_state(fa.state_dict())
    fa.load_state_dict(tail, strict=False)
    fa.eval()
    ra = fa.get_alignment(wav1[0].numpy())
    out["g4_alignment"] = ra["alignment"]
    out["g4_known"] = known
    rf = fa.get_faptai_output(wav1[0].numpy())
    out["g4_tvs"] = np.stack([np.asarray(rf["tvs_pred"][k], dtype=np.float32) for k in
                              ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")], axis=-1)
    out["g4_frame_phns"] = np.asarray(rf["pred_frame_phns"], dtype=np.int64)
    tv_t = rng.standard_normal((1, T, 9), dtype=np.float32)
    out["g4_tvt"] = tv_t
    with torch.no_grad():
        rff = fa(0, wav1, torch.tensor([32000]), None, None, *[torch.from_numpy(tv_t[:, :, i]) for i in range(9)])
    out["g4_losses"] = np.asarray([float(rff["loss"]), float(rff["tv_loss"]), float(rff["align_loss"])])
    print("G4 alignment", out["g4_alignment"].shape, "losses", out["g4_losses"])

    # ------------------------------------------------------------------ G3: Wav2Vec2_PR.forward on the 12x768 'group' backbone
    cfg_b = hf_config("base")
    dir_b = os.path.join(tmp, "base")
    save_backbone(cfg_b, 1, dir_b)
    pr = ref_pr.Wav2Vec2_PR(cfg_b, None, dir_b, VOCAB)
    hw, hb = W.linear_params(104, 46, 768)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.eval()
    lens3 = [32000, 27000, 16000]
    wav3 = W.waveforms(3, 32000, lens3, seed=3234)
    labels, _ = W.phoneme_sequences(3, 10, 40, 2, 45, seed=7, pad=-100)
    labels[2, 5:] = -100          # short
    r3 = pr(wav3, torch.tensor(lens3), labels)
    r3["phoneme_logits"].retain_grad()
    r3["loss"].backward()
    out["g3_labels"] = labels.numpy()
    out["g3_loss"] = np.asarray([float(r3["loss"])])
    out["g3_logits"] = r3["phoneme_logits"].detach().numpy()
    out["g3_log_probs"] = r3["log_probs"].detach().numpy()
    out["g3_grad_logits"] = r3["phoneme_logits"].grad.numpy()
    out["g3_hidden"] = r3["hidden_states"].detach().numpy()[:, ::8, ::16].copy()    # sub-sampled (fixture size)
    print("G3 loss", out["g3_loss"], out["g3_logits"].shape)
    # single-utterance C1 shape: base backbone, 4 s
    wav_c1 = W.waveforms(1, 64000, None, seed=1234)
    out["c1_logits"] = pr.get_ctc_logits(wav_c1[0].numpy())
    print("C1", out["c1_logits"].shape)

    # ------------------------------------------------------------------ G5: modules
    lp = ref_modules.LowPassFilterLayer("cpu", 10, 49, 9)
    out["g5_taps"] = lp.lowpass.weight.detach().numpy().reshape(-1)
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(5)).standard_normal((2, 120, 9), dtype=np.float32))
    out["g5_lp_in"] = x.numpy()
    out["g5_lp_out"] = lp(x).numpy()
    att = torch.log_softmax(torch.from_numpy(
        np.random.Generator(np.random.PCG64(6)).standard_normal((3, 1, 80, 60), dtype=np.float32)), -1)
    out["g5_fs_in"] = att.numpy()
    out["g5_fs_text"] = np.asarray([30, 59, 5])
    out["g5_fs_mel"] = np.asarray([80, 70, 33])
    out["g5_fs_loss"] = np.asarray([float(ref_modules.ForwardSumLoss()(att, [30, 59, 5], [80, 70, 33]))])
    pe = ref_modules.PositionalEncoding(128, 0.0, 60)
    out["g5_pe"] = pe.pe.numpy()

    # ------------------------------------------------------------------ G6: Viterbi vs torchaudio.functional.forced_align
    g6 = np.random.Generator(np.random.PCG64(9))
    vit = []
    for trial in range(24):
        Tt, C = int(g6.integers(6, 60)), int(g6.integers(3, 10))
        L = int(g6.integers(1, max(2, Tt // 3)))
        mode = trial % 3
        x = g6.standard_normal((Tt, C)).astype(np.float32) if mode == 0 else (
            np.zeros((Tt, C), np.float32) if mode == 1 else g6.integers(0, 3, (Tt, C)).astype(np.float32))
        tg = g6.integers(1, C, size=L).astype(np.int32)
        if L > 2 and trial % 2:
            tg[1] = tg[0]
        lpx = torch.log_softmax(torch.from_numpy(x), -1)
        p, s = torchaudio.functional.forced_align(lpx[None], torch.from_numpy(tg)[None], blank=0)
        vit.append((x, tg, p[0].numpy().astype(np.int32), s[0].numpy()))
    out["g6_n"] = np.asarray([len(vit)])
    for i, (x, tg, p, s) in enumerate(vit):
        out[f"g6_x{i}"], out[f"g6_t{i}"], out[f"g6_p{i}"], out[f"g6_s{i}"] = x, tg, p, s

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
