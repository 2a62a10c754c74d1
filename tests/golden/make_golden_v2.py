"""Golden fixtures at the BASELINE config sizes, produced by the REFERENCE's own classes (read-only import from
/root/reference, transformers 5.5.0 / torch CPU fp32) on the deterministic synthetic weights / inputs of
aptai_b200/synth.py.  Complements make_golden.py (whose cases stop at 2 s / 4 s):

  L8   APTAI.get_aptai_output, 24x1024, one 8 s utterance  (T = 399: the query-tile-pair attention kernel)
  L20  APTAI.get_aptai_output, 24x1024, one 20 s utterance (T = 999: the longest utterance of BASELINE config 5)
  R4   APTAI.forward, 24x1024, ragged batch of 4 x <= 8 s  (padding semantics, masked losses)
  C2   Wav2Vec2_PR.forward, 24x1024, 16 x <= 8 s, CTC loss + d loss / d logits (BASELINE config 2)
  F8   Force_APTAI (B = 1: the reference's RNN raises NameError for B > 1), 8 s, 45 known phonemes
  API  Wav2Vec2_PR.get_embeddings / pred_phn_seq / predict_phonemes_durations / get_embeddings_grad on the 12x768
       backbone with the decoder replaced by oracle/ctc_decode.py (flashlight-text is not installable here)

Run once in the build container:   python tests/golden/make_golden_v2.py     (about 10 minutes of CPU)
"""
import os
import pickle
import sys
import tempfile
import time
import types

import numpy as np
import torch
import transformers  # noqa: F401  (must precede the stubs)
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
for n in ["editdistance", "librosa", "librosa.filters", "librosa.sequence"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["librosa.filters"].mel = None
sys.modules["librosa.sequence"].dtw = None
REF = os.environ.get("APTAI_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "models"), REF]

import aptai as ref_aptai  # noqa: E402
import force_aptai as ref_force  # noqa: E402
import w2v2_pr as ref_pr  # noqa: E402

from make_golden import NO_REG, VOCAB, hf_config, save_backbone  # noqa: E402,F401
from oracle import ctc_decode  # noqa: E402
from oracle import weights as W  # noqa: E402

TVN = ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")
torch.set_num_threads(os.cpu_count())


def margin(logits):
    s = np.sort(logits, -1)
    return (s[..., -1] - s[..., -2]).astype(np.float32)


def main():
    tmp = tempfile.mkdtemp(prefix="aptai_golden2_")
    out = {}
    t0 = time.time()
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

    # ---- L8 / L20: single utterances at 8 s and 20 s
    for tag, L, seed in (("l8", 128000, 4234), ("l20", 320000, 5234)):
        wav = W.waveforms(1, L, None, seed=seed)
        r = m.get_aptai_output(wav[0].numpy())
        out[f"{tag}_logits"] = r["phn_fc_logits"].astype(np.float32)
        out[f"{tag}_tvs"] = np.stack([np.asarray(r["tvs_pred"][k], dtype=np.float32) for k in TVN], -1)
        print(tag, out[f"{tag}_logits"].shape, f"{time.time() - t0:.0f}s", flush=True)

    # ---- R4: ragged batch through APTAI.forward; the logits are captured by a forward hook on the phoneme head
    lens = [128000, 113000, 96000, 71000]
    flen = [399, 352, 299, 221]
    T = 399
    wav = W.waveforms(4, 128000, lens, seed=6234)
    rng = np.random.Generator(np.random.PCG64(31))
    phn = np.zeros((4, T), dtype=np.int64)
    tvt = np.full((4, T, 9), -100.0, dtype=np.float32)
    for b in range(4):
        phn[b, : flen[b]] = rng.integers(1, 46, size=flen[b])
        tvt[b, : flen[b]] = rng.standard_normal((flen[b], 9), dtype=np.float32)
    cap = {}
    hk = m.phn_head.register_forward_hook(lambda mod, i, o: cap.__setitem__("logits", o.detach().numpy().copy()))
    with torch.no_grad():
        r = m(0, wav, torch.tensor(lens), torch.from_numpy(phn), *[torch.from_numpy(tvt[:, :, i]) for i in range(9)])
    hk.remove()
    out["r4_phn"], out["r4_tvt"] = phn.astype(np.int8), tvt
    out["r4_losses"] = np.asarray([float(r["loss"]), float(r["mse_loss"]), float(r["ce_loss"])], dtype=np.float64)
    out["r4_tvs"] = r["tvs_pred"].numpy()
    out["r4_pred"] = r["phn_fc_pred"].numpy().astype(np.int8)
    out["r4_logits"] = cap["logits"].astype(np.float32)
    print("r4 losses", out["r4_losses"], f"{time.time() - t0:.0f}s", flush=True)
    del m

    # ---- C2: Wav2Vec2_PR.forward, 16 x <= 8 s, CTC loss and its gradient w.r.t. the logits
    pr = ref_pr.Wav2Vec2_PR(cfg_l, None, dir_l, VOCAB)
    hw, hb = W.linear_params(103, 46, 1024)
    with torch.no_grad():
        pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
    pr.eval()
    g2 = np.random.Generator(np.random.PCG64(41))
    lens2 = [128000] + [int(x) for x in g2.integers(96000, 128001, size=15)]
    wav2 = W.waveforms(16, 128000, lens2, seed=7234)
    labels, _ = W.phoneme_sequences(16, 10, 59, 2, 45, seed=7, pad=-100)
    r = pr(wav2, torch.tensor(lens2), labels)
    (gl,) = torch.autograd.grad(r["loss"], r["phoneme_logits"])
    lg = r["phoneme_logits"].detach().numpy()
    out["c2_lens"] = np.asarray(lens2, dtype=np.int64)
    out["c2_labels"] = labels.numpy()
    out["c2_loss"] = np.asarray([float(r["loss"])], dtype=np.float64)
    out["c2_pred"] = lg.argmax(-1).astype(np.int8)
    out["c2_margin"] = margin(lg)
    out["c2_logits_sub"] = lg[:, ::8].copy()                      # every 8th frame (fixture size)
    out["c2_grad_sub"] = gl.numpy()[:, ::8].copy()
    print("c2 loss", out["c2_loss"], f"{time.time() - t0:.0f}s", flush=True)

    # ---- F8: Force_APTAI on that recogniser, B = 1, 8 s, 45 known phonemes
    ck = os.path.join(tmp, "pr_large", "best-model-ckpt")
    os.makedirs(ck)
    torch.save(pr.state_dict(), os.path.join(ck, "pytorch_model.bin"))
    pickle.dump(pr.get_config(), open(os.path.join(ck, "model_cfg.pkl"), "wb"))
    del pr
    phn_seq, _ = W.phoneme_sequences(1, 45, 45, 1, 45, seed=12, pad=0)
    known = phn_seq[0].numpy().astype(np.int64)

    class _Hyp:
        def __init__(self, toks):
            self.tokens = torch.as_tensor(toks)
            self.timesteps = torch.arange(len(toks))

    torchaudio.models.decoder.ctc_decoder = lambda **kw: (lambda em: [[_Hyp(known)] for _ in range(em.shape[0])])
    fa = ref_force.Force_APTAI(os.path.join(tmp, "pr_large"), "cpu", VOCAB)
    from helpers import force_tail_state
    fa.load_state_dict(force_tail_state(fa.state_dict()), strict=False)
    fa.eval()
    wav8 = W.waveforms(1, 128000, None, seed=4234)
    out["f8_known"] = known
    out["f8_alignment"] = fa.get_alignment(wav8[0].numpy())["alignment"].astype(np.float32)
    rf = fa.get_faptai_output(wav8[0].numpy())
    out["f8_tvs"] = np.stack([np.asarray(rf["tvs_pred"][k], dtype=np.float32) for k in TVN], -1)
    out["f8_frame_phns"] = np.asarray(rf["pred_frame_phns"], dtype=np.int64)
    tv_t = np.random.Generator(np.random.PCG64(51)).standard_normal((1, 399, 9), dtype=np.float32)
    out["f8_tvt"] = tv_t
    with torch.no_grad():
        rff = fa(0, wav8, torch.tensor([128000]), None, None, *[torch.from_numpy(tv_t[:, :, i]) for i in range(9)])
    out["f8_losses"] = np.asarray([float(rff["loss"]), float(rff["tv_loss"]), float(rff["align_loss"])])
    print("f8 losses", out["f8_losses"], f"{time.time() - t0:.0f}s", flush=True)
    del fa

    # ---- API rows of Wav2Vec2_PR on the 12x768 'group' backbone, decoder = oracle/ctc_decode.py
    torchaudio.models.decoder.ctc_decoder = ctc_decode.ctc_decoder
    cfg_b = hf_config("base")
    dir_b = os.path.join(tmp, "base")
    save_backbone(cfg_b, 1, dir_b)
    prb = ref_pr.Wav2Vec2_PR(cfg_b, None, dir_b, VOCAB)
    hw, hb = W.linear_params(104, 46, 768)
    with torch.no_grad():
        prb.pr_head.weight.copy_(hw * 8); prb.pr_head.bias.copy_(hb)   # x8: decisive logits -> well-separated argmax
    lens3 = [32000, 27000]
    wav3 = W.waveforms(2, 32000, lens3, seed=8234)
    e = prb.get_embeddings(wav3, torch.tensor(lens3))
    out["api_emb_features"] = e["features_hidden"].numpy()[:, ::16, ::4].copy()
    out["api_emb_last"] = e["last_transf_hidden"].numpy()[:, ::16, ::4].copy()
    out["api_emb_logits"] = e["phoneme_logits"].astype(np.float32)
    out["api_emb_frame_lens"] = e["frame_seq_lens"]
    for b in range(2):
        out[f"api_emb_seq{b}"] = np.asarray(e["phn_pred_seq_idx"][b], dtype=np.int64)
    w1 = W.waveforms(1, 32000, None, seed=9234)
    p = prb.pred_phn_seq(w1, VOCAB)
    out["api_seq_idx"] = np.asarray(p["phn_seq_idx"], dtype=np.int64)
    out["api_seq_ipa"] = np.asarray(p["phn_seq_ipa"])
    d = prb.predict_phonemes_durations(w1, VOCAB)
    out["api_dur_idx"] = np.asarray(d["phn_seq_idx"], dtype=np.int64)
    out["api_dur"] = np.asarray(d["phn_seq_dur"], dtype=np.float64)
    # get_embeddings_grad: values + the gradient of a scalar of its three logit outputs w.r.t. two parameters
    prb.eval()
    eg = prb.get_embeddings_grad(wav3, torch.tensor(lens3), VOCAB, 4, 9)
    cw = torch.from_numpy(np.random.Generator(np.random.PCG64(61)).standard_normal((3, 2, 99, 46), dtype=np.float32))
    s = (eg["phoneme_logits_last"] * cw[0]).sum() + (eg["phoneme_logits_inter"] * cw[1]).sum() \
        + (eg["phoneme_logits_latter"] * cw[2]).sum()
    p1 = prb.wav2vec2.encoder.layers[2].feed_forward.output_dense.weight
    p2 = prb.wav2vec2.encoder.layers[7].attention.q_proj.weight
    p3 = prb.pr_head.weight
    g1, g2_, g3 = torch.autograd.grad(s, [p1, p2, p3])
    out["api_grad_logits_inter"] = eg["phoneme_logits_inter"].detach().numpy()[:, ::4].copy()
    out["api_grad_logits_latter"] = eg["phoneme_logits_latter"].detach().numpy()[:, ::4].copy()
    out["api_grad_inter_hidden"] = eg["intermediate_hidden"].detach().numpy()[:, ::16, ::4].copy()
    out["api_grad_norms"] = np.asarray([float(g1.norm()), float(g2_.norm()), float(g3.norm())])
    out["api_grad_g1"] = g1.numpy()[::16, ::64].copy()
    out["api_grad_g2"] = g2_.numpy()[::16, ::16].copy()
    out["api_grad_g3"] = g3.numpy()[:, ::16].copy()
    print("api", out["api_seq_idx"][:8], out["api_dur"][:4], out["api_grad_norms"], f"{time.time() - t0:.0f}s", flush=True)

    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
