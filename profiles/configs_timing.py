"""Wall-clock / CUDA-event timings of BASELINE.json configs 1-3 at their full sizes on one B200 (inference halves; the
training steps are timed by profiles/train_bench.py) with size-independent property checks.  Prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from aptai_b200 import APTAI, Force_APTAI, Wav2Vec2_PR, ops  # noqa: E402
from aptai_b200.backbone import register_in_memory_checkpoint  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402
from aptai_b200.synth import backbone_state_dict, linear_params, phoneme_sequences, waveforms  # noqa: E402

dev = torch.device("cuda:0")


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def wall_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


# ---- config 1: one 4 s utterance, base-sized backbone, phoneme + articulatory heads (wall clock, numpy out)
cfg_b = W2V2Config.base(**bench.NO_REG)
register_in_memory_checkpoint("mem://c1", backbone_state_dict(cfg_b, 1))
m1 = APTAI(dev, bench.VOCAB, "mem://c1", cfg_b, None, 0.0, 0.0)
m1 = m1.to(dev).eval()
wav1 = waveforms(1, 64000, None, seed=1234)[0].numpy()
ms = wall_time(lambda: m1.get_aptai_output(wav1))
print(json.dumps({"config": 1, "what": "APTAI.get_aptai_output, 4 s, 12x768 'group' backbone, host numpy in/out",
                  "ms": ms, "audio_s_per_s": 4.0 / (ms * 1e-3)}))
del m1

# ---- config 2 (forward half + fused CTC loss and d loss/d logits): B=16 x 8 s, 24x1024
cfg_l = W2V2Config.large(**bench.NO_REG)
sd_l = backbone_state_dict(cfg_l, 0)
register_in_memory_checkpoint("mem://c2", sd_l)
pr = Wav2Vec2_PR(cfg_l, None, "mem://c2", bench.VOCAB)
hw, hb = linear_params(103, 46, 1024)
with torch.no_grad():
    pr.pr_head.weight.copy_(hw); pr.pr_head.bias.copy_(hb)
pr = pr.to(dev).eval()
lens = [128000 - 8000 * (i % 4) for i in range(16)]
wav = waveforms(16, 128000, lens, seed=7).to(dev)
labels, _ = phoneme_sequences(16, 10, 59, 2, 45, seed=7, pad=-100)
lt = torch.tensor(lens, device=dev)
r = pr(wav, lt, labels.to(dev), want_grad=True)
g = r["grad_logits"]
assert torch.isfinite(r["loss"]) and float(g.sum(-1).abs().max()) < 1e-4
assert float(g[1, 374:].abs().max()) == 0.0                 # frames beyond the 7.5 s utterance: exactly zero
ms = ev_time(lambda: pr(wav, lt, labels.to(dev), want_grad=True))
print(json.dumps({"config": 2, "what": "Wav2Vec2_PR.forward B=16x8s 24x1024 (+CTC loss and d loss/d logits; full training step: train_bench.py)",
                  "ms": ms, "audio_s_per_s": sum(lens) / 16000 / (ms * 1e-3), "loss": float(r["loss"])}))

# ---- config 3: Force_APTAI forced alignment, B=64 x 8 s, known phoneme sequences
fa = Force_APTAI("unused", dev, bench.VOCAB, w2v2_pr=pr).to(dev).eval()
B = 64
lens3 = [128000 - 4000 * (i % 8) for i in range(B)]
wav3 = waveforms(B, 128000, lens3, seed=11).to(dev)
seqs, sl = phoneme_sequences(B, 10, 59, 1, 45, seed=11, pad=0)
seq_list = [seqs[b, : int(sl[b])].numpy() for b in range(B)]
lt3 = torch.tensor(lens3, device=dev)
paths, scores, status = fa.forced_align(wav3, lt3, seq_list)
p = paths.cpu().numpy()
flen = [cfg_l.conv_out_length(n) for n in lens3]
for b in range(B):
    assert int(status[b]) == 0
    pb = p[b, : flen[b]]
    col = [int(x) for i, x in enumerate(pb) if x != 0 and (i == 0 or pb[i - 1] != x)]
    # collapsing the frame path (merge repeats, drop blanks) must give back the known sequence, up to repeats that
    # CTC separates by a blank
    dec = []
    prev = -1
    for x in pb:
        if x != prev and x != 0:
            dec.append(int(x))
        prev = x
    assert dec == [int(v) for v in seq_list[b]], b
    assert (p[b, flen[b]:] == -1).all()
ms = ev_time(lambda: fa.forced_align(wav3, lt3, seq_list))
_, _, logits = pr._logits(wav3, lt3)
lp = ops.softmax_rows(logits.contiguous(), log=True)
ms_v = ev_time(lambda: fa.forced_align(wav3, lt3, seq_list, log_probs=lp))
print(json.dumps({"config": 3, "what": "Force_APTAI.forced_align B=64x8s (encoder + log-softmax + CTC Viterbi); path collapses to the known sequence for all 64",
                  "ms": ms, "viterbi_only_ms": ms_v, "audio_s_per_s": sum(lens3) / 16000 / (ms * 1e-3)}))
t3 = fa._trunk(wav3[:8], lt3[:8], seq_list[:8])
print(json.dumps({"config": 3, "what": "Force_APTAI cross-attention alignment matrix B=8 (batch > 1 works; the reference raises NameError)",
                  "att_shape": list(t3["att"].shape)}))
# the model's own forward at the same size: recogniser + cross-attention alignment matrix + BiLSTM + low-pass + losses
tvt3 = torch.randn((B, 399, 9), generator=torch.Generator().manual_seed(5)).to(dev)
tv_cols = [tvt3[:, :, i].contiguous() for i in range(9)]
with torch.no_grad():
    ms_f = ev_time(lambda: fa(0, wav3, lt3, None, None, *tv_cols, phn_seqs=seq_list))
print(json.dumps({"config": 3, "what": "Force_APTAI.forward B=64x8s eval (recogniser + cross-attention + BiLSTM + low-pass + MSE / forward-sum losses + frame phonemes)",
                  "ms": ms_f, "audio_s_per_s": sum(lens3) / 16000 / (ms_f * 1e-3)}))
