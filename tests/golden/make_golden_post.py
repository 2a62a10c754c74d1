"""Golden vectors for the callers' data formats, produced by the REFERENCE's own functions:
utility.py (phn_frames2dur, phn_frame_id2phn, tvs_metric_rmse, tvs_metric_ppc, get_stats, evaluate_overlap),
train/train_aptai.py `_collate_fn`, data/dataset_hprc.py `interpolate_signal`, and torchaudio.functional.resample
(the call data/dataset_hprc.py:68-72 makes).  Functions living in modules whose imports need corpora tooling are
extracted with `ast` from the read-only reference file and executed here; nothing is copied into the repo.

Run once in the build container:   python tests/golden/make_golden_post.py
"""
import ast
import os
import sys
import types

import numpy as np
import scipy.interpolate
import torch
import transformers  # noqa: F401
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
import utility as ref_util  # noqa: E402


def extract(path, name, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    rng = np.random.Generator(np.random.PCG64(77))
    out = {}
    # ---- segments
    frames = np.repeat(rng.integers(1, 46, size=40), rng.integers(1, 9, size=40)).astype(np.int64)
    out["p_frames"] = frames
    dur = ref_util.phn_frames2dur(frames.tolist())
    out["p_dur_start"] = np.asarray([d[0] for d in dur])
    out["p_dur_end"] = np.asarray([d[1] for d in dur])
    out["p_dur_phn"] = np.asarray([d[2] for d in dur], dtype=np.int64)
    out["p_id2phn"] = np.asarray(ref_util.phn_frame_id2phn(frames.tolist()), dtype=np.int64)
    # ---- TV metrics
    gt = rng.standard_normal((173, 9)).astype(np.float32)
    pred = (gt + 0.3 * rng.standard_normal((173, 9))).astype(np.float32)
    out["m_gt"], out["m_pred"] = gt, pred
    rm = ref_util.tvs_metric_rmse(gt, pred)
    pc = ref_util.tvs_metric_ppc(gt, pred)
    out["m_rmse"] = np.asarray([rm[k] for k in rm], dtype=np.float64)
    out["m_pcc"] = np.asarray([pc[k][0] for k in pc], dtype=np.float64)
    # ---- boundary stats
    y = np.sort(np.round(rng.uniform(0, 8, size=37), 2))
    yhat = np.sort(np.round(np.concatenate([y[:30] + rng.choice([-0.03, -0.02, -0.01, 0.0, 0.01, 0.02, 0.03], 30),
                                            rng.uniform(0, 8, 5)]), 2))
    out["b_y"], out["b_yhat"] = y, yhat
    out["b_stats"] = np.asarray(ref_util.get_stats(y, yhat, tolerance=0.02), dtype=np.float64)
    a = [rng.integers(1, 5, size=n).tolist() for n in (50, 77, 3)]
    b = [[(v if rng.random() > 0.3 else 1 + v % 4) for v in s] for s in a]
    out["o_a"] = np.asarray(sum(a, []), dtype=np.int64)
    out["o_b"] = np.asarray(sum(b, []), dtype=np.int64)
    out["o_lens"] = np.asarray([len(s) for s in a])
    out["o_overlap"] = np.asarray([ref_util.evaluate_overlap(a, b)])
    # ---- interpolate_signal
    interp = extract(os.path.join(REF, "data", "dataset_hprc.py"), "interpolate_signal", {"np": np, "scipy": scipy})
    sig = rng.standard_normal((211, 9))
    out["i_sig"] = sig
    out["i_out_97"] = interp(sig, 97)
    out["i_out_400"] = interp(sig, 400)
    # ---- collate
    collate = extract(os.path.join(REF, "train", "train_aptai.py"), "_collate_fn", {"torch": torch})
    names = ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")
    batch = []
    for i, (L, T) in enumerate([(3000, 9), (1234, 3), (4000, 12)]):
        batch.append({"audio": torch.from_numpy(rng.standard_normal(L).astype(np.float32)), "audio_len": L,
                      "phn_frames_49hz": rng.integers(1, 46, size=T).tolist(),
                      "tvs_norm_49hz": {k: rng.standard_normal(T).astype(np.float32) for k in names}})
    col = collate(batch)
    for i, x in enumerate(batch):
        out[f"c_audio{i}"] = x["audio"].numpy()
        out[f"c_phn{i}"] = np.asarray(x["phn_frames_49hz"], dtype=np.int64)
        out[f"c_tv{i}"] = np.stack([x["tvs_norm_49hz"][k] for k in names], -1)
    out["c_audio_inputs"] = col["audio_inputs"].numpy()
    out["c_audio_lengths"] = col["audio_lengths"].numpy()
    out["c_phn_frames"] = col["phn_frames_49hz"].numpy()
    out["c_tvs"] = np.stack([col[k].numpy() for k in names], -1)
    # ---- resample (dataset_hprc.py:68-72)
    wav = (0.1 * rng.standard_normal(22050)).astype(np.float32)
    out["r_wav"] = wav
    out["r_44100"] = torchaudio.functional.resample(waveform=torch.from_numpy(wav)[None], orig_freq=44100,
                                                    new_freq=16_000)[0].numpy()
    out["r_22050"] = torchaudio.functional.resample(waveform=torch.from_numpy(wav[:9999])[None], orig_freq=22050,
                                                    new_freq=16_000)[0].numpy()
    out["r_8000"] = torchaudio.functional.resample(waveform=torch.from_numpy(wav[:4000])[None], orig_freq=8000,
                                                   new_freq=16_000)[0].numpy()
    path = os.path.join(HERE, "golden_post_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
