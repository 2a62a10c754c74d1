#!/usr/bin/env python
"""Training-step timing (BASELINE configs 2 and 4), one process per GPU:

  config 4: APTAI training step (articulatory regression + phoneme CE), 24x1024 backbone, batch 32 per GPU of 4-8 s
            utterances, frozen conv encoder, fused Adam, data-parallel gradient all-reduce overlapped with the backward
  config 2: Wav2Vec2_PR CTC forward + backward, batch 16 x 8 s
  config 3: Force_APTAI training step (frozen recogniser forward in eval mode, tail forward + backward + Adam), batch
            64 x 8 s with known phoneme sequences of 10-59 phonemes

  python profiles/train_bench.py --config 4 --steps 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
      profiles/train_bench.py --config 4 --steps 5

Prints one JSON line per run (rank 0): ms per step (CUDA events, max over ranks), audio-s/s over all ranks,
forward / backward / optimizer split of rank 0, and model FLOP/s (3 x forward FLOPs of the trainable part + 1 x conv).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import APTAI, Force_APTAI, Wav2Vec2_PR, lib  # noqa: E402
from aptai_b200.backbone import register_in_memory_checkpoint  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402
from aptai_b200.synth import backbone_state_dict, linear_params, phoneme_sequences  # noqa: E402
from aptai_b200.train import FusedAdam  # noqa: E402

NO_REG = dict(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
              final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)
VOCAB = {"(blank)": 0, "(...)": 1, **{f"p{i}": i for i in range(2, 46)}}


def flops_train(cfg, lens, mult=3.0):
    H, F, N = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    conv = rest = 0.0
    for L in lens:
        t, cin = L, 1
        for k, s, c in zip(cfg.conv_kernel, cfg.conv_stride, cfg.conv_dim):
            t = (t - k) // s + 1
            conv += 2.0 * c * cin * k * t
            cin = c
        T = t
        rest += 2.0 * T * cin * H + 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
        rest += N * T * (8.0 * H * H + 4.0 * H * F) + N * 4.0 * T * T * H + 2.0 * T * H * 55
    return conv + mult * rest


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--model", default="large", choices=["large", "base"])
    ap.add_argument("--unfrozen", action="store_true",
                    help="train the conv feature encoder too (the reference's default for the recogniser)")
    ap.add_argument("--regularised", action="store_true",
                    help="XLS-R's published training regularisers: hidden/attention dropout 0.1, layerdrop 0.1, "
                         "SpecAugment mask_time_prob 0.075, head dropouts 0.1")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    reg = dict(NO_REG)
    if args.regularised:
        reg.update(hidden_dropout=0.1, attention_dropout=0.1, activation_dropout=0.0, feat_proj_dropout=0.0,
                   final_dropout=0.1, layerdrop=0.1, apply_spec_augment=True, mask_time_prob=0.075, mask_time_length=10,
                   mask_time_min_masks=2)
    cfg = (W2V2Config.large if args.model == "large" else W2V2Config.base)(**reg)
    hd = 0.1 if args.regularised else 0.0
    H = cfg.hidden_size
    name = register_in_memory_checkpoint("mem://train-bench", backbone_state_dict(cfg, 0))
    B = args.batch or {4: 32, 2: 16, 3: 64}[args.config]
    L = 128000
    rng = np.random.Generator(np.random.PCG64(21 + rank))
    g = torch.Generator().manual_seed(1234 + rank)
    if args.config == 4:
        model = APTAI(dev, VOCAB, name, cfg, None, phn_drop=hd, tv_drop=hd)
        tvw, tvb = linear_params(101, 9, H)
        pw, pb = linear_params(102, 46, H)
        with torch.no_grad():
            model.tv_head[2].weight.copy_(tvw); model.tv_head[2].bias.copy_(tvb)
            model.phn_head[2].weight.copy_(pw); model.phn_head[2].bias.copy_(pb)
        lens = rng.integers(64000, 128001, size=B)
        lens[0] = L
    elif args.config == 3:
        model = Force_APTAI("unused", dev, VOCAB, w2v2_pr=Wav2Vec2_PR(cfg, None, name, VOCAB))
        lens = np.full((B,), L)
    else:
        model = Wav2Vec2_PR(cfg, None, name, VOCAB)
        if not args.unfrozen:
            model.wav2vec2.freeze_feature_encoder()
        lens = np.full((B,), L)
    model = model.to(dev).train()
    wav = torch.empty((B, L)).normal_(0.0, 0.1, generator=g)
    for b in range(B):
        wav[b, int(lens[b]):] = 0
    flen = [cfg.conv_out_length(int(n)) for n in lens]
    T = cfg.conv_out_length(L)
    kwargs = {}
    if args.config == 3:
        tvt = torch.from_numpy(rng.standard_normal((B, T, 9), dtype=np.float32)).to(dev)
        seqs, sl = phoneme_sequences(B, 10, 59, 1, 45, seed=11 + rank, pad=0)
        kwargs["phn_seqs"] = [seqs[b, : int(sl[b])].numpy().astype(np.int64) for b in range(B)]
        batch = (0, wav.to(dev), torch.as_tensor(lens).to(dev), None, None,
                 *[tvt[:, :, i].contiguous() for i in range(9)])
    elif args.config == 4:
        phn = np.zeros((B, T), dtype=np.int64)
        tvt = np.full((B, T, 9), -100.0, dtype=np.float32)
        for b in range(B):
            phn[b, : flen[b]] = rng.integers(1, 46, size=flen[b])
            tvt[b, : flen[b]] = rng.standard_normal((flen[b], 9), dtype=np.float32)
        batch = (0, wav.to(dev), torch.as_tensor(lens).to(dev), torch.from_numpy(phn).to(dev),
                 *[torch.from_numpy(tvt[:, :, i]).to(dev) for i in range(9)])
    else:
        labels, _ = phoneme_sequences(B, 10, 59, 2, 45, seed=7 + rank, pad=-100)
        batch = (wav.to(dev), torch.as_tensor(lens).to(dev), labels.to(dev))
    opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-5)
    if world > 1:
        model.enable_data_parallel()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    losses = []

    def step(times=None):
        e = [ev() for _ in range(4)]
        e[0].record()
        opt.zero_grad()
        out = model(*batch, **kwargs)
        e[1].record()
        out["loss"].backward()
        e[2].record()
        opt.step()
        e[3].record()
        losses.append(out["loss"].detach())
        if times is not None:
            times.append(e)

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.launch_count()
    t0, t1 = ev(), ev()
    times = []
    t0.record()
    for _ in range(args.steps):
        step(times)
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        fwd = np.mean([e[0].elapsed_time(e[1]) for e in times])
        bwd = np.mean([e[1].elapsed_time(e[2]) for e in times])
        optm = np.mean([e[2].elapsed_time(e[3]) for e in times])
        audio_s = float(sum(lens)) / 16000.0
        fl = flops_train(cfg, [int(n) for n in lens], 1.0 if args.config == 3 else 3.0)   # config 3: frozen backbone
        print(json.dumps({
            "workload": f"config{args.config}: " + {4: "APTAI training step", 2: "Wav2Vec2_PR CTC fwd+bwd+Adam",
                                                    3: "Force_APTAI training step (frozen recogniser, trained tail)"}[args.config] +
                        f", {args.model} backbone, batch {B}/GPU, max 8 s, "
                        f"{'conv encoder trained' if args.unfrozen and args.config == 2 else 'frozen conv encoder'}, fused Adam"
                        + (", dropout 0.1 (hidden/attention/heads) + LayerDrop 0.1 + SpecAugment 0.075" if args.regularised else "")
                        + (", DP all-reduce overlapped with backward" if world > 1 else ""),
            "n_gpus": world, "ms_per_step": ms, "audio_s_per_s": world * audio_s / (ms * 1e-3),
            "fwd_ms": fwd, "bwd_ms": bwd, "opt_ms": optm, "model_tflops_per_gpu": fl / (ms * 1e-3) / 1e12,
            "launches_per_step": (lib.launch_count() - l0) / args.steps,
            "loss_first_last": [float(losses[0]), float(losses[-1])],
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
