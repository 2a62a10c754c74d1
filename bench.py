#!/usr/bin/env python
"""bench.py — audio-seconds per second of the APTAI forward + alignment hot path (BASELINE.json metric).

Workload (BASELINE config 5): 4096 synthetic utterances, duration U[2, 20] s at 16 kHz, XLS-R-sized 24x1024 'layer'
backbone with random-init weights, APTAI inference (encoder -> TV head + low-pass, phoneme head + argmax) followed
by log-softmax + CTC-Viterbi forced alignment against known synthetic phoneme sequences.  One step = one pass over
all utterances, length-bucketed (padding semantics of the reference: pad to the batch maximum).
Every rank runs the full workload on its own GPU (utterance-sharded path, no collective): weak scaling.

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the reference path (torch-CPU
                                                           restatement of transformers.Wav2Vec2Model + APTAI heads)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from aptai_b200 import sweep  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402

NO_REG = dict(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
              final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)
VOCAB = {"(blank)": 0, "(...)": 1, **{f"p{i}": i for i in range(2, 46)}}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(cfg, n_utts, seed, max_rows):
    lengths = sweep.synth_durations(n_utts, 2.0, 20.0, seed=seed)
    batches = sweep.make_batches(cfg, lengths, bucket_width=32, max_rows=max_rows)
    return lengths, batches


def synth_batch_host(lengths, batch, seed):
    """Pinned host buffers of one batch: waveform fp32 [B,L] (0.1*N(0,1), zero beyond each length), lengths, known
    phoneme sequences int32 [B,59] (ids 1..45, length U{10..59} capped by the frame count)."""
    B, L = len(batch.indices), batch.samples
    g = torch.Generator().manual_seed(seed)
    wav = torch.empty((B, L), dtype=torch.float32).pin_memory()
    wav.normal_(0.0, 0.1, generator=g)
    lens = torch.tensor([lengths[i] for i in batch.indices], dtype=torch.int64)
    for b in range(B):
        wav[b, int(lens[b]):] = 0.0
    rng = np.random.Generator(np.random.PCG64(seed))
    tg = np.zeros((B, 59), dtype=np.int32)
    tl = np.zeros((B,), dtype=np.int32)
    for b in range(B):
        n = int(rng.integers(10, 60))
        tl[b] = n
        seq = rng.integers(1, 46, size=n)
        for j in range(1, n):                       # no adjacent repeats: always feasible for T >= 99 frames
            if seq[j] == seq[j - 1]:
                seq[j] = 1 + (seq[j] % 45)
        tg[b, :n] = seq
    return wav, lens.pin_memory(), torch.from_numpy(tg).pin_memory(), torch.from_numpy(tl).pin_memory()


def make_model(cfg, dev):
    from aptai_b200 import APTAI
    from aptai_b200.backbone import register_in_memory_checkpoint
    from aptai_b200.synth import backbone_state_dict, linear_params
    name = register_in_memory_checkpoint("mem://bench", backbone_state_dict(cfg, 0))
    m = APTAI(dev, VOCAB, name, cfg, None, phn_drop=0.0, tv_drop=0.0)
    tvw, tvb = linear_params(101, 9, cfg.hidden_size)
    pw, pb = linear_params(102, 46, cfg.hidden_size)
    with torch.no_grad():
        m.tv_head[2].weight.copy_(tvw); m.tv_head[2].bias.copy_(tvb)
        m.phn_head[2].weight.copy_(pw); m.phn_head[2].bias.copy_(pb)
    return m.to(dev).eval()


def hot_path(model, wav, lens, tg, tl):
    """One batch through the public API: APTAI.predict = encoder + heads + low-pass + argmax + forced alignment."""
    r = model.predict(wav, lens, phn_targets=tg, phn_target_lens=tl)
    return r["tvs_pred"], r["phn_fc_pred"], r["align_paths"]


class GemmTimer:
    """CUDA-event bracket around every launch of the GEMM family + its algorithmic FLOPs (roofline leg)."""

    def __init__(self):
        self.ev, self.flops = [], 0.0

    def hook(self, a):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        K = a.taps * a.kb_per_tap * 64
        kv = K if a.a_col_per_nblk == 0 else a.taps * a.a_col_per_nblk      # grouped conv: real group width
        self.flops += 2.0 * a.segs * a.rows_per_seg * a.N * kv
        self.ev.append((e0, e1))
        return e1.record

    def total_ms(self):
        return sum(e0.elapsed_time(e1) for e0, e1 in self.ev)


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the aptai_b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...") goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from aptai_b200 import lib, ops
    cfg = W2V2Config.large(**NO_REG)
    lengths, batches = build_workload(cfg, args.utterances, seed=rank, max_rows=args.max_rows)
    audio_s = sum(lengths) / 16000.0
    model = make_model(cfg, dev)
    host = [synth_batch_host(lengths, b, 1000 * rank + i) for i, b in enumerate(batches)]
    devb = [tuple(t.to(dev, non_blocking=True) for t in h) for h in host]
    torch.cuda.synchronize()
    padded_frames = sum(len(b.indices) * b.frames for b in batches)
    valid_frames = sum(cfg.conv_out_length(l) for l in lengths)

    def step_resident():
        for (wav, lens, tg, tl) in devb:
            hot_path(model, wav, lens, tg, tl)

    def step_e2e(sink):
        for (wav, lens, tg, tl) in host:
            w, l, g, t = (x.to(dev, non_blocking=True) for x in (wav, lens, tg, tl))
            tv, pred, paths = hot_path(model, w, l, g, t)
            sink.append((tv.to("cpu", non_blocking=True), pred.to("cpu", non_blocking=True),
                         paths.to("cpu", non_blocking=True)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = lib.launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    timer = GemmTimer()
    ops.set_gemm_hook(timer.hook)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ops.set_gemm_hook(None)
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = lib.launch_count() - launches0
    gemm_ms = timer.total_ms()
    gemm_flops = timer.flops
    # end-to-end: pinned host -> device copies and device -> host results inside the timed region
    sink = []
    step_e2e(sink)
    barrier()
    sink.clear()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        sink.clear()
        step_e2e(sink)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(e1) if False else t0.elapsed_time(t1)
    h2d = sum(sum(t.numel() * t.element_size() for t in h) for h in host)
    d2h = sum(len(b.indices) * b.frames * (9 * 4 + 8 + 4) for b in batches)
    tms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max, ms_e2e_max = float(tms[0]), float(tms[1])
    if rank == 0:
        peak_tf, peak_hbm, how = peaks()
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": "audio-sec/sec APTAI fwd+align", "value": world * audio_s * args.steps / (ms_max * 1e-3),
            "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"config5: {args.utterances} utterances U[2,20]s 16kHz per GPU, XLS-R-sized 24x1024 "
                                   "'layer' backbone, APTAI heads+low-pass+argmax + log-softmax + CTC-Viterbi alignment",
                       "batches": len(batches), "bucket_width_frames": 32, "max_rows": args.max_rows,
                       "padded_over_valid_frames": padded_frames / valid_frames,
                       "l2": "inputs larger than L2 (each batch's activations >> 126 MB; weights 631 MB)",
                       "audio_s_per_gpu_per_step": audio_s, "parallelism": f"utterance-sharded x{world}, no collective"},
            "e2e": {"value": world * audio_s * args.steps / (ms_e2e_max * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all launches of the GEMM family)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "peak_source": how,
                         "gemm_share_of_step": gemm_ms / ms, "gemm_launches": len(timer.ev)},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, lengths, batches, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cfg, lengths, batches, budget_s=15.0):
    """The oracle port (torch-CPU fp32 restatement of the reference path) on a bounded sample of the workload."""
    from oracle import ctc as octc
    from oracle import heads as oh
    from oracle import w2v2 as ow
    from oracle.weights import backbone_state_dict, linear_params
    torch.set_num_threads(os.cpu_count())
    sd = backbone_state_dict(cfg, 0)
    tvw, tvb = linear_params(101, 9, cfg.hidden_size)
    pw, pb = linear_params(102, 46, cfg.hidden_size)
    taps = oh.lowpass_taps()
    b = batches[len(batches) // 2]                         # median-length bucket
    idx = b.indices[:4]
    lens = [lengths[i] for i in idx]
    L = max(lens)
    g = torch.Generator().manual_seed(1)
    wav = torch.empty((len(idx), L)).normal_(0.0, 0.1, generator=g)
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    rng = np.random.Generator(np.random.PCG64(1))
    tgs = [rng.permutation(45)[:20] + 1 for _ in idx]

    def once():
        with torch.no_grad():
            h = ow.forward(sd, cfg, wav, lens)[-1]
            _, tv, logits = oh.aptai_heads(h, tvw, tvb, pw, pb, taps)
            lp = torch.log_softmax(logits, -1).numpy()
            for i, n in enumerate(lens):
                octc.viterbi_align(lp[i, : cfg.conv_out_length(n)], tgs[i], blank=0)

    once()
    t0 = time.perf_counter()
    n = 0
    while True:
        once()
        n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n * sum(lens) / 16000.0 / dt, "unit": "audio-s/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"{n} passes over {len(idx)} utterances of the median bucket "
                                      f"({sum(lens) / 16000.0:.1f} audio-s per pass, fp32, {dt:.1f} s of CPU work)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = W2V2Config.large(**NO_REG)
    lengths, batches = build_workload(cfg, args.utterances, seed=0, max_rows=args.max_rows)
    world = int(os.environ.get("WORLD_SIZE", 1))
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(cfg, lengths, batches, budget_s=per_step)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "audio-sec/sec APTAI fwd+align", "value": v, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config5 sample: {base['sample']}", "parallelism": "host CPU threads"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utterances", type=int, default=4096)
    ap.add_argument("--max-rows", type=int, default=75776)   # 4 x (74 CTA pairs x 256 rows)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
