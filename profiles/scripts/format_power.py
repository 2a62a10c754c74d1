"""Why is precision="fp16" slower than bf16 on the same kernels?  Sustained runs (seconds, so the power cap acts) of the
FFN1 GEMM and of the whole 24-layer step of one batch in both operand formats, with nvidia-smi clock / power samples
taken during each run.  Usage: python profiles/scripts/format_power.py [seconds]"""
import os, statistics, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from aptai_b200 import ops


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for l in self.p.stdout:
            try:
                a, b = l.split(",")
                self.rows.append((float(a), float(b)))
            except Exception:
                pass

    def stop(self, skip=5):
        self.p.terminate()
        r = self.rows[skip:] or self.rows
        return statistics.median(x[0] for x in r), statistics.median(x[1] for x in r), len(r)


def sustained(fn, seconds):
    fn(); torch.cuda.synchronize()
    smi = Smi()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    mhz, watts, k = smi.stop()
    return e0.elapsed_time(e1) / n, mhz, watts, k


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    dev = torch.device("cuda:0")
    M, K, N = 49152, 1024, 4096
    g = torch.Generator(device=dev).manual_seed(0)
    a32 = torch.randn((M, K), device=dev, generator=g)
    w32 = torch.randn((N, K), device=dev, generator=g) * 0.03
    b = torch.zeros((N,), device=dev)
    for name, scale in (("randn", 1.0), ("zeros", 0.0)):
        for dt in (torch.bfloat16, torch.float16):
            a, w = (a32 * scale).to(dt), (w32 * scale).to(dt)
            out = torch.empty((M, N), dtype=dt, device=dev)
            ms, mhz, watts, k = sustained(lambda: ops.linear(a, w, b, act=1, out_bf16=out), secs)
            print(f"FFN1+GELU {name:6s} {str(dt)[6:]:9s} {ms * 1e3:7.1f} us  {2 * M * K * N / ms / 1e9:7.1f} TFLOP/s  "
                  f"{mhz:6.0f} MHz  {watts:6.0f} W  ({k} samples)", flush=True)
    # whole inference step of one batch (B = 120 x 8 s) in both modes
    from aptai_b200 import APTAI
    from aptai_b200.backbone import register_in_memory_checkpoint
    from aptai_b200.config import W2V2Config
    from aptai_b200.synth import backbone_state_dict, waveforms
    cfg = W2V2Config.large(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, final_dropout=0.0,
                           layerdrop=0.0, apply_spec_augment=False)
    name = register_in_memory_checkpoint("mem://fp", backbone_state_dict(cfg, 0))
    vocab = {"(blank)": 0, "(...)": 1, **{f"p{i}": i for i in range(2, 46)}}
    m = APTAI(dev, vocab, name, cfg, None, phn_drop=0.0, tv_drop=0.0).to(dev).eval()
    B, L = 120, 128000
    wav = waveforms(B, L, None, seed=3).to(dev)
    lens = torch.full((B,), L, device=dev)
    for mode in ("bf16", "fp16", "bf16", "fp16"):
        m.set_precision(mode)
        ms, mhz, watts, k = sustained(lambda: m.predict(wav, lens), secs)
        print(f"APTAI.predict B=120x8s {mode}: {ms:7.2f} ms  {B * 8 / ms * 1e3:8.0f} audio-s/s  {mhz:6.0f} MHz  {watts:6.0f} W",
              flush=True)


if __name__ == "__main__":
    main()
