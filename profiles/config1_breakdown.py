"""Config 1 (one 4 s utterance, 12x768 backbone): where the 2 ms go — GPU time of the replayed CUDA graph vs the
eager launch sequence vs the host-side wrapper (numpy in/out)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aptai_b200 import APTAI  # noqa: E402
from aptai_b200.backbone import register_in_memory_checkpoint  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402
from aptai_b200.synth import backbone_state_dict, waveforms  # noqa: E402

dev = torch.device("cuda:0")
cfg = W2V2Config.base(**bench.NO_REG)
register_in_memory_checkpoint("mem://c1", backbone_state_dict(cfg, 1))
m = APTAI(dev, bench.VOCAB, "mem://c1", cfg, None, 0.0, 0.0).to(dev).eval()
wav = waveforms(1, 64000, None, seed=1234)
w = wav.to(dev)
ln = torch.tensor([64000], device=dev)


def ev(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def wall(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


with torch.no_grad():
    m.use_cuda_graphs = True
    graphed = ev(lambda: m._single_graphed(w, ln))
    m.use_cuda_graphs = False
    eager = ev(lambda: m._single_graphed(w, ln))
    eager_wall = wall(lambda: m.get_aptai_output(wav[0].numpy()))
    m.use_cuda_graphs = True
    graph_wall = wall(lambda: m.get_aptai_output(wav[0].numpy()))
print(json.dumps({"config": 1, "gpu_ms_graph_replay": graphed, "gpu_ms_eager_launches": eager,
                  "wall_ms_get_aptai_output_eager": eager_wall, "wall_ms_get_aptai_output_graph": graph_wall,
                  "audio_s_per_s_graph": 4.0 / (graph_wall * 1e-3)}))
