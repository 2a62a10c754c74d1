"""One APTAI.predict pass in the accuracy mode (precision="f32x3") for ncu launch lists: PROFILE_B x PROFILE_L."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from aptai_b200.config import W2V2Config
B = int(os.environ.get("PROFILE_B", 32)); L = int(os.environ.get("PROFILE_L", 128000))
dev = torch.device("cuda:0")
model = bench.make_model(W2V2Config.large(**bench.NO_REG), dev)
model.set_precision("f32x3")
g = torch.Generator().manual_seed(0)
wav = torch.empty((B, L)).normal_(0.0, 0.1, generator=g).to(dev)
lens = torch.full((B,), L, dtype=torch.int64, device=dev)
tg = (torch.arange(59, dtype=torch.int32)[None] % 45 + 1).repeat(B, 1).to(dev)
tl = torch.full((B,), 40, dtype=torch.int32, device=dev)
for _ in range(2):
    r = model.predict(wav, lens, phn_targets=tg, phn_target_lens=tl)
torch.cuda.synchronize()
print("ok")
