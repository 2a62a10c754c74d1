"""TC vs SIMT conv-0 against an fp64 reference: where do they differ by more than one fp16 rounding step?"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from aptai_b200 import ops
dev = torch.device("cuda", 0)
def rnd(shape, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)
B, L = 3, 16000
wav = rnd((B, L), 0.1, 11); w = rnd((512, 10), (2.0 / 10) ** 0.5, 12); bs = rnd((512,), 0.2, 13)
gam = 1 + rnd((512,), 0.1, 14); bet = rnd((512,), 0.1, 15)
ref = F.conv1d(wav.double()[:, None], w.double()[:, None], bs.double(), stride=5)
ref = F.layer_norm(ref.transpose(1, 2), (512,), gam.double(), bet.double(), 1e-5)
pre = ref.clone()
ref = F.gelu(ref)
ops.CONV0_TC = 1
y_tc = ops.conv0(wav, w, bs, gam, bet, 1, out_dtype=torch.float16).double()
ops.CONV0_TC = 0
y_si = ops.conv0(wav, w, bs, gam, bet, 1, out_dtype=torch.float16).double()
ulp = 2.0 ** -10 * ref.abs().clamp_min(2.0 ** -14)
for name, y in (("tc", y_tc), ("simt", y_si)):
    e = (y - ref).abs() / ulp
    print(name, "max err in ulps", e.max().item(), "mean", e.mean().item(), "count > 1", (e > 1).sum().item())
d = (y_tc - y_si).abs()
idx = (d > 1.01 * 2.0 ** -10 * y_si.abs().clamp_min(2.0 ** -14)).nonzero()
print("flagged", idx.shape[0])
for i in idx[:12].tolist():
    b, t, c = i
    print(i, "pre", pre[b, t, c].item(), "ref", ref[b, t, c].item(), "tc", y_tc[b, t, c].item(), "simt", y_si[b, t, c].item())
