"""One fused-LayerNorm conv GEMM (conv layer 1 shape: k = 3, stride 2, 512 -> 512 channels, fp16) for ncu captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import ops
dev = torch.device("cuda:0")
B, T_in, C = int(os.environ.get("PROFILE_B", 32)), 25599, 512
x = ops.alloc_rows_bf16(B, T_in, C, dev, dtype=torch.float16)
x.normal_(0.0, 1.0)
w = (torch.randn((512, 3 * C), device=dev) * 0.03).half()
b = torch.randn((512,), device=dev) * 0.1
g = 1 + 0.1 * torch.randn((512,), device=dev)
e = 0.1 * torch.randn((512,), device=dev)
for _ in range(3):
    y = ops.conv_igemm(x, w, b, 3, 2, ln_gamma=g, ln_beta=e, act=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = ops.conv_igemm(x, w, b, 3, 2, ln_gamma=g, ln_beta=e, act=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
T_out = (T_in - 3) // 2 + 1
print(f"conv1-shaped fused-LN GEMM B={B}: {ms:.3f} ms, {2.0 * B * T_out * 512 * 1536 / ms / 1e9:.1f} TFLOP/s")
