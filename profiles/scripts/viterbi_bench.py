"""CTC-Viterbi kernel timing at bench-like shapes (B utterances x T frames x 46 classes, S labels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from aptai_b200 import ops
dev = torch.device("cuda", 0)
for B, T, S in ((190, 399, 59), (120, 399, 40), (75, 999, 100), (75, 999, 200)):
    g = torch.Generator().manual_seed(0)
    lp = torch.log_softmax(torch.randn((B, T, 46), generator=g), -1).to(dev)
    tg = torch.randint(1, 46, (B, S), generator=g, dtype=torch.int32).to(dev)
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.ctc_viterbi(lp, tg, il, tl, blank=0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.ctc_viterbi(lp, tg, il, tl, blank=0)
    b.record(); torch.cuda.synchronize()
    print(f"B={B} T={T} S={S}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
