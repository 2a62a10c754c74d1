"""world_size-2 gloo worker: the data-parallel gradient path (GradBuffer layout, GradReducer bucketing/averaging,
parameter broadcast) on CPU tensors — host logic only, no kernels."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200.train import GradBuffer, GradReducer, broadcast_parameters  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(rank)


class Layer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.k_proj = torch.nn.Linear(8, 8)
        self.v_proj = torch.nn.Linear(8, 8)
        self.q_proj = torch.nn.Linear(8, 8)
        self.ff = torch.nn.Linear(8, 16)


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.embed = torch.nn.Parameter(torch.randn(5))
        self.layers = torch.nn.ModuleList([Layer() for _ in range(6)])
        self.head = torch.nn.Linear(8, 3)


net = Net()
broadcast_parameters(net, 0)
ref = [p.detach().clone() for p in net.parameters()]
gathered = [torch.zeros_like(ref[3]) for _ in range(world)]
dist.all_gather(gathered, ref[3])
assert all(torch.equal(g, gathered[0]) for g in gathered), "weights must be identical after the broadcast"

groups = []
for i in range(6):
    groups.append([f"layers.{i}.{n}.weight" for n in ("q_proj", "k_proj", "v_proj")])
    groups.append([f"layers.{i}.{n}.bias" for n in ("q_proj", "k_proj", "v_proj")])
gb = GradBuffer(list(net.named_parameters()), groups)
# fused q/k/v gradients are adjacent and in q, k, v order
fw = gb.fused(groups[0], (24, 8))
fw[:8] = 1.0; fw[8:16] = 2.0; fw[16:] = 3.0
L0 = net.layers[0]
assert float(L0.q_proj.weight.grad.mean()) == 1.0 and float(L0.k_proj.weight.grad.mean()) == 2.0
assert float(L0.v_proj.weight.grad.mean()) == 3.0
gb.zero()
# every parameter's grad is a view of the flat buffer
for n, p in gb.params:
    assert gb.owns(p), n

red = GradReducer(gb, "layers.", 6, layers_per_bucket=4)
for n, p in gb.params:
    p.grad.fill_(float(rank + 1))
for i in range(5, -1, -1):            # the backward visits the layers in decreasing order
    red.layer_done(i)
red.finish()
expect = sum(range(1, world + 1)) / world
for n, p in gb.params:
    assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), n
# deferred waits (the optimizer-overlap mode): finish() returns with the reductions registered as pending regions that
# cover every parameter exactly once, in launch order (last layers first); wait_pending() completes them
red2 = GradReducer(gb, "layers.", 6, layers_per_bucket=4, defer_wait=True)
for n, p in gb.params:
    p.grad.fill_(float(rank + 1))
for i in range(5, -1, -1):
    red2.layer_done(i)
red2.finish()
assert len(gb.pending) == 4, len(gb.pending)          # layers 4-5, layers 0-3, the head of the buffer, its tail
covered = sorted((lo, hi) for _, lo, hi in gb.pending)
assert covered[0][0] == 0 and covered[-1][1] == gb.numel and all(a[1] <= b[0] for a, b in zip(covered, covered[1:]))
for n, p in gb.params:
    lo = gb.offsets[n]
    assert any(a <= lo and lo + p.numel() <= b for a, b in covered), n      # whole tensors per region
assert gb.pending[0][1] == red2.spans[4][0]           # the last layers' bucket was launched first
gb.wait_pending()
assert gb.pending == []
for n, p in gb.params:
    assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), n
# the plain (non-overlapped) reduction gives the same result
for n, p in gb.params:
    p.grad.fill_(float(rank + 1))
for w in gb.allreduce(bucket_bytes=1024, async_op=True):
    w.wait()
for n, p in gb.params:
    assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), n
dist.barrier()
if rank == 0:
    print("GLOO_DP_OK")
dist.destroy_process_group()
