import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import sweep  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = W2V2Config.large()
lengths = sweep.synth_durations(256, 2.0, 20.0, seed=0)
batches = sweep.make_batches(cfg, lengths, 32, 49152)
mine = sweep.shard_lpt(batches, world)[rank]
ids = torch.zeros(256, dtype=torch.int64)
for b in mine:
    ids[b.indices] += 1
dist.all_reduce(ids)
assert bool((ids == 1).all()), "shards must partition the utterances"
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == float(world)
load = torch.tensor([sum(b.flops for b in mine)], dtype=torch.float64)
loads = [torch.zeros_like(load) for _ in range(world)]
dist.all_gather(loads, load)
assert max(loads).item() / min(loads).item() < 1.15
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
