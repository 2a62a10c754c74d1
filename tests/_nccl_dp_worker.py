"""world_size-N NCCL worker (one process per GPU): data-parallel APTAI training step.  Checks that the overlapped,
layer-bucketed all-reduce averages the gradients (== mean over ranks of the single-rank gradients of the same
batches) and that the ranks' weights stay identical after the fused Adam step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aptai_b200 import APTAI  # noqa: E402
from aptai_b200.backbone import register_in_memory_checkpoint  # noqa: E402
from aptai_b200.config import W2V2Config  # noqa: E402
from aptai_b200.synth import backbone_state_dict, waveforms  # noqa: E402
from aptai_b200.train import FusedAdam  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = W2V2Config.large(num_hidden_layers=6, hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0,
                       feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False)
vocab = {"(blank)": 0, "(...)": 1, **{f"p{i}": i for i in range(2, 46)}}
name = register_in_memory_checkpoint("mem://dp", backbone_state_dict(cfg, rank))     # different init per rank ...
m = APTAI(dev, vocab, name, cfg, None, phn_drop=0.0, tv_drop=0.0).to(dev).train()


def batch(r):
    rng = np.random.Generator(np.random.PCG64(50 + r))
    lens = [32000, 24000]
    wav = waveforms(2, 32000, lens, seed=900 + 10 * r)
    T = 99
    phn = np.zeros((2, T), dtype=np.int64)
    tvt = np.full((2, T, 9), -100.0, dtype=np.float32)
    for b, n in enumerate([99, 74]):
        phn[b, :n] = rng.integers(1, 46, size=n)
        tvt[b, :n] = rng.standard_normal((n, 9), dtype=np.float32)
    return (0, wav.to(dev), torch.tensor(lens, device=dev), torch.from_numpy(phn).to(dev),
            *[torch.from_numpy(tvt[:, :, i]).to(dev) for i in range(9)])


m.enable_data_parallel(layers_per_bucket=4)                                          # ... broadcast makes them equal
probe = m.wav2vec2.encoder.layers[2].feed_forward.output_dense.weight
gathered = [torch.empty_like(probe) for _ in range(world)]
dist.all_gather(gathered, probe.detach())
assert all(torch.equal(g, gathered[0]) for g in gathered), "weights differ after the broadcast"

# reference: every rank computes the single-rank gradients of ALL ranks' batches without the reducer, and averages
red = m._reducer
object.__setattr__(m, "_reducer", None)
gb = m.grad_buffer()
ref = torch.zeros_like(gb.flat)
for r in range(world):
    gb.zero()
    m(*batch(r))["loss"].backward()
    ref += gb.flat / world
object.__setattr__(m, "_reducer", red)
opt = FusedAdam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
opt.zero_grad()
m(*batch(rank))["loss"].backward()                                                   # own batch + overlapped all-reduce
torch.cuda.synchronize()
err = float((gb.flat - ref).norm() / ref.norm())
# the wgrad / dQ reductions use fp32 atomics: run-to-run summation order differs at the 1e-6 level
assert err < 1e-4, f"rank {rank}: all-reduced gradient differs from the mean of single-rank gradients ({err:.2e})"
opt.step()
dist.all_gather(gathered, probe.detach())
assert all(torch.equal(g, gathered[0]) for g in gathered), "weights diverged after the data-parallel step"

# ---- optimizer overlap: the reducer leaves the last waits to FusedAdam, which updates each all-reduce region as soon as
# its reduction has landed.  Same batches, same starting weights and moments -> bit-identical weights after the step.
import copy  # noqa: E402

state0 = copy.deepcopy(m.state_dict())
adam0 = copy.deepcopy(opt.state_dict())


def one_step(overlap):
    m.load_state_dict(state0)
    opt.load_state_dict(copy.deepcopy(adam0))
    m.enable_data_parallel(layers_per_bucket=4, broadcast=False, overlap_optimizer=overlap)
    opt.zero_grad()
    m(*batch(rank))["loss"].backward()
    if overlap:
        assert len(m.grad_buffer().pending) >= 3, "deferred mode must leave the reductions pending"
    opt.step()
    assert m.grad_buffer().pending == []
    torch.cuda.synchronize()
    return [p.detach().clone() for p in m.parameters()]


w_plain = one_step(False)
w_over = one_step(True)
# the backward's fp32 atomics make two runs differ at the 1e-7 level; a region updated BEFORE its all-reduce landed
# (or twice, or not at all) would differ by the whole Adam step (lr 1e-4)
worst = max(float((a - b).abs().max()) for a, b in zip(w_plain, w_over))
assert worst < 2e-6, f"rank {rank}: optimizer-overlap step differs from the plain data-parallel step ({worst:.2e})"
dist.all_gather(gathered, probe.detach())
assert all(torch.equal(g, gathered[0]) for g in gathered), "weights diverged after the overlapped step"

# ---- Force_APTAI: the tail's flat gradient buffer is averaged by one all-reduce at the end of the backward
from aptai_b200 import Force_APTAI, Wav2Vec2_PR  # noqa: E402

torch.manual_seed(100 + rank)                                                        # different tail init per rank
fa = Force_APTAI("unused", dev, vocab, w2v2_pr=Wav2Vec2_PR(cfg, None, name, vocab)).to(dev).train()
fa.frame_drop.p = fa.pe_phn.dropout.p = fa.rnn.linear[1].p = 0.0


def fbatch(r):
    rng = np.random.Generator(np.random.PCG64(70 + r))
    lens = [32000, 24000]
    wav = waveforms(2, 32000, lens, seed=950 + 10 * r)
    tvt = torch.from_numpy(rng.standard_normal((2, 99, 9), dtype=np.float32)).to(dev)
    seqs = [rng.integers(1, 46, size=int(n)).astype(np.int64) for n in rng.integers(8, 30, size=2)]
    return (0, wav.to(dev), torch.tensor(lens, device=dev), None, None,
            *[tvt[:, :, i].contiguous() for i in range(9)]), seqs


fgb = fa.enable_data_parallel()
fprobe = fa.rnn.lstm.weight_hh_l0
fg = [torch.empty_like(fprobe) for _ in range(world)]
dist.all_gather(fg, fprobe.detach())
assert all(torch.equal(g, fg[0]) for g in fg), "Force_APTAI weights differ after the broadcast"
dp = fa._dp
object.__setattr__(fa, "_dp", None)
fref = torch.zeros_like(fgb.flat)
for r in range(world):
    fgb.zero()
    args, seqs = fbatch(r)
    fa(*args, phn_seqs=seqs)["loss"].backward()
    fref += fgb.flat / world
object.__setattr__(fa, "_dp", dp)
fopt = FusedAdam([p for p in fa.parameters() if p.requires_grad], lr=1e-4)
fopt.zero_grad()
args, seqs = fbatch(rank)
fa(*args, phn_seqs=seqs)["loss"].backward()
torch.cuda.synchronize()
ferr = float((fgb.flat - fref).norm() / fref.norm())
assert ferr < 1e-4, f"rank {rank}: Force_APTAI all-reduced gradient differs from the mean ({ferr:.2e})"
fopt.step()
dist.all_gather(fg, fprobe.detach())
assert all(torch.equal(g, fg[0]) for g in fg), "Force_APTAI weights diverged after the data-parallel step"
dist.barrier()
if rank == 0:
    print(f"NCCL_DP_OK world={world} grad_rel_err={err:.2e} overlap_step_max_abs_diff={worst:.2e} "
          f"force_grad_rel_err={ferr:.2e}")
dist.destroy_process_group()
