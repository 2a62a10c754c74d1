"""Training-step plumbing for the drop-in modules: flat gradient buffer, fused Adam, data-parallel gradient
all-reduce, and the autograd hook that lets the reference's loop (`train/train_aptai.py:431-443`:
`optimizer.zero_grad(); out = model(...); out['loss'].backward(); optimizer.step()`) drive the hand-written
backward kernels.

Design (B200-first):
  * every trainable parameter's `.grad` is a view into ONE flat fp32 buffer (`GradBuffer`), laid out so that the
    q/k/v projection gradients of a layer are adjacent — the fused [3H, H] wgrad writes all three at once — and so
    that the data-parallel all-reduce is a handful of large NCCL calls over contiguous memory instead of ~400;
  * no activation recomputation: the reference turns on gradient checkpointing (models/aptai.py:38) to fit 16-40 GB
    GPUs; a 32 x 8 s batch keeps ~12 GB of saved activations, which 180 GB of HBM3e holds outright;
  * the optimizer is one kernel launch over all tensors (`aptai_adam_step`), bit-compatible with torch.optim.Adam.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import lib as _lib
from .lib import check

F32 = torch.float32
_ALIGN = 64   # elements: every tensor starts on a 256-byte boundary of the flat buffer


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class GradBuffer:
    """Flat fp32 gradient storage; `p.grad` of every registered parameter is a view into it."""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]],
                 fused_groups: Sequence[Sequence[str]] = ()):
        params = [(n, p) for n, p in named_params if p.requires_grad]
        by_name = dict(params)
        order: List[str] = []
        seen = set()
        self._fused: Dict[Tuple[str, ...], Tuple[int, int]] = {}
        group_of = {}
        for grp in fused_groups:
            grp = tuple(grp)
            if all(n in by_name for n in grp):
                for n in grp:
                    group_of[n] = grp
        for n, _ in params:
            if n in seen:
                continue
            for m in group_of.get(n, (n,)):
                order.append(m)
                seen.add(m)
        dev = params[0][1].device
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in order:
            p = by_name[n]
            grp = group_of.get(n)
            if grp is None or n == grp[0]:
                off = (off + _ALIGN - 1) // _ALIGN * _ALIGN      # members of a fused group stay contiguous
            self.offsets[n] = off
            off += p.numel()
        self.numel = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros((self.numel,), dtype=F32, device=dev)
        self.params = [(n, by_name[n]) for n in order]
        for n, p in self.params:
            if p.dtype != F32:
                raise TypeError(f"GradBuffer: parameter {n} is {p.dtype}; the training path keeps fp32 master weights")
            p.grad = self.flat[self.offsets[n]: self.offsets[n] + p.numel()].view(p.shape)
        for grp in set(group_of.values()):
            n0 = grp[0]
            tot = sum(by_name[n].numel() for n in grp)
            self._fused[grp] = (self.offsets[n0], tot)

    def view(self, name: str) -> torch.Tensor:
        return dict(self.params)[name].grad

    def fused(self, names: Sequence[str], shape) -> torch.Tensor:
        off, tot = self._fused[tuple(names)]
        return self.flat[off: off + tot].view(shape)

    def zero(self) -> None:
        self.flat.zero_()

    def owns(self, p: torch.nn.Parameter) -> bool:
        g = p.grad
        return (g is not None and g.untyped_storage().data_ptr() == self.flat.untyped_storage().data_ptr())

    # ---- data parallel ------------------------------------------------------------------------------------
    def allreduce(self, group=None, bucket_bytes: int = 256 << 20, average: bool = True, async_op: bool = False):
        """Sum (or average) the gradients over the data-parallel group: a few large all-reduces over the flat
        buffer (NCCL over NVLink/NVSwitch; gloo in the CPU tests)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return []
        world = dist.get_world_size(group)
        n = max(1, bucket_bytes // 4)
        works = []
        for s in range(0, self.numel, n):
            chunk = self.flat[s: s + n]
            if average:
                chunk.div_(world)      # pre-divide: keeps the sum in range and matches DDP's averaging
            works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
        return works


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (the reference's optimizer, train/train_aptai.py:350-356) as ONE kernel launch per step.

    Same constructor arguments and update rule (L2 `weight_decay` added to the gradient, bias correction,
    eps outside the square root); works with torch LR schedulers (reads `param_groups[0]['lr']`).  If the
    parameters' `.grad` are not yet views of a `GradBuffer`, one is created."""

    CHUNK = 1 << 16

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 grad_buffer: Optional[GradBuffer] = None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no trainable parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam: a single parameter group is supported")
        if not params[0].is_cuda:
            raise RuntimeError("FusedAdam: parameters must be on a CUDA (sm_100) device; there is no CPU path")
        if grad_buffer is None or not all(grad_buffer.owns(p) for p in params):
            grad_buffer = GradBuffer([(f"p{i}", p) for i, p in enumerate(params)])
        self.gb = grad_buffer
        dev = params[0].device
        self._params = params
        flat0 = self.gb.flat.data_ptr()
        offs = [(p.grad.data_ptr() - flat0) // 4 for p in params]
        nums = [p.numel() for p in params]
        for p in params:
            if not p.is_contiguous() or p.dtype != F32:
                raise TypeError("FusedAdam: parameters must be contiguous fp32")
        self._ptrs = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev)
        self._offs = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._nums = torch.tensor(nums, dtype=torch.int64, device=dev)
        chunks = []
        for i, n in enumerate(nums):
            for s in range(0, n, self.CHUNK):
                chunks.append((i, s))
        ck = torch.zeros((len(chunks), 2), dtype=torch.int64)
        for j, (i, s) in enumerate(chunks):
            ck[j, 0] = i            # int32 tensor index in the low word (little endian), padding word zero
            ck[j, 1] = s
        self._chunks = ck.to(dev)
        self._n_chunks = len(chunks)
        self.exp_avg = torch.zeros_like(self.gb.flat)
        self.exp_avg_sq = torch.zeros_like(self.gb.flat)
        self._step = 0
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none: bool = False) -> None:     # keep the views, clear the storage
        self.gb.zero()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self._step += 1
        check(_lib.load().aptai_adam_step(self._ptrs.data_ptr(), self._offs.data_ptr(), self._nums.data_ptr(),
                                          self._chunks.data_ptr(), self._n_chunks, self.CHUNK,
                                          self.gb.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                          float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                          float(g["weight_decay"]), self._step, float(self.grad_scale), _stream()),
              "adam_step")
        # the kernel wrote through raw pointers: tell autograd / the kernel-weight cache that the values changed
        torch.autograd.graph.increment_version(self._params)
        return loss


class _BackwardHook(torch.autograd.Function):
    """Makes `out['loss'].backward()` run the hand-written backward: the forward stores a closure that launches the
    backward kernels (writing into the GradBuffer views); autograd only delivers the upstream scalar gradient."""

    @staticmethod
    def forward(ctx, loss_value: torch.Tensor, anchor: torch.Tensor, run_backward):
        ctx.run_backward = run_backward
        return loss_value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.run_backward(grad_out)
        return None, None, None


def attach_backward(loss_value: torch.Tensor, anchor: torch.nn.Parameter, run_backward) -> torch.Tensor:
    """`anchor` is any trainable parameter: it makes the returned scalar require grad."""
    return _BackwardHook.apply(loss_value, anchor, run_backward)
