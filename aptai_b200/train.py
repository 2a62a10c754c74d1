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

import bisect
import ctypes as C
import weakref
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import lib as _lib
from .lib import check

F32 = torch.float32
_ALIGN = 64   # elements: every tensor starts on a 256-byte boundary of the flat buffer


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_BUFFERS: Dict[int, "weakref.ReferenceType"] = {}      # flat storage pointer -> GradBuffer (FusedAdam finds pending waits)


class GradBuffer:
    """Flat fp32 gradient storage; `p.grad` of every registered parameter is a view into it.

    `pending`: all-reduces of regions of the buffer that have been LAUNCHED but not yet waited for
    (`GradReducer(defer_wait=True)`): a list of (wait, lo, hi) in launch order.  `FusedAdam.step()` consumes it region
    by region; anything else that reads `.grad` first calls `wait_pending()`."""

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
        self.pending: List[Tuple] = []
        _BUFFERS[self.flat.untyped_storage().data_ptr()] = weakref.ref(self)
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

    def wait_pending(self) -> None:
        """Make the current stream wait for every launched-but-unawaited all-reduce of this buffer."""
        pend, self.pending = self.pending, []
        for wait, _, _ in pend:
            wait()

    def zero(self) -> None:
        self.wait_pending()
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


class GradReducer:
    """Data-parallel gradient exchange overlapped with the backward pass (BASELINE config 4).

    The flat buffer is laid out in parameter order, so the gradients of encoder layers [i, i+k) are one contiguous
    region that is final as soon as the backward has passed layer i: `layer_done(i)` launches an asynchronous
    all-reduce of that region (NCCL runs it on its own stream over NVLink/NVSwitch while the remaining layers'
    kernels keep the SMs busy); `finish()` reduces what is left (heads, positional conv, projection, norms) and
    makes the compute stream wait for all of it.  Averaging uses NCCL's AVG reduction (SUM + divide elsewhere)."""

    def __init__(self, gb: GradBuffer, layer_prefix: str, n_layers: int, layers_per_bucket: int = 4, group=None,
                 defer_wait: bool = False):
        """`defer_wait`: `finish()` launches the last reductions but leaves the waiting to the consumer of the
        gradients (`gb.pending`): `FusedAdam.step()` then updates each region as soon as ITS all-reduce has landed, so
        the optimizer runs under the reductions that are still in flight instead of behind the last one.  Anything
        else that reads `.grad` between backward and step must call `gb.wait_pending()` (the reference's training
        loops read nothing there: train/train_aptai.py:431-443)."""
        self.gb, self.group, self.k = gb, group, max(1, layers_per_bucket)
        self.defer_wait = defer_wait
        self.n_layers = n_layers
        spans = []
        for i in range(n_layers):
            pre = f"{layer_prefix}{i}."
            offs = [(gb.offsets[n], gb.offsets[n] + p.numel()) for n, p in gb.params if n.startswith(pre)]
            spans.append((min(o[0] for o in offs), max(o[1] for o in offs)))
        for a, b in zip(spans[:-1], spans[1:]):
            if a[1] > b[0]:
                raise RuntimeError("GradReducer: encoder layers are not laid out in order in the gradient buffer")
        self.spans = spans
        self.works = []
        self._done = [False] * n_layers

    def _launch(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        chunk = self.gb.flat[lo:hi]
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl":
            w = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self.works.append((w.wait, lo, hi))
        else:
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

            def wait(w=w, chunk=chunk, world=world):
                w.wait()
                chunk.div_(world)
            self.works.append((wait, lo, hi))

    def layer_done(self, i: int) -> None:
        """Called by the backward after layer i's gradients are complete (layers arrive in decreasing order)."""
        self._done[i] = True
        if i % self.k == 0:
            hi_layer = min(self.n_layers, i + self.k) - 1
            self._launch(self.spans[i][0], self.spans[hi_layer][1])

    def finish(self) -> None:
        self._launch(0, self.spans[0][0])
        self._launch(self.spans[-1][1], self.gb.numel)
        self.gb.pending.extend(self.works)
        self.works = []
        self._done = [False] * self.n_layers
        if not self.defer_wait:
            self.gb.wait_pending()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Every rank starts from rank `src`'s weights (DDP's construction-time broadcast)."""
    ts = list(module.parameters()) + list(module.buffers())
    for t in ts:
        dist.broadcast(t.data, src=src, group=group)
    # `.data` writes do not bump the version counters the kernel-weight cache (Wav2Vec2Backbone.plan) keys on
    torch.autograd.graph.increment_version(ts)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (the reference's optimizer, train/train_aptai.py:350-356) as ONE kernel launch per step.

    Same constructor arguments and update rule (L2 `weight_decay` added to the gradient, bias correction,
    eps outside the square root); works with torch LR schedulers (reads `param_groups[0]['lr']`).
    The gradients must live in one flat buffer — the models' `grad_buffer()` (attached on their first training
    forward) or an explicit `GradBuffer`; the optimizer binds to whatever buffer the `.grad` views point into when
    `step()` runs, so it can be constructed before the first forward exactly like torch.optim.Adam."""

    CHUNK = 1 << 16

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdam: no trainable parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam: a single parameter group is supported")
        if not params[0].is_cuda:
            raise RuntimeError("FusedAdam: parameters must be on a CUDA (sm_100) device; there is no CPU path")
        for p in params:
            if not p.is_contiguous() or p.dtype != F32:
                raise TypeError("FusedAdam: parameters must be contiguous fp32")
        dev = params[0].device
        self._params = params
        nums = [p.numel() for p in params]
        soffs, off = [], 0
        for n in nums:
            soffs.append(off)
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        self._ptrs = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev)
        self._soffs = torch.tensor(soffs, dtype=torch.int64, device=dev)
        self._soffs_host = soffs
        self._nums = torch.tensor(nums, dtype=torch.int64, device=dev)
        chunks = [(i, s) for i, n in enumerate(nums) for s in range(0, n, self.CHUNK)]
        ck = torch.tensor(chunks, dtype=torch.int64).reshape(-1, 2)   # {int32 tensor | pad} (little endian), int64 start
        self._chunks = ck.to(dev)
        self._n_chunks = len(chunks)
        self.exp_avg = torch.zeros((off,), dtype=F32, device=dev)
        self.exp_avg_sq = torch.zeros((off,), dtype=F32, device=dev)
        self._step = 0
        self.grad_scale = 1.0
        self._bound = None        # (storage ptr, first grad ptr) the offsets below were computed for
        self._goffs = None
        self._gbase = 0

    def _bind(self) -> bool:
        """Locate the flat gradient buffer behind the parameters' .grad views; False if no gradient exists yet."""
        grads = [p.grad for p in self._params]
        if all(g is None for g in grads):
            return False
        if any(g is None for g in grads):
            raise RuntimeError("FusedAdam: some parameters have no gradient; the fused step needs all of them in one "
                               "flat buffer (model.grad_buffer() / GradBuffer)")
        st = grads[0].untyped_storage().data_ptr()
        key = (st, tuple(g.data_ptr() for g in grads[:4]), grads[-1].data_ptr())
        if key == self._bound:
            return True
        for g in grads:
            if g.untyped_storage().data_ptr() != st or g.dtype != F32 or not g.is_contiguous():
                raise RuntimeError("FusedAdam: gradients must be contiguous fp32 views of ONE flat buffer "
                                   "(model.grad_buffer() / GradBuffer)")
        self._gbase = st
        goffs = [(g.data_ptr() - st) // 4 for g in grads]
        dev = self._params[0].device
        self._goffs = torch.tensor(goffs, dtype=torch.int64, device=dev)
        # chunk table in the order of the flat gradient buffer: a region [lo, hi) of it (one all-reduce bucket) is
        # then a contiguous range of chunks, found by bisection on the chunks' gradient offsets
        chunks = sorted(((i, s) for i, p in enumerate(self._params) for s in range(0, p.numel(), self.CHUNK)),
                        key=lambda c: goffs[c[0]] + c[1])
        self._chunk_goff = [goffs[i] + s for i, s in chunks]
        self._chunks = torch.tensor(chunks, dtype=torch.int64).reshape(-1, 2).to(dev)
        self._n_chunks = len(chunks)
        self._bound = key
        return True

    def _grad_buffer(self) -> Optional[GradBuffer]:
        ref = _BUFFERS.get(self._gbase)
        return ref() if ref is not None else None

    def zero_grad(self, set_to_none: bool = False) -> None:     # keep the views, clear the storage
        for p in self._params:
            if p.grad is not None:
                if self._bind():
                    gb = self._grad_buffer()
                    if gb is not None:
                        gb.wait_pending()
                    g0 = self._params[0].grad
                    # one memset over the whole flat buffer instead of one per tensor
                    torch.empty(0, dtype=F32, device=g0.device).set_(g0.untyped_storage()).zero_()
                return

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if not self._bind():
            return loss
        g = self.param_groups[0]
        self._step += 1

        def launch(c0: int, c1: int) -> None:
            if c1 <= c0:
                return
            check(_lib.load().aptai_adam_step(self._ptrs.data_ptr(), self._goffs.data_ptr(), self._soffs.data_ptr(),
                                              self._nums.data_ptr(), self._chunks.data_ptr() + 16 * c0, c1 - c0, self.CHUNK,
                                              self._gbase, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                              float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                              float(g["eps"]), float(g["weight_decay"]), self._step,
                                              float(self.grad_scale), _stream()), "adam_step")

        gb = self._grad_buffer()
        pend = []
        if gb is not None and gb.pending:
            pend, gb.pending = gb.pending, []
        if not pend:
            launch(0, self._n_chunks)
        else:
            # data parallel with deferred waits: update every all-reduce region as soon as its reduction has landed
            # (the regions were launched last-layers-first), under the reductions still in flight
            done = []
            for wait, lo, hi in pend:
                wait()
                c0, c1 = bisect.bisect_left(self._chunk_goff, lo), bisect.bisect_left(self._chunk_goff, hi)
                launch(c0, c1)
                done.append((c0, c1))
            done.sort()
            at = 0
            for c0, c1 in done:                  # anything no region covered (none with GradReducer's buckets)
                launch(at, c0)
                at = max(at, c1)
            launch(at, self._n_chunks)
        # the kernel wrote through raw pointers: tell autograd / the kernel-weight cache that the values changed
        torch.autograd.graph.increment_version(self._params)
        return loss

    # ---- checkpointing: torch.optim.Adam's layout, so optimizer.pt files are interchangeable with the reference's
    # (train/train_phoneme_recognizer.py:396 loads it, :483 saves it)
    def _moment_views(self, i: int):
        off, n = int(self._soffs_host[i]), self._params[i].numel()
        shape = self._params[i].shape
        return self.exp_avg[off: off + n].view(shape), self.exp_avg_sq[off: off + n].view(shape)

    def state_dict(self):
        sd = super().state_dict()
        state = {}
        if self._step > 0:
            for i in range(len(self._params)):
                m, v = self._moment_views(i)
                state[i] = {"step": torch.tensor(float(self._step)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("FusedAdam.load_state_dict: parameter groups do not match this optimizer")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        state = state_dict.get("state", {})
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, pid in enumerate(groups[0]["params"]):
            st = state.get(pid, state.get(str(pid)))
            if st is None:
                continue
            m, v = self._moment_views(i)
            m.copy_(st["exp_avg"].to(m.device, F32).view_as(m))
            v.copy_(st["exp_avg_sq"].to(v.device, F32).view_as(v))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FusedAdam.load_state_dict: per-parameter step counts differ; the fused step keeps one")
        self._step = steps.pop() if steps else 0


class _BackwardHook(torch.autograd.Function):
    """Makes `out['loss'].backward()` run the hand-written backward: the forward stores a closure that launches the
    backward kernels (writing into the GradBuffer views); autograd only delivers the upstream scalar gradient."""

    @staticmethod
    def forward(ctx, loss_value: torch.Tensor, anchor: torch.Tensor, run_backward):
        ctx.run_backward = run_backward
        return loss_value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        run, ctx.run_backward = ctx.run_backward, None
        if run is None:
            raise RuntimeError("aptai_b200: backward through this loss a second time (its activations were freed, "
                               "as stock autograd does without retain_graph)")
        # the closure owns every saved activation of the step: dropping it here frees them even if the caller keeps
        # the loss tensor (and with it this node) alive, e.g. `sum_train_loss += train_loss`
        # (train/train_aptai.py:446)
        run(grad_out)
        return None, None, None


def attach_backward(loss_value: torch.Tensor, anchor: torch.nn.Parameter, run_backward) -> torch.Tensor:
    """`anchor` is any trainable parameter: it makes the returned scalar require grad."""
    return _BackwardHook.apply(loss_value, anchor, run_backward)
