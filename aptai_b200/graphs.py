"""CUDA-graph replay for the launch-bound single-utterance calls (`get_aptai_output`, `get_ctc_logits`, ...:
BASELINE config 1).  One 4 s utterance through a 12-layer backbone is ~110 kernel launches of a few microseconds
each; issued one by one from Python the call is bound by launch overhead, not by the GPU.  The first call for a
given input length runs eagerly (lazy initialisation: kernel attributes, the weight plan), the second captures the
whole launch sequence — TMA descriptors are kernel parameters, so they are baked into the graph nodes — and later
calls copy the waveform into the static input buffer and replay one graph.

A captured graph is only valid while every buffer it references stays where it was: the cache is keyed by the
backbone's plan object and its buffer generation (weights are refreshed IN PLACE by `aptai_prepare_weights`, so an
optimizer step or load_state_dict does not invalidate the graphs; moving the model does)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Tuple

import torch


class GraphCache:
    def __init__(self, max_entries: int = 8):
        self.max_entries = max_entries
        self._entries: "OrderedDict[tuple, dict]" = OrderedDict()

    def clear(self) -> None:
        self._entries.clear()

    def run(self, key: tuple, fn: Callable[..., Tuple[torch.Tensor, ...]], *inputs: torch.Tensor):
        """`fn(*inputs)` -> tuple of tensors.  Returns the outputs (static buffers of the graph once captured: copy
        or consume them before the next call with the same key)."""
        e = self._entries.get(key)
        if e is None:
            e = {"calls": 0}
            self._entries[key] = e
            while len(self._entries) > self.max_entries:
                self._entries.popitem(last=False)
        else:
            self._entries.move_to_end(key)
        e["calls"] += 1
        if e["calls"] == 1:                      # eager: lazy initialisation must not happen under capture
            return fn(*inputs)
        if "graph" not in e:
            static_in = tuple(t.clone() for t in inputs)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = fn(*static_in)
            e.update(graph=g, static_in=static_in, static_out=static_out)
        for dst, src in zip(e["static_in"], inputs):
            dst.copy_(src, non_blocking=True)
        e["graph"].replay()
        return e["static_out"]
