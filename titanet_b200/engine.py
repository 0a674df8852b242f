"""Step runners for the TitaNet hot path: CUDA-graph capture of a whole training step and
the data-parallel gradient exchange.

The reference's training step (src/learn.py:88-135) is ``model(spectrograms, speakers) ->
loss.backward()`` issued op by op from Python; here the same step (waveform -> mel -> model
-> loss -> backward) is captured once into a CUDA graph and replayed, so the ~900 kernel
launches of a TitaNet-S step cost one graph launch on the host.  Dropout stays fresh
across replays because the masks are keyed by a device-resident seed that a captured
kernel advances (csrc: ``tn_seed_next``).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib
from . import _ops as ops


class GraphedTrainStep:
    """Capture ``mel -> model(x, speakers) -> loss.backward()`` for fixed shapes.

    ``step(wave, labels)`` copies the inputs into static device buffers (async, pinned host
    tensors welcome), replays the graph and returns the static ``loss`` tensor; the
    parameters' ``.grad`` tensors are static too (overwritten by every replay), so an
    optimizer or an all-reduce can follow.  ``use_graph=False`` runs the same step eagerly.
    """

    def __init__(self, model: torch.nn.Module, mel, batch: int, n_samples: int, device, use_graph: bool = True,
                 warmup: int = 3, after_backward: Optional[Callable[[], None]] = None, use_arena: bool = True,
                 lengths: Optional[torch.Tensor] = None):
        self.model, self.mel, self.device = model, mel, torch.device(device)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.wave = torch.zeros(batch, n_samples, device=self.device)
        self.labels = torch.zeros(batch, dtype=torch.int64, device=self.device)
        # ragged batches (datasets.collate_fn semantics): samples per utterance, a static int32 buffer like the others
        self.lengths = None if lengths is None else lengths.to(self.device, torch.int32).clone()
        self.after_backward = after_backward
        self.loss = self.emb = self.preds = None
        self.graph = None
        self.launches_per_step = 0
        # every zero-initialised accumulator and every parameter gradient of a step lives in one buffer that a
        # single memset clears (ops.ZeroArena): the first (measuring) step sizes it
        self.arena = ops.ZeroArena(self.device) if use_arena else None
        if use_graph:
            self._capture(warmup)

    def _body(self):
        prev = ops.set_arena(self.arena)
        try:
            if self.arena is not None:
                self.arena.begin_step()
            for p in self.params:
                p.grad = None
            self.emb, self.preds, self.loss = self.model(self.mel.batch(self.wave, self.lengths), speakers=self.labels)
            self.loss.backward()
            if self.arena is not None and self.arena.measuring:
                self.arena.finish_measuring()
            if self.after_backward is not None:
                self.after_backward()
        finally:
            ops.set_arena(prev)

    def _capture(self, warmup: int):
        try:    # warm-up runs on a side stream by design; the AccumulateGrad stream-mismatch warning is noise here
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # the eager warm-up steps left a step's worth of activations cached in the default pool; the capture below allocates
        # its own (graph-private) pool of the same size, so return the cached blocks first: the peak reservation is ONE step
        # (TitaNet-S at batch 2048: ~90 GB instead of ~180 GB)
        for p in self.params:
            p.grad = None
        self.emb = self.preds = self.loss = None
        torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.kernel_launches()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.kernel_launches() - before

    def load(self, wave: torch.Tensor, labels: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        self.wave.copy_(wave, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        if lengths is not None:
            if self.lengths is None:
                raise ValueError("this step was captured without per-utterance lengths")
            self.lengths.copy_(lengths, non_blocking=True)

    # ---- double-buffered input staging -------------------------------------------------------------------------------
    # ``prefetch`` copies the NEXT step's (pinned) host inputs into a staging buffer on a copy stream while the current step
    # is still running; ``run_prefetched`` waits for that copy, moves the staging buffer into the static inputs (a 12 MB
    # device-to-device copy: ~4 us) and replays the step.  What a data loader with pinned buffers does; without it every
    # step waits ~0.25 ms for its own PCIe transfer.
    def prefetch(self, wave: torch.Tensor, labels: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        if not hasattr(self, "_stage"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [(torch.empty_like(self.wave), torch.empty_like(self.labels),
                            None if self.lengths is None else torch.empty_like(self.lengths), torch.cuda.Event()) for _ in range(2)]
            self._stage_i = 0
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self._consumed:
                e.record(torch.cuda.current_stream(self.device))
        w, l, n, ev = self._stage[self._stage_i]
        self._copy_stream.wait_event(self._consumed[self._stage_i])      # the step that read this buffer last has consumed it
        with torch.cuda.stream(self._copy_stream):
            w.copy_(wave, non_blocking=True)
            l.copy_(labels, non_blocking=True)
            if n is not None and lengths is not None:
                n.copy_(lengths, non_blocking=True)
            ev.record(self._copy_stream)

    def run_prefetched(self) -> torch.Tensor:
        i = self._stage_i
        w, l, n, ev = self._stage[i]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        self.wave.copy_(w, non_blocking=True)
        self.labels.copy_(l, non_blocking=True)
        if n is not None:
            self.lengths.copy_(n, non_blocking=True)
        self._consumed[i].record(cur)
        self._stage_i = 1 - i
        return self.run()

    def run(self) -> torch.Tensor:
        """One step on whatever is in the static input buffers."""
        if self.graph is not None:
            self.graph.replay()
        else:
            before = _lib.kernel_launches()
            self._body()
            self.launches_per_step = _lib.kernel_launches() - before
        return self.loss

    def step(self, wave: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.load(wave, labels)
        return self.run()


class GradAllReduce:
    """Data-parallel gradient exchange: one all-reduce (SUM, then 1/world) over one flat fp32
    buffer holding every parameter gradient (24.6 MB for TitaNet-S), NCCL over NVLink.
    BatchNorm statistics stay per replica (DDP semantics)."""

    def __init__(self, params, world_size: int, group=None, arena=None):
        self.params = list(params)
        self.world = world_size
        self.group = group
        self.arena = arena          # the step's ops.ZeroArena (GraphedTrainStep.arena): gradients are exchanged in place
        self.flat = None

    def _arena_span(self, grads):
        """The slice of the active ZeroArena that holds every gradient, as one fp32 tensor (or None)."""
        arena = self.arena if self.arena is not None else ops._ARENA
        if arena is None or arena.buf is None:
            return None
        base, size = arena.buf.data_ptr(), arena.buf.numel()
        lo, hi = size, 0
        for g in grads:
            off = g.data_ptr() - base
            if off < 0 or off + g.numel() * g.element_size() > size or g.dtype != torch.float32 or not g.is_contiguous():
                return None
            lo, hi = min(lo, off), max(hi, off + g.numel() * 4)
        lo -= lo % 16
        hi += (-hi) % 16
        return arena.buf[lo:min(hi, size)].view(torch.float32)

    def __call__(self):
        if self.world <= 1:
            return
        import torch.distributed as dist
        grads = [p.grad for p in self.params]
        span = self._arena_span(grads)
        if span is not None:
            # gradients already share one buffer: exchange it in place (no pack / unpack copies)
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(span, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(span, group=self.group)
                span.mul_(1.0 / self.world)
            return
        if self.flat is None:
            self.flat = torch.empty(sum(g.numel() for g in grads), device=grads[0].device, dtype=grads[0].dtype)
        torch._foreach_copy_(list(self.flat.split([g.numel() for g in grads])), [g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat, group=self.group)
        self.flat.mul_(1.0 / self.world)
        torch._foreach_copy_([g.reshape(-1) for g in grads], list(self.flat.split([g.numel() for g in grads])))
