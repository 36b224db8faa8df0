"""Batch-sharded data parallelism for HWGATE training.

One process per GPU.  The batch of sign sequences is split across ranks; every
op of the model is per-sample (LayerNorm only, windows never cross samples), so
the only exchange is ONE mean all-reduce of the parameter gradients per step
(the loss is a batch mean, SmoothCrossEntropy.py:39).  Gradients live in a few
flat buckets in reverse registration order - the order backward produces them
(`.grad` is a view into its bucket: no pack / unpack copies) - and each
bucket's in-place all-reduce is launched on a side stream as soon as its last
gradient has been accumulated, so NCCL (NVLink 5 / NVSwitch) overlaps the rest
of backward.  Inference shards the batch with no collective.

The reference has no distributed code; `torch.distributed` (NCCL on GPU, gloo
in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Join the process group torchrun set up.  Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(n: int, rank: int, world: int) -> slice:
    """Contiguous slice of a batch of n samples owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def sync_threshold_rng(seed: int, module: Optional[torch.nn.Module] = None) -> None:
    """The training threshold of MSA.forward is ONE scalar per call for the whole
    batch, drawn from the CPU generator (HWGATE.py:96): all ranks must draw the
    same sequence.  With `module`, every MSA gets ONE shared dedicated generator
    seeded with `seed` (same on every rank), so data loading and augmentation -
    which consume the global CPU generator at rank-dependent rates - cannot pull
    the ranks' thresholds apart.  Without `module` only the global generator is
    seeded (single-process use: the reference's own behaviour, configs.py:55-59)."""
    torch.manual_seed(seed)
    if module is None:
        return
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    for m in module.modules():
        if hasattr(m, "_draw_threshold") and hasattr(m, "_thr_gen"):
            m._thr_gen = gen


class GradientAllReduce:
    """Bucketed, backward-overlapped mean all-reduce of a module's gradients, with no pack / unpack copies:
    every parameter's `.grad` IS a view into its flat bucket, autograd accumulates into it in place, and each
    bucket is all-reduced in place on a side stream as soon as its last gradient of the step has arrived.

        sync = GradientAllReduce(model, bucket_bytes=8 << 20)
        sync.zero_grad()         # instead of model.zero_grad(): one memset per bucket, .grad stays a bucket view
        loss.backward()          # hooks launch one async all-reduce per complete bucket
        sync.finish()            # the current stream waits for the buckets; .grad holds the mean over ranks

    A `.grad` that was replaced behind our back (e.g. `zero_grad(set_to_none=True)`) is copied into its bucket
    slot and re-pointed when its hook fires, so the result is the same, only slower.
    With `trace=True` CUDA events are recorded around every bucket's collective (`trace_report`).
    """

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None, trace: bool = False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):  # backward reaches the last-registered parameters first
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {}
        self._view = {}
        self._flat: List[torch.Tensor] = []
        for bi, b in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device)
            self._flat.append(flat)
            off = 0
            for p in b:
                self._bucket_of[p] = bi
                self._view[p] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self._avg_native = False
        if self.world > 1:
            try:
                self._avg_native = dist.get_backend(group) == "nccl"
            except Exception:
                self._avg_native = False
        self._hooks = []
        self._stream = None
        self._trace = trace
        self._events: List[tuple] = []
        if self.world > 1:
            if self.params and self.params[0].is_cuda:
                self._stream = torch.cuda.Stream()
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.zero_grad()

    def zero_grad(self) -> None:
        """Zero every bucket (one memset each) and make every `.grad` its bucket view."""
        for flat in self._flat:
            flat.zero_()
        for p in self.params:
            v = self._view[p]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
        self.reset()

    def reset(self) -> None:
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._events = []

    def _launch(self, bi: int) -> None:
        flat = self._flat[bi]
        self._launched[bi] = True

        def reduce():
            ev0 = ev1 = None
            if self._trace and flat.is_cuda:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            if self._avg_native:
                self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                flat.div_(self.world)
                self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            if ev1 is not None:
                self._works[bi].wait()       # stream-level wait (NCCL): orders ev1 after the collective
                ev1.record()
                self._events.append((bi, flat.numel() * flat.element_size(), ev0, ev1))

        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                reduce()
        else:
            reduce()

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        v = self._view[p]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():   # .grad was replaced: move it into the bucket
            v.copy_(p.grad)
            p.grad = v
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> None:
        """Make the current stream wait for every bucket; afterwards `.grad` holds the mean over ranks."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:  # parameters that received no gradient this step
                self._launch(bi)
        for bi in range(len(self.buckets)):
            self._works[bi].wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        events = self._events
        self.reset()
        self._events = events

    def trace_report(self, t0: "torch.cuda.Event") -> list:
        """[(bucket, bytes, start_ms, end_ms)] relative to the event `t0` (call after a synchronize)."""
        return [(bi, nbytes, t0.elapsed_time(e0), t0.elapsed_time(e1)) for bi, nbytes, e0, e1 in self._events]

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
