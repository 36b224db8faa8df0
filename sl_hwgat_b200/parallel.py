"""Batch-sharded data parallelism for HWGATE training.

One process per GPU.  The batch of sign sequences is split across ranks; every
op of the model is per-sample (LayerNorm only, windows never cross samples), so
the only exchange is ONE mean all-reduce of the parameter gradients per step
(the loss is a batch mean, SmoothCrossEntropy.py:39).  Gradients are packed
into a few flat buckets in reverse registration order - the order backward
produces them - and each bucket's all-reduce is launched on a side stream as
soon as its last gradient has been accumulated, so NCCL (NVLink 5 / NVSwitch)
overlaps the rest of backward.  Inference shards the batch with no collective.

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


def sync_threshold_rng(seed: int) -> None:
    """The training threshold of MSA.forward is ONE scalar per call for the whole
    batch, drawn from the CPU generator (HWGATE.py:96): all ranks must draw the
    same sequence, so they seed the CPU generator identically."""
    torch.manual_seed(seed)


class GradientAllReduce:
    """Bucketed, backward-overlapped mean all-reduce of a module's gradients.

        sync = GradientAllReduce(model, bucket_bytes=8 << 20)
        loss.backward()          # hooks launch one async all-reduce per full bucket
        sync.finish()            # wait, copy the averaged values back into .grad
    """

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None):
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
        self._flat: List[torch.Tensor] = []
        for bi, b in enumerate(self.buckets):
            self._flat.append(torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device))
            for p in b:
                self._bucket_of[p] = bi
        self._pending = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._hooks = []
        self._stream = None
        if self.world > 1:
            if self.params and self.params[0].is_cuda:
                self._stream = torch.cuda.Stream()
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.reset()

    def reset(self) -> None:
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def _launch(self, bi: int) -> None:
        bucket, flat = self.buckets[bi], self._flat[bi]
        self._launched[bi] = True

        def pack_and_reduce():
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            flat.div_(self.world)
            self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                pack_and_reduce()
        else:
            pack_and_reduce()

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> None:
        """Block the current stream on every bucket and scatter the means back into .grad."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:  # parameters that received no gradient this step
                self._launch(bi)
        for bi, bucket in enumerate(self.buckets):
            self._works[bi].wait()
            if self._stream is not None:
                torch.cuda.current_stream().wait_stream(self._stream)
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(self._flat[bi][off:off + n].view_as(p))
                off += n
        self.reset()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
