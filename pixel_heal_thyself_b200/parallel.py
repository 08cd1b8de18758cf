"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch.

Batch / learning-rate policy: ``trainer.batch_size`` is the PER-RANK batch (weak scaling, the benchmark's contract), so
the global batch is world x batch_size; the gradient is the MEAN over the global batch (SUM all-reduce, 1/world folded
into the Adam kernel) and lr_g / lr_d are used unscaled, exactly as a single process with the larger batch would.

The reference is single-device (base_trainer.py:46).  The generator has no
cross-sample operation and L1 is a mean, so data parallelism is exact: each rank
runs the hot path on its shard of the patch batch and the only exchange is one
SUM all-reduce of the generator's flat gradient arena (9.28 M fp32 = 37 MB per
step), averaged inside the fused Adam kernel (grad_scale = 1/world).  By default
it is ONE all-reduce right after backward (~0.1-0.15 ms exposed); with
PHT_GRAD_ALLREDUCE=overlap it is issued bucket by bucket on a side stream while
backward is still producing the earlier layers' gradients (decoder -> block 4..0
-> encoders) -- measured slower (2 GPUs 11.06 vs 10.89 ms/step, 4 GPUs 11.39 vs
11.07): NCCL's CTAs occupy SMs that the persistent one-CTA-per-SM kernels need.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_distributed() -> tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the default group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # PHT_DIST_BACKEND=gloo: several ranks may share ONE GPU (NCCL refuses duplicate devices); gloo all-reduces CUDA
        # tensors through the host -- used by the single-GPU data-parallel parity test, never for throughput
        backend = os.environ.get("PHT_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            local = local % torch.cuda.device_count()
            torch.cuda.set_device(local)
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_indices(n: int, rank: int, world: int, batch: int, perm: torch.Tensor) -> torch.Tensor:
    """Rank's strided slice of a global permutation.  One process: the whole permutation (the reference's DataLoader
    keeps the final partial batch: ceil(n / batch) iterations, base_trainer.py:363).  Several ranks: trimmed so that
    every rank runs the same number of full batches (a ragged last step would dead-lock the gradient all-reduce)."""
    if world == 1:
        return perm
    per_rank = (n // (world * batch)) * batch
    return perm[rank::world][:per_rank]


def bucket_ranges(offsets: dict[str, tuple[int, int]], order: list[str], total: int) -> list[tuple[str, int, int]]:
    """Contiguous [lo, hi) ranges of the flat gradient arena in the order backward completes them:
    decoder, transformer_blocks.{last..0}, encoders.  ``order`` is the parameter registration order."""
    def group(name: str) -> str:
        if name.startswith("decoder."):
            return "decoder"
        if name.startswith("transformer_blocks."):
            return "block" + name.split(".")[1]
        return "encoders"

    groups: dict[str, list[int]] = {}
    for n in order:
        lo = offsets[n][0]
        g = groups.setdefault(group(n), [lo, lo])
        g[0], g[1] = min(g[0], lo), max(g[1], lo)
    starts = sorted((v[0], k) for k, v in groups.items())
    out = {}
    for i, (lo, k) in enumerate(starts):
        hi = starts[i + 1][0] if i + 1 < len(starts) else total
        out[k] = (lo, hi)
    blocks = sorted((k for k in out if k.startswith("block")), key=lambda s: -int(s[5:]))
    seq = (["decoder"] if "decoder" in out else []) + blocks + (["encoders"] if "encoders" in out else [])
    return [(k, out[k][0], out[k][1]) for k in seq]


class GradBucketer:
    """All-reduce of a flat gradient arena, in one of two schedules (environment ``PHT_GRAD_ALLREDUCE``):

    ``end`` (default): ONE all-reduce of the whole arena on the compute stream in ``finish()``: nothing overlaps the
    backward pass, so NCCL's CTAs never take SMs away from the persistent one-CTA-per-SM kernels (a 148-CTA kernel
    that finds an SM occupied runs that CTA's tiles as a second wave).

    ``overlap``: ``ready(tag)`` is called by the backward schedule as soon as the gradients of a bucket are final (on
    the compute stream); the all-reduce of that slice is enqueued on a side stream behind an event, and ``finish()``
    makes the compute stream wait for all outstanding buckets."""

    def __init__(self, get_flat_grad, offsets, order, total, group=None, allreduce=None):
        """``allreduce(tensor)``: replaces the plain SUM all-reduce of the ``end`` schedule (the trainer's CUDA-graph
        recorder splits its capture around the collective there)."""
        self.allreduce = allreduce
        self.get_flat_grad = get_flat_grad
        self.total = total
        self.mode = os.environ.get("PHT_GRAD_ALLREDUCE", "end")
        assert self.mode in ("overlap", "end"), "PHT_GRAD_ALLREDUCE must be overlap or end"
        self.ranges = {k: (lo, hi) for k, lo, hi in bucket_ranges(offsets, order, total)}
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._stream = None
        self._pending = []
        self.launched: list[str] = []

    def ready(self, tag: str) -> None:
        if self.world == 1 or tag not in self.ranges:
            return
        lo, hi = self.ranges[tag]
        g = self.get_flat_grad()[lo:hi]
        self.launched.append(tag)
        if self.mode == "end":
            return
        if g.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            g.record_stream(self._stream)
        else:
            self._pending.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, aliased: bool = True) -> None:
        """Complete the step's exchange.  ``aliased`` = every ``p.grad`` aliases the arena backward wrote
        (``FlatAdam.gather_grads``, which the caller runs first): if not, the gradients were only just gathered into the
        arena, so whatever the overlapped buckets reduced during backward is stale and the whole arena is reduced here."""
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self.world > 1 and (self.mode == "end" or not aliased):
            if self.mode == "overlap":
                raise RuntimeError("PHT_GRAD_ALLREDUCE=overlap needs p.grad to alias the flat gradient arena (use "
                                   "zero_grad(set_to_none=True), no gradient accumulation / hooks), or use the default "
                                   "PHT_GRAD_ALLREDUCE=end")
            if self.allreduce is not None:
                self.allreduce(self.get_flat_grad()[:self.total])
            else:
                dist.all_reduce(self.get_flat_grad()[:self.total], op=dist.ReduceOp.SUM, group=self.group)
        self.launched.clear()


def allreduce_module_grads(module: torch.nn.Module, world: int, allreduce=None) -> None:
    """Average the gradients of a stock PyTorch module (the critic) across ranks with one flat all-reduce."""
    if world == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    flat = torch._utils._flatten_dense_tensors(grads)
    if allreduce is not None:
        allreduce(flat)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
