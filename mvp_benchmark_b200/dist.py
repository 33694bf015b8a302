"""Multi-GPU plumbing for the operator hot path: one process per GPU, the batch sharded across ranks,
NO collective on the operator path (every op is independent per cloud) and ONE all-reduce
of fp32 gradients per training step (what replaces the reference's single-process nn.DataParallel,
completion/train.py:49,141).  Backend-agnostic: NCCL over NVLink on the B200 box, gloo in CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world,
    local_rank); a single-process run (no RANK in the env) returns (0, 1, 0) without initialising."""
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)  # binds the communicator to this rank's GPU up front
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_bounds(total, rank, world):
    """Contiguous slice [lo, hi) of `total` clouds owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank, world, dim=0):
    lo, hi = shard_bounds(t.size(dim), rank, world)
    return t.narrow(dim, lo, hi - lo)


def _pack(bucket):
    """Flat fp32-or-native buffer of a bucket of parameters: every parameter's gradient (zeros where this rank
    has none) followed by one 'used' flag per parameter.  The layout depends on the PARAMETER list only, never on
    which gradients happen to exist on this rank, so all ranks always exchange buffers of equal size and meaning."""
    ref = bucket[0]
    parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(ref.dtype) for p in bucket]
    used = torch.tensor([0.0 if p.grad is None else 1.0 for p in bucket], dtype=ref.dtype, device=ref.device)
    return torch.cat(parts + [used])


def _unpack(bucket, flat, scale):
    """Hands the reduced gradients back WITHOUT copies and, normally, without a host synchronisation: the flat buffer
    is scaled once in place and every parameter's .grad becomes a view of its slice (what DDP calls
    gradient_as_bucket_view).  Only a bucket holding a parameter that THIS rank did not use needs the 'used' flags on
    the host — to keep grad None for a parameter no rank used (as in a single-process step) and to hand over the sum
    for one that only other ranks used, so that the optimiser steps identically everywhere.  (Round 1 copied every
    slice back with two small kernels per parameter and read the flags of every bucket with .tolist(): ~290 launches
    and three host synchronisations per step — 4 ms on a launch-bound 25 ms step.)"""
    nflag = len(bucket)
    if scale != 1.0:
        flat[:flat.numel() - nflag].mul_(scale)
    used = None
    if any(p.grad is None for p in bucket):
        used = flat[flat.numel() - nflag:].tolist()
    off = 0
    for k, p in enumerate(bucket):
        n = p.numel()
        if used is None or used[k] > 0:
            g = flat[off:off + n].view_as(p)
            p.grad = g if g.dtype == p.dtype else g.to(p.dtype)
        off += n


def allreduce_gradients(params, world=None, bucket_bytes=32 << 20, average=True, weight=None):
    """Sum (or average) the .grad of `params` across ranks with as few collectives as possible: gradients are
    flattened into buckets of `bucket_bytes` (sized for launch latency, not link count — NVSwitch gives every pair
    full bandwidth).  Buckets are built from EVERY parameter that requires a gradient, in the order given, with
    zeros standing in for a gradient this rank did not produce: the collectives and their sizes are the same on all
    ranks whatever each rank's graph looked like.  Returns the number of collectives issued.

    `average=True` divides the sum by the world size: the gradient of the mean of the per-rank mean losses.  With
    UNEVEN shards (7 clouds over 2 ranks = 4 + 3) that is not the gradient of the global mean; pass
    `weight = local_count / global_count` (and average=True) to scale each rank's gradient before the sum instead."""
    if not dist.is_initialized():
        return 0
    world = world or dist.get_world_size()
    params = [p for p in params if p.requires_grad]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = _pack(bucket)
        if weight is not None:
            flat[:flat.numel() - len(bucket)] *= float(weight)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        _unpack(bucket, flat, 1.0 / world if (average and weight is None) else 1.0)
        calls += 1
        bucket, size = [], 0

    for p in params:
        nbytes = p.numel() * p.element_size()
        if size + nbytes > bucket_bytes and bucket:
            flush()
        bucket.append(p)
        size += nbytes
    flush()
    return calls


def max_over_ranks(value, device):
    """max of a python float over ranks (timing: a multi-GPU number is the slowest rank's)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized():
        dist.barrier()


class OverlappedGradientAllReduce:
    """The gradient all-reduce of a data-parallel step, overlapped with the backward pass (what replaces
    nn.DataParallel's reduce-to-GPU-0, completion/train.py:49,141).  Parameters are assigned to buckets of
    `bucket_bytes` up front, in REVERSE registration order (gradients arrive roughly last layer first).  A bucket is
    flattened and all-reduced asynchronously once its last gradient has been accumulated
    (`register_post_accumulate_grad_hook`) AND every bucket before it has been launched: collectives are matched
    across ranks by issue order, so they are issued strictly in bucket order on every rank, whichever order a
    rank's graph completes them in.  `finish()` — call it after `backward()` — launches what is left (a parameter
    without a gradient on this rank counts as zeros), waits, and writes the averaged gradients back; a parameter
    used by another rank only receives that rank's gradient, one used by no rank keeps grad None (a per-parameter
    'used' flag travels with each bucket).

        reducer = OverlappedGradientAllReduce(model.parameters())
        loss.backward(); reducer.finish(); optimizer.step()

    NCCL runs the collectives on its own stream (torch's ProcessGroup synchronises it with the producer stream);
    gloo (CPU tests) behaves the same through the returned work handles."""

    def __init__(self, params, bucket_bytes=32 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.average = average
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._reset()

    def _reset(self):
        n = len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * n
        self._flat = [None] * n
        self._next = 0  # buckets [0, _next) have been launched

    def _launch(self, i):
        flat = _pack(self.buckets[i])
        self._flat[i] = flat
        self._work[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True) if dist.is_initialized() else None

    def _on_grad(self, p):
        self._pending[self._bucket_of[id(p)]] -= 1
        while self._next < len(self.buckets) and self._pending[self._next] == 0:  # strictly in bucket order
            self._launch(self._next)
            self._next += 1

    def finish(self):
        """Launch the buckets that are still open (in order), wait for all of them, write the (averaged) sums
        back.  Returns the number of collectives of this step."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        for i, bucket in enumerate(self.buckets):
            if self._work[i] is not None:
                self._work[i].wait()
            _unpack(bucket, self._flat[i], 1.0 / world if (self.average and world > 1) else 1.0)
        n = len(self.buckets)
        self._reset()
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
