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


def allreduce_gradients(params, world=None, bucket_bytes=32 << 20, average=True):
    """Sum (or average) the .grad of `params` across ranks with as few collectives as possible: grads are
    flattened into buckets of `bucket_bytes` (sized for launch latency, not link count — NVSwitch gives
    every pair full bandwidth).  Returns the number of collectives issued."""
    if not dist.is_initialized():
        return 0
    world = world or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        calls += 1
        bucket, size = [], 0

    for g in grads:
        nbytes = g.numel() * g.element_size()
        if size + nbytes > bucket_bytes and bucket:
            flush()
        bucket.append(g)
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
    `bucket_bytes` up front, in REVERSE registration order (gradients arrive roughly last layer first); a bucket is
    flattened and all-reduced asynchronously the moment its last gradient has been accumulated
    (`register_post_accumulate_grad_hook`), while autograd keeps running; `finish()` — call it after `backward()` —
    launches whatever is left (parameters that received no gradient in this step count as zeros, so every rank issues
    the same collectives whatever its graph looked like), waits, and writes the averaged gradients back.

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
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _launch(self, i):
        bucket = self.buckets[i]
        ref = next((p for p in bucket if p.grad is not None), bucket[0])
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket]) \
            if len(bucket) > 1 or bucket[0].grad is None else bucket[0].grad.reshape(-1).clone()
        flat = flat.to(ref.dtype)
        self._flat[i] = flat
        self._work[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True) if dist.is_initialized() else None

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def finish(self):
        """Launch the buckets that are still open, wait for all of them, write the (averaged) sums back.  Returns the
        number of collectives of this step."""
        world = dist.get_world_size() if dist.is_initialized() else 1
        for i in range(len(self.buckets)):
            if self._work[i] is None and self._flat[i] is None:
                self._launch(i)
        for i, bucket in enumerate(self.buckets):
            if self._work[i] is not None:
                self._work[i].wait()
            flat = self._flat[i]
            if self.average and world > 1:
                flat /= world
            off = 0
            for p in bucket:
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        n = len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * n
        self._flat = [None] * n
        return n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
