"""Data-parallel plumbing of the Trainer, kept free of CUDA so the N > 1 logic is testable with gloo
on CPU tensors (tests/test_dist_cpu.py).  One process per GPU; torch.distributed (NCCL over
NVLink 5 / NVSwitch in production) only moves bytes — what is reduced and how it is scaled is
decided here.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(a, rank: int, world_size: int):
    """Rank r trains on rows [r*B/n, (r+1)*B/n) of the global batch."""
    if world_size == 1:
        return a
    n = a.shape[0]
    if n % world_size:
        raise ValueError(f'batch {n} is not divisible by world size {world_size}')
    per = n // world_size
    return a[rank * per:(rank + 1) * per]


def grad_scale(loss_is_mean: bool, world_size: int) -> float:
    """Factor applied to the SUM of per-rank gradients so the update equals the single-process one:
    MSELoss divides by the LOCAL y.size (loss.py:29) → mean of the rank gradients; CrossEntropyLoss
    is an un-normalised sum over the batch (loss.py:36-39) → plain sum."""
    return 1.0 / world_size if loss_is_mean else 1.0


def allreduce_sum(tensors) -> None:
    """SUM all-reduce of each tensor in place (the optimizer's gradient arena blocks: a few large
    contiguous buffers, so NVLink/NVSwitch bandwidth, not launch latency, sets the cost)."""
    if world()[1] == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def global_loss(local: torch.Tensor, loss_is_mean: bool) -> torch.Tensor:
    """Loss of the global batch from the per-rank losses (what train.py:32 prints)."""
    rank, n = world()
    if n == 1:
        return local
    t = local.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if loss_is_mean:
        t /= n
    return t


def broadcast_from_rank0(tensors) -> None:
    if world()[1] == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=0)


def dropout_range(n: int, base: int, rank: int, world_size: int):
    """Philox counter range of one DropOut call on a batch-sharded tensor.

    The global tensor (world_size shards of n elements, rank-major = batch-major) is element
    g = rank*n + i of the stream starting at `base`; every rank advances `base` by the size of the
    GLOBAL tensor, so the union of the per-rank masks is the single-process mask whatever the world
    size.  Sizes are rounded up to 4 so each 128-bit vector is one Philox call.
    Returns (offset for this rank, next base)."""
    n4 = (n + 3) // 4 * 4
    return base + rank * n4, base + world_size * n4


class BucketTracker:
    """Which gradient buckets (arena blocks) are complete during a backward pass.

    The optimizer's gradient arena is filled in first-use = backward order, so bucket b holds the
    gradients of a contiguous run of layers; once every gradient that lives in b has been handed to
    `Optimizer.update` in this pass, b can be all-reduced while the rest of backward still runs
    (SURVEY.md §8e).  Pure host logic: `mark(identifier)` returns the bucket that just became
    complete, or None."""

    def __init__(self, bucket_of):
        self.bucket_of = dict(bucket_of)                  # identifier -> bucket index
        self.expected = {}
        for b in self.bucket_of.values():
            self.expected[b] = self.expected.get(b, 0) + 1
        self.begin_step()

    def begin_step(self):
        self.seen = set()
        self.count = {b: 0 for b in self.expected}

    def mark(self, identifier):
        b = self.bucket_of.get(identifier)
        if b is None or identifier in self.seen:
            return None
        self.seen.add(identifier)
        self.count[b] += 1
        return b if self.count[b] == self.expected[b] else None


class AsyncBucketReducer:
    """SUM all-reduce of buckets as they complete, off the compute stream.

    CUDA tensors: the collective is enqueued on a side stream that first waits for the compute stream's
    work so far (the kernels that wrote the bucket); `finish()` makes the compute stream wait for the side
    stream.  CPU tensors (gloo, tests): asynchronous work handles, waited in `finish()`."""

    def __init__(self):
        self._side = None
        self._handles = []
        self.reduced = set()

    def reduce(self, index, tensor):
        if world()[1] > 1:
            if tensor.is_cuda:
                if self._side is None:
                    self._side = torch.cuda.Stream()
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
            else:
                self._handles.append(dist.all_reduce(tensor, op=dist.ReduceOp.SUM, async_op=True))
        self.reduced.add(index)

    def finish(self):
        for h in self._handles:
            h.wait()
        self._handles = []
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        done, self.reduced = self.reduced, set()
        return done
