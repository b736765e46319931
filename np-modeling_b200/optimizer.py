"""Optimizers — drop-in for the reference's `optimizer` module (optimizer.py:12-69).

Same surface: `Optimizer.update(obj, attribute, gradient)`, `update_variable(identifier,
variable, gradient)`, `SGDOptimizer(learning_rate)`, `AdamOptimizer(learning_rate, beta1, beta2,
epsilon)` (epsilon inside the sqrt, per-identifier step count starting at 1).

B200 design: parameters, gradients and Adam moments are device buffers; an update is not applied
tensor by tensor but queued and flushed as ONE multi-tensor kernel (npm_sgd_multi /
npm_adam_multi) when the outermost `Layer.__call__(..., backprop=True)` returns — every gradient
of a backward pass is computed from pre-update weights in the reference too (each parameter's
update is the last use of that parameter, optimizer.py:13-18 called at the end of each
backward), so deferring is result-identical.  Gradients live in a flat arena so that
data-parallel training all-reduces a few large contiguous blocks (see train.py).
"""
import abc
import dataclasses
import os

import numpy as np
import torch

from npm_b200 import device
from npm_b200._lib import OPT_CHUNK, C, TensorEntry


class Optimizer(metaclass=abc.ABCMeta):
    def update(self, obj: object, attribute: str, gradient) -> None:
        identifier = f'{id(obj)}.{attribute}'
        variable = getattr(obj, attribute)
        variable = self.update_variable(identifier, variable, gradient)
        setattr(obj, attribute, variable)

    @abc.abstractmethod
    def update_variable(self, identifier: str, variable, gradient):
        pass

    # ---- B200 extensions (no-ops for user-defined optimizers) ---------------
    def grad_buffer(self, obj: object, attribute: str, shape):
        """Device buffer a layer should write the gradient of `obj.attribute` into."""
        return device.empty(shape)

    def grad_buffer_pack(self, obj: object, attributes, shape):
        """One `[len(attributes), *shape]` buffer whose leading-axis slices are the gradients of
        `obj.<attribute>` (parameters a layer keeps adjacent so one GEMM produces all their gradients)."""
        return device.empty((len(attributes),) + tuple(int(s) for s in shape))

    def _enter(self):
        pass

    def _exit(self):
        pass

    def flush(self):
        pass


class _Arena:
    """Bump allocator over a few large fp32 blocks: gradients end up contiguous (in first-use,
    i.e. backward, order), so a data-parallel all-reduce touches whole blocks."""

    ALIGN = 64            # floats (256 B)
    # Largest block = largest all-reduce bucket.  The arena is filled in backward order, so the LAST block holds the
    # first layers' gradients and its all-reduce cannot overlap any compute.  A/B on 8 x B200 (NPM_DP_BUCKET_MB,
    # profiles/r02_bench_8gpu.txt): 64 MiB buckets 93.3 ms/step, 256 MiB 92.5 ms/step — the exposed tail shrinks, the
    # number of NCCL launches that contend with the persistent tcgen05 kernels for SMs grows; 256 MiB stays.
    MAX_BLOCK = (int(os.environ.get('NPM_DP_BUCKET_MB', '256')) << 20) // 4

    def __init__(self):
        self.blocks = []   # [tensor, used]
        self.total = 0
        self.last_block = -1
        self.version = 0   # bumped by every allocation (a bucket schedule is valid for one version)

    def alloc(self, shape):
        n = int(np.prod(shape)) if len(shape) else 1
        need = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if not self.blocks or self.blocks[-1][1] + need > self.blocks[-1][0].numel():
            size = max(need, min(self.MAX_BLOCK, max(1 << 20, 2 * self.total)))
            self.blocks.append([device.zeros(size).t, 0])
        blk = self.blocks[-1]
        view = blk[0][blk[1]:blk[1] + n].view(*shape)
        blk[1] += need
        self.total += need
        self.last_block = len(self.blocks) - 1
        self.version += 1
        return device.DeviceArray(view)

    def used_views(self):
        return [b[0][:b[1]] for b in self.blocks if b[1] > 0]


class _FusedOptimizer(Optimizer):
    """Queue + flush machinery shared by SGD and Adam."""

    def _init_fused(self):
        self._pending = []        # (identifier, variable, gradient)
        self._depth = 0
        self._grads = {}          # identifier -> persistent grad buffer
        self._grad_packs = {}     # (id(obj), attributes, shape) -> [n, *shape] buffer whose slices are in _grads
        self._arena = _Arena()
        self._tables = {}         # key -> (device table, n_chunks, keepalive)
        self.grad_sync = None     # callable(optimizer) run before a flush applies (data parallel)
        self.grad_scale = 1.0     # multiplies every gradient at apply time
        self.bucket_ready = None  # callable(optimizer, block index): every gradient of an arena block has been produced
        self._grad_block = {}     # identifier -> arena block index
        self._tracker = None      # npm_b200.dist.BucketTracker, rebuilt when the arena grows
        self._tracker_version = -1

    def grad_buffer(self, obj, attribute, shape):
        identifier = f'{id(obj)}.{attribute}'
        if any(p[0] == identifier for p in self._pending):
            # the same parameter a second time inside one bracket (a layer applied twice): its queued update still
            # points at this persistent buffer, so apply it before the buffer is handed out to be overwritten
            self.flush()
        buf = self._grads.get(identifier)
        shape = tuple(int(s) for s in shape)
        if buf is None or buf.shape != shape:
            buf = self._arena.alloc(shape)
            self._grads[identifier] = buf
            self._grad_block[identifier] = self._arena.last_block
        return buf

    def grad_buffer_pack(self, obj, attributes, shape):
        shape = tuple(int(s) for s in shape)
        key = (id(obj), tuple(attributes), shape)
        if any(p[0] in {f'{id(obj)}.{a}' for a in attributes} for p in self._pending):
            self.flush()
        pack = self._grad_packs.get(key)
        if pack is None:
            pack = self._arena.alloc((len(attributes),) + shape)
            self._grad_packs[key] = pack
            for i, attribute in enumerate(attributes):
                self._grads[f'{id(obj)}.{attribute}'] = pack[i]
                self._grad_block[f'{id(obj)}.{attribute}'] = self._arena.last_block
        return pack

    def _enter(self):
        self._depth += 1

    def _exit(self):
        self._depth -= 1
        if self._depth == 0:
            self.flush()

    def update_variable(self, identifier, variable, gradient):
        variable = device.param(variable)
        gradient = device.asdevice(gradient)
        assert variable.size == gradient.size, f'{identifier}: {variable.shape} vs {gradient.shape}'
        if any(p[0] == identifier for p in self._pending):
            self.flush()   # same parameter twice in one backward: keep the reference's sequential semantics
        self._pending.append((identifier, variable, gradient))
        if self.bucket_ready is not None and self._depth > 0:
            self._mark_bucket(identifier, gradient)
        if self._depth == 0:
            self.flush()
        return variable

    def _mark_bucket(self, identifier, gradient):
        """Data parallel: tell `bucket_ready` when the last gradient of an arena block arrives.  Only once the arena
        layout is stable (no allocation since the previous flush) and only for gradients that live in the arena."""
        if self._tracker is None or self._tracker_version != self._arena.version:
            return
        own = self._grads.get(identifier)
        if own is None or own.ptr != gradient.ptr:
            return
        done = self._tracker.mark(identifier)
        if done is not None:
            self.bucket_ready(self, done)

    def _table(self, entries, moments):
        planes = [device.weight_planes_of(v) for (_, v, _) in entries]
        key = tuple((v.ptr, g.ptr, v.size, pl) for (_, v, g), pl in zip(entries, planes))
        hit = self._tables.get(key)
        if hit is not None:
            self._mark_planes(entries, planes)
            return hit
        arr = (TensorEntry * len(entries))()
        chunks = 0
        for i, (ident, v, g) in enumerate(entries):
            arr[i].param = v.ptr
            arr[i].grad = g.ptr
            if planes[i] is not None:
                arr[i].planes, arr[i].plane_stride = planes[i]
            if moments is not None:
                m, s = moments(ident, v)
                arr[i].m, arr[i].v = m.ptr, s.ptr
            arr[i].numel = v.size
            arr[i].chunk_begin = chunks
            chunks += (v.size + OPT_CHUNK - 1) // OPT_CHUNK
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = host.to(device._device())
        hit = (dev, chunks)
        if len(self._tables) > 64:
            self._tables.clear()
        self._tables[key] = hit
        self._mark_planes(entries, planes)
        return hit

    @staticmethod
    def _mark_planes(entries, planes):
        """The update kernel about to be launched with this table rewrites the split-bf16 image of every weight whose
        planes are in it (the launch is stream-ordered before anything that can read them)."""
        for (_, v, _), pl in zip(entries, planes):
            if pl is not None:
                device.mark_weight_planes_fresh(v)

    def flush(self):
        if not self._pending:
            return
        if self.grad_sync is not None:
            self.grad_sync(self)
        if self.bucket_ready is not None:
            if self._tracker_version != self._arena.version:
                from npm_b200.dist import BucketTracker
                self._tracker = BucketTracker(self._grad_block)
                self._tracker_version = self._arena.version
            self._tracker.begin_step()
        pending, self._pending = self._pending, []
        self._apply(pending)

    @abc.abstractmethod
    def _apply(self, pending):
        pass


class SGDOptimizer(_FusedOptimizer):
    def __init__(self, learning_rate: float) -> None:
        self._learning_rate = learning_rate
        self._init_fused()

    def _apply(self, pending):
        table, chunks = self._table(pending, None)
        C.npm_sgd_multi(table.data_ptr(), len(pending), chunks, float(self._learning_rate),
                        float(self.grad_scale), device.stream())



@dataclasses.dataclass
class AdamOptimizerConfig:
    learning_rate: float
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-7

    def __post_init__(self, *args, **kwargs):
        self._steps = {}
        self._momentums = {}
        self._velocities = {}


class AdamOptimizer(AdamOptimizerConfig, _FusedOptimizer):
    def __post_init__(self, *args, **kwargs):
        super().__post_init__(*args, **kwargs)
        self._init_fused()

    def _moments(self, identifier, variable):
        m = self._momentums.get(identifier)
        if m is None or m.size != variable.size:
            m = device.zeros(variable.shape)
            v = device.zeros(variable.shape)
            self._momentums[identifier] = m
            self._velocities[identifier] = v
        return m, self._velocities[identifier]

    def _apply(self, pending):
        # group by step count t (optimizer.py:53: per-identifier, starting at 1); in practice one group
        groups = {}
        for item in pending:
            groups.setdefault(self._steps.get(item[0], 1), []).append(item)
        for t, entries in groups.items():
            table, chunks = self._table(entries, self._moments)
            C.npm_adam_multi(table.data_ptr(), len(entries), chunks, float(self.learning_rate), float(self.beta1),
                             float(self.beta2), float(self.epsilon), int(t), float(self.grad_scale), device.stream())
            for ident, _, _ in entries:
                self._steps[ident] = t + 1
