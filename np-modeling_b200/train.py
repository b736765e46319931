"""Trainer — drop-in for the reference's `train` module (train.py:13-46), made batch-sharded
data parallel.

Single process: identical to the reference loop (forward chain → loss → `loss(backprop=True)` →
reverse chain with `optimizer_`), printing `Step:` / `Loss:` each step.  Inputs arrive as host
arrays and are copied host→device every step; the loss is read back every step.

Data parallel (one process per GPU, `torch.distributed` initialised with NCCL — or gloo for the
CPU tests of this logic): rank r trains on rows [r*B/n, (r+1)*B/n) of `inputs`/`targets`.  Parameters
are broadcast from rank 0 after the lazy initialisation of step 0; gradients, which the optimizer keeps
in a flat arena filled in backward order, are all-reduced (SUM) block by block — each block on a side
stream as soon as the backward pass has produced its last gradient, so the exchange overlaps the rest of
backward; what is left is reduced right before the fused update — and scaled by 1/n when the loss is a
mean over the local shard (MSELoss) or by 1 when it is a sum (CrossEntropyLoss) — so the update equals the single-process one on the full batch
(SURVEY.md §8e).  The printed loss is the global one.
"""
import logging
import os
from typing import Optional, Sequence

import numpy as np
import torch

import loss
import optimizer
from layers import layer
from npm_b200 import device
from npm_b200 import dist as npm_dist

_world = npm_dist.world


def iter_parameters(obj, _seen=None):
    """Yield (owner, attribute) for every parameter reachable from a layer (depth-first, attribute
    order = creation order, identical on every rank)."""
    if _seen is None:
        _seen = set()
    if id(obj) in _seen:
        return
    _seen.add(id(obj))
    if isinstance(obj, (list, tuple)):
        for o in obj:
            yield from iter_parameters(o, _seen)
        return
    if not isinstance(obj, layer.Layer):
        return
    for name, value in list(vars(obj).items()):
        if isinstance(value, (layer.Layer, list, tuple)):
            yield from iter_parameters(value, _seen)
        elif isinstance(obj, layer.StatefulLayer) and name in getattr(obj, 'PARAMETERS', _DEFAULT_PARAMS):
            if isinstance(value, (device.DeviceArray, np.ndarray)):
                yield obj, name


_DEFAULT_PARAMS = ('_w', '_b', '_gamma', '_beta', '_wq', '_wk', '_wv', '_wo', '_bq', '_bk', '_bv', '_bo')


class Trainer:
    def __init__(self,
                 layers: Sequence[layer.Layer],
                 loss_: Optional[loss.Loss] = None,
                 verbose: bool = True,
                 shard_inputs: bool = True,
                 cuda_graph: bool = False):
        """`verbose=False` silences the reference's Step/Loss prints; `shard_inputs=False` says the
        arrays passed to train()/eval() already are this rank's shard.  Extension over the reference:
        `inputs` may be a tuple — it is splatted into the first layer (e.g. `(q, kv)` for a
        DecoderStack), and a tuple gradient is reduced to its first element between layers.
        `cuda_graph=True` (single process, SGDOptimizer, no stochastic layer, `verbose=False`): after two eager steps of
        a `train()` call the whole step — forward chain, loss, reverse chain, fused update — is captured once into a
        CUDA graph and replayed for the remaining steps, one launch per step instead of one per kernel (a launch-bound
        model such as cfg1's MLP is otherwise bound by ~15 us of host work per kernel); anything that does not
        qualify, or whose capture fails, runs the eager loop."""
        self._layers = layers
        self._loss = loss_ or loss.MSELoss()
        self._verbose = verbose
        self._shard_inputs = shard_inputs
        self._synced = False
        self._cuda_graph = cuda_graph
        self._graphs = {}          # (input shapes, optimizer) -> captured step
        self._pinned = {}
        self._prefetched = {}      # key -> (host object, device arrays, event) uploaded ahead of time on the copy stream
        self._copy_stream = None

    # ---- host → device staging ------------------------------------------------------------
    def _shard(self, a):
        rank, world = _world()
        if isinstance(a, tuple):
            return tuple(self._shard(x) for x in a)
        if world == 1 or not self._shard_inputs or isinstance(a, device.DeviceArray):
            return a
        return npm_dist.shard_rows(a, rank, world)

    def prefetch(self, inputs, targets) -> None:
        """Start the host->device copies of the NEXT `train()` / `eval()` call's inputs on a copy stream, so that they
        overlap the step that is still running (an input pipeline calls this right after handing the current batch to
        `train`).  The next call recognises the same host objects and waits for the copies instead of issuing them on
        the compute stream — at 8 GPUs the eight 100 MB uploads per step otherwise sit in the critical path."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        for key, host in (('inputs', inputs), ('targets', targets)):
            with torch.cuda.stream(self._copy_stream):
                dev = self._to_device(key, self._shard(host), _no_prefetch=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._prefetched[key] = (host, dev, ev)      # keyed on the object the caller will pass again

    def _take_prefetched(self, key, a):
        hit = self._prefetched.get(key)
        if hit is None:
            return None

        def same(x, y):
            if isinstance(x, tuple):
                return isinstance(y, tuple) and len(x) == len(y) and all(same(p, q) for p, q in zip(x, y))
            return x is y
        host, dev, ev = hit
        if not same(host, a):
            return None
        del self._prefetched[key]
        torch.cuda.current_stream().wait_event(ev)
        for t in (dev if isinstance(dev, tuple) else (dev,)):
            if isinstance(t, device.DeviceArray):
                t.t.record_stream(torch.cuda.current_stream())     # allocated on the copy stream, consumed on this one
        return dev

    def _to_device(self, key, a, _no_prefetch: bool = False, orig=None):
        """Copy one step's host input into HBM through a reusable pinned staging buffer (`orig`: the object the caller
        passed before sharding — what `prefetch` was given)."""
        if not _no_prefetch and '.' not in key:
            dev = self._take_prefetched(key, a if orig is None else orig)
            if dev is not None:
                return dev
        if isinstance(a, tuple):
            return tuple(self._to_device(f'{key}.{i}', x, True) for i, x in enumerate(a))
        if isinstance(a, device.DeviceArray):
            return a
        if isinstance(a, torch.Tensor) and a.is_pinned() and a.dtype == torch.float32:
            return device.from_pinned(a)        # already page-locked: one async H2D copy
        a = np.asarray(a)
        if not torch.cuda.is_available():
            raise RuntimeError('np-modeling_b200 needs a CUDA device; there is no CPU fallback')
        # Two pinned staging buffers per input, used alternately, each guarded by the event of its last H2D copy: the
        # copy is asynchronous and queued behind ~1000 kernels while the host runs ahead, so rewriting a staging buffer
        # whose copy has not executed yet would upload the NEXT batch's bytes for the current step
        # (`for xb, yb in data: trainer.train(xb, yb, 1, opt)`).
        slot = self._pinned.get(key)
        if slot is None or tuple(slot['bufs'][0].shape) != a.shape:
            slot = {'bufs': [torch.empty(a.shape, dtype=torch.float32).pin_memory() for _ in range(2)],
                    'events': [None, None], 'next': 0}
            self._pinned[key] = slot
        i = slot['next']
        slot['next'] = i ^ 1
        if slot['events'][i] is not None:
            slot['events'][i].synchronize()
        pin = slot['bufs'][i]
        pin.numpy()[...] = a
        out = device.from_pinned(pin)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        slot['events'][i] = ev
        return out

    # ---- data-parallel plumbing -----------------------------------------------------------
    def _broadcast_parameters(self):
        rank, world = _world()
        if world == 1 or self._synced:
            return
        params = [owner._p(name) for owner, name in iter_parameters(list(self._layers))]
        npm_dist.broadcast_from_rank0([p.t for p in params])
        for p in params:
            p.touched()                 # written behind the arrays' back
        self._synced = True

    def _install_grad_sync(self, optimizer_):
        rank, world = _world()
        if world == 1 or not isinstance(optimizer_, optimizer._FusedOptimizer):
            return
        optimizer_.grad_scale = npm_dist.grad_scale(getattr(self._loss, 'dp_mean', True), world)

        reducer = npm_dist.AsyncBucketReducer()

        def bucket_ready(opt, b):
            # every gradient of arena block b exists: all-reduce it on the side stream while backward continues
            blk = opt._arena.blocks[b]
            reducer.reduce(b, blk[0][:blk[1]])

        def sync(opt):
            done = reducer.finish()
            npm_dist.allreduce_sum([blk[0][:blk[1]] for b, blk in enumerate(opt._arena.blocks) if blk[1] > 0 and b not in done])
            # gradients that did not come from the arena (user-supplied buffers)
            arena_ptrs = [(b[0].data_ptr(), b[0].data_ptr() + b[0].numel() * 4) for b in opt._arena.blocks]
            npm_dist.allreduce_sum([g.t for _, _, g in opt._pending
                                    if not any(lo <= g.ptr < hi for lo, hi in arena_ptrs)])

        optimizer_.grad_sync = sync
        optimizer_.bucket_ready = None if os.environ.get('NPM_DP_NO_OVERLAP') else bucket_ready

    def _global_loss(self, l):
        rank, world = _world()
        if world == 1:
            return l
        t = l._t if isinstance(l, device.DeviceScalar) else torch.tensor([float(l)], device=device._device())
        return device.DeviceScalar(device.DeviceArray(npm_dist.global_loss(t, getattr(self._loss, 'dp_mean', True))))

    def _forward(self, y):
        for i, layer_ in enumerate(self._layers):
            y = layer_(*y) if (i == 0 and isinstance(y, tuple)) else layer_(y)
        return y

    # ---- the reference loop (train.py:20-39) ------------------------------------------------
    def train(self, inputs, targets, steps: int, optimizer_: optimizer.Optimizer) -> None:
        orig_inputs, orig_targets = inputs, targets
        inputs, targets = self._shard(inputs), self._shard(targets)
        self._install_grad_sync(optimizer_)
        fused = isinstance(optimizer_, optimizer.Optimizer)
        if _world()[1] > 1 and not self._synced:
            # parameters are created lazily by the first forward (layers/layer.py:33-35): run one
            # untrained forward so they exist, then make every rank start from rank 0's values
            self._forward(self._to_device('inputs', inputs))
            self._broadcast_parameters()

        i = 0
        while i < steps:
            if self._graph_ready(optimizer_, i, steps):
                done = self._train_graphed(inputs, targets, steps - i, optimizer_, orig_inputs, orig_targets)
                if done:
                    return
                self._cuda_graph = False       # capture failed: eager from here on
            if self._verbose:
                print('Step: ', i)
            i += 1

            logging.info('Running forward pass')
            y = self._to_device('inputs', inputs, orig=orig_inputs)
            t = self._to_device('targets', targets, orig=orig_targets)
            y = self._forward(y)
            l = self._loss(y, t)

            logging.info('Running backward pass')
            dy = self._loss(backprop=True)
            # One bracket around the whole reverse chain: every layer's updates are applied by a
            # single fused kernel (after one gradient all-reduce when data parallel).
            if fused:
                optimizer_._enter()
            try:
                for layer_ in reversed(self._layers):
                    dy = layer_(dy, backprop=True, optimizer_=optimizer_)
                    if isinstance(dy, tuple):
                        dy = dy[0]
            finally:
                if fused:
                    optimizer_._exit()
            self.last_loss = self._global_loss(l)
            if self._verbose:
                # printed after the backward pass is enqueued so the host never stalls the GPU
                # mid-step; the text is the reference's (train.py:32)
                print('Loss: ', self.last_loss)

    # ---- CUDA-graph replay of a whole step (opt-in, see __init__) --------------------------------
    def _graph_ready(self, optimizer_, i, steps) -> bool:
        if not self._cuda_graph or self._verbose or _world()[1] > 1 or not isinstance(optimizer_, optimizer.SGDOptimizer):
            return False
        if optimizer_.grad_sync is not None or optimizer_.bucket_ready is not None:
            return False
        return i >= 2 and steps - i >= 2          # two eager steps first: lazy init, allocator pools, kernel attributes

    def _one_step(self, y, t, optimizer_):
        y = self._forward(y)
        l = self._loss(y, t)
        dy = self._loss(backprop=True)
        optimizer_._enter()
        try:
            for layer_ in reversed(self._layers):
                dy = layer_(dy, backprop=True, optimizer_=optimizer_)
                if isinstance(dy, tuple):
                    dy = dy[0]
        finally:
            optimizer_._exit()
        return l

    def _train_graphed(self, inputs, targets, steps, optimizer_, orig_inputs, orig_targets) -> bool:
        from layers import normalizations
        as_tuple = lambda v: v if isinstance(v, tuple) else (v,)
        first_in = self._to_device('inputs', inputs, orig=orig_inputs)
        first_t = self._to_device('targets', targets, orig=orig_targets)
        # a captured step is tied to the buffers it was captured with: input shapes, the optimizer (its gradient arena)
        # and every parameter's storage (an attribute rebound by the user gets a new buffer, and with it a new capture)
        key = (tuple(a.shape for a in as_tuple(first_in)), first_t.shape, id(optimizer_),
               tuple(owner._p(name).ptr for owner, name in iter_parameters(list(self._layers))))
        entry = self._graphs.get(key)
        if entry is None:
            # static input buffers: the graph reads these addresses on every replay
            stat_in = tuple(device.empty(a.shape) for a in as_tuple(first_in))
            stat_t = device.empty(first_t.shape)
            for dst, src in zip(stat_in + (stat_t,), as_tuple(first_in) + (first_t,)):
                dst.t.copy_(src.t)
            before = normalizations._philox['offset']
            graph = torch.cuda.CUDAGraph()
            try:
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):
                    l = self._one_step(stat_in if isinstance(first_in, tuple) else stat_in[0], stat_t, optimizer_)
            except Exception as exc:                      # noqa: BLE001 — any capture problem means: stay eager
                logging.warning('Trainer: CUDA-graph capture failed (%s); running eagerly', exc)
                torch.cuda.synchronize()
                return False
            if normalizations._philox['offset'] != before or not isinstance(l, device.DeviceScalar):
                # a stochastic layer baked one dropout mask into the graph, or the loss is not a device value: not replayable
                logging.warning('Trainer: the step is not replayable as a CUDA graph; running eagerly')
                return False
            entry = self._graphs[key] = (graph, stat_in, stat_t, l._t)
            # capture does not execute: the captured step still has to run for this iteration
        graph, stat_in, stat_t, loss_t = entry
        for k in range(steps):
            if k > 0:
                ins = as_tuple(self._to_device('inputs', inputs, orig=orig_inputs))
                tgt = self._to_device('targets', targets, orig=orig_targets)
            else:
                ins, tgt = as_tuple(first_in), first_t
            for dst, src in zip(stat_in + (stat_t,), ins + (tgt,)):
                if dst.ptr != src.ptr:
                    dst.t.copy_(src.t, non_blocking=True)
            graph.replay()
            self.last_loss = device.DeviceScalar(device.DeviceArray(loss_t.clone()))
        return True

    # ---- checkpoint / resume (the reference has none) -------------------------
    def save_checkpoint(self, path: str, optimizer_: Optional[optimizer.Optimizer] = None) -> None:
        """Parameters (in `iter_parameters` order), the Adam state (step count and both moments per parameter) and the
        dropout stream position → one .npz.  Rank 0 writes; parameters are replicated under data parallel."""
        from layers import normalizations
        torch.cuda.synchronize()
        out = {}
        params = list(iter_parameters(list(self._layers)))
        for i, (owner, name) in enumerate(params):
            out[f'p{i}'] = np.asarray(owner._p(name))
            if isinstance(optimizer_, optimizer.AdamOptimizer):
                ident = f'{id(owner)}.{name}'
                if ident in optimizer_._momentums:
                    out[f'm{i}'] = np.asarray(optimizer_._momentums[ident])
                    out[f'v{i}'] = np.asarray(optimizer_._velocities[ident])
                    out[f't{i}'] = np.int64(optimizer_._steps.get(ident, 1))
        out['n_params'] = np.int64(len(params))
        out['dropout'] = np.array([normalizations._philox['seed'], normalizations._philox['offset']], dtype=np.uint64)
        if _world()[0] == 0:
            np.savez(path, **out)

    def load_checkpoint(self, path: str, optimizer_: Optional[optimizer.Optimizer] = None, example_inputs=None) -> None:
        """Restore what `save_checkpoint` wrote.  Parameters are created lazily by the first forward
        (layers/layer.py:33-35), so pass `example_inputs` (one batch) unless the layers have already been called."""
        from layers import normalizations
        if example_inputs is not None:
            self._forward(self._to_device('inputs', self._shard(example_inputs)))
        with np.load(path if str(path).endswith('.npz') else str(path) + '.npz') as z:
            params = list(iter_parameters(list(self._layers)))
            assert int(z['n_params']) == len(params), f"checkpoint has {int(z['n_params'])} parameters, the model {len(params)}"
            for i, (owner, name) in enumerate(params):
                p = owner._p(name)
                assert p.shape == z[f'p{i}'].shape, f'{name}: {p.shape} vs {z[f"p{i}"].shape}'
                p.copy_from(z[f'p{i}'])                       # in place: aliases and packed blocks stay valid
                if isinstance(optimizer_, optimizer.AdamOptimizer) and f'm{i}' in z.files:
                    ident = f'{id(owner)}.{name}'
                    m, v = optimizer_._moments(ident, p)
                    m.copy_from(z[f'm{i}'])
                    v.copy_from(z[f'v{i}'])
                    optimizer_._steps[ident] = int(z[f't{i}'])
            seed, offset = (int(x) for x in z['dropout'])
            normalizations.set_dropout_seed(seed, offset)
        self._synced = True      # every rank loaded the same values

    def eval(self, inputs, targets) -> None:
        inputs, targets = self._shard(inputs), self._shard(targets)
        y = self._to_device('inputs', inputs)
        t = self._to_device('targets', targets)
        y = self._forward(y)
        l = self._global_loss(self._loss(y, t))
        self.last_loss = l
        if self._verbose:
            print('Loss: ', l)
