"""Host-side data-parallel logic on CPU with gloo, world_size 2: batch sharding,
SUM all-reduce of the gradient arena + loss-dependent scaling reproduce the single-process
full-batch gradient / update; the Philox dropout ranges tile the global tensor."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from npm_b200 import dist as D
        from oracle import np_oracle as O
        from oracle import philox
        assert D.world() == (rank, world)
        rng = np.random.default_rng(0)                      # same data on every rank
        x = rng.standard_normal((8, 6))
        t = rng.standard_normal((8, 4))
        onehot = np.eye(4)[rng.integers(0, 4, size=8)]
        w = rng.standard_normal((6, 4)) * 0.3
        b = rng.standard_normal(4) * 0.1
        xs, ts, oh = (D.shard_rows(a, rank, world) for a in (x, t, onehot))
        assert xs.shape == (8 // world, 6)

        res = {}
        for name, is_mean in (('mse', True), ('ce', False)):
            # local gradient on this rank's shard (oracle = the reference's arithmetic)
            if name == 'mse':
                y = O.linear_fwd(xs, w, b)
                dy = O.mse_bwd(y, ts)
                local_loss = O.mse_fwd(y, ts)
            else:
                y, z = O.dense_fwd(xs, w, b, activation='softmax')
                dy = O.softmax_bwd(y, O.ce_bwd(y, oh))
                local_loss = O.ce_fwd(y, oh)
            _, dw, db = O.linear_bwd(xs, w, dy)
            arena = torch.from_numpy(np.concatenate([dw.ravel(), db.ravel()]).copy())   # one flat block
            D.allreduce_sum([arena])
            scale = D.grad_scale(is_mean, world)
            g = arena.numpy() * scale
            loss = D.global_loss(torch.tensor([local_loss], dtype=torch.float64), is_mean).item()
            # single-process reference on the full batch
            if name == 'mse':
                yf = O.linear_fwd(x, w, b)
                dyf = O.mse_bwd(yf, t)
                full_loss = O.mse_fwd(yf, t)
            else:
                yf, _ = O.dense_fwd(x, w, b, activation='softmax')
                dyf = O.softmax_bwd(yf, O.ce_bwd(yf, onehot))
                full_loss = O.ce_fwd(yf, onehot)
            _, dwf, dbf = O.linear_bwd(x, w, dyf)
            np.testing.assert_allclose(g, np.concatenate([dwf.ravel(), dbf.ravel()]), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(loss, full_loss, rtol=1e-12)
            # the Adam step every rank then takes equals the single-process one
            wa, _, _ = O.adam_step(w, g[:24].reshape(6, 4), np.zeros((6, 4)), np.zeros((6, 4)), 1, 1e-2)
            wf, _, _ = O.adam_step(w, dwf, np.zeros((6, 4)), np.zeros((6, 4)), 1, 1e-2)
            np.testing.assert_allclose(wa, wf, rtol=1e-12)
            res[name] = loss

        # parameter broadcast: every rank ends with rank 0's values
        p = torch.full((5,), float(rank + 1))
        D.broadcast_from_rank0([p])
        assert torch.equal(p, torch.ones(5))

        # dropout: the per-rank Philox ranges tile the global tensor's range
        n_local, base, seed = 40, 1000, 77
        off, nxt = D.dropout_range(n_local, base, rank, world)
        mine = philox.dropout_mask(n_local, 0.75, seed, off)
        whole = philox.dropout_mask(n_local * world, 0.75, seed, base)
        assert np.array_equal(mine, whole[rank * n_local:(rank + 1) * n_local])
        assert nxt == base + world * n_local
        # bucketed, overlapped all-reduce: buckets are reduced as their last gradient arrives, the rest at the end;
        # the result equals one all-reduce of everything
        bucket_of = {'l2.w': 0, 'l2.b': 0, 'l1.w': 1, 'l1.b': 1, 'l0.w': 2}
        tracker = D.BucketTracker(bucket_of)
        blocks = [torch.full((4,), float(rank + 1 + b)) for b in range(3)]
        reducer = D.AsyncBucketReducer()
        fired = []
        for ident in ('l2.w', 'l2.b', 'l2.b', 'l1.w', 'unknown', 'l1.b'):       # l0.w never arrives in this pass
            b = tracker.mark(ident)
            if b is not None:
                fired.append(b)
                reducer.reduce(b, blocks[b])
        assert fired == [0, 1]
        done = reducer.finish()
        assert done == {0, 1}
        D.allreduce_sum([blk for b, blk in enumerate(blocks) if b not in done])
        for b in range(3):
            assert torch.equal(blocks[b], torch.full((4,), float(sum(r + 1 + b for r in range(world)))))
        tracker.begin_step()
        assert tracker.mark('l0.w') == 2 and tracker.mark('l0.w') is None
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_data_parallel_logic_gloo_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0, f'worker exited with {p.exitcode}'
    got = dict(out.get(timeout=10) for _ in range(2))
    assert got[0] == got[1]                      # both ranks report the same global losses


def test_shard_rows_rejects_ragged_batches():
    sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
    from npm_b200 import dist as D
    with pytest.raises(ValueError):
        D.shard_rows(np.zeros((7, 2)), 0, 2)
    assert D.shard_rows(np.arange(8).reshape(4, 2), 1, 2).tolist() == [[4, 5], [6, 7]]
    assert D.grad_scale(True, 4) == 0.25 and D.grad_scale(False, 4) == 1.0
    assert D.dropout_range(10, 0, 1, 2) == (12, 24)
