"""tools/dp_check.py — data-parallel equivalence on N GPUs (run under torchrun, one rank per GPU).

A 3-layer DecoderStack is trained for 3 Adam steps (a) batch-sharded across the ranks, gradients all-reduced bucket
by bucket overlapped with backward, and (b) with every rank holding the FULL batch (the mean of identical gradients =
the single-process update).  The parameters after (a) and (b) must agree to fp32 round-off, with and
without the overlap (NPM_DP_NO_OVERLAP=1).  Prints DP_CHECK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
sys.path.insert(0, ROOT)


def run(shard, seed=0):
    import loss
    import optimizer
    import train
    from layers.adapters import DecoderStack
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(seed)                     # identical on every rank
    b, s, d, h, f = 2 * world, 128, 256, 4, 512
    q = rng.standard_normal((b, s, d)).astype(np.float32)
    kv = rng.standard_normal((b, s, d)).astype(np.float32)
    t = rng.standard_normal((b, s, d)).astype(np.float32)
    np.random.seed(1234)                                  # identical lazy initialisation on every rank
    stack = DecoderStack(3, h, f, True, 0.0)
    tr = train.Trainer([stack], loss.MSELoss(), verbose=False, shard_inputs=shard)
    adam = optimizer.AdamOptimizer(learning_rate=1e-3)
    # scale the unscaled reference init so that three layers stay in range
    stack(q[:1], kv[:1])                                   # lazy initialisation
    for owner, name in train.iter_parameters([stack]):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            setattr(owner, name, (v / np.sqrt(v.shape[-1] if v.ndim > 1 else 1.0) * 0.5).astype(np.float32))
    tr.train((q, kv), t, 3, adam)
    torch.cuda.synchronize()
    return {f'{i}.{name}': np.asarray(getattr(owner, name)).copy()
            for i, (owner, name) in enumerate(train.iter_parameters([stack]))}, float(tr.last_loss)


def main():
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl')
    import npm_b200
    npm_b200.set_precision(os.environ.get('NPM_DP_CHECK_PRECISION', 'bf16x3'))
    sharded, loss_s = run(True)
    full, loss_f = run(False)
    worst = 0.0
    for k in full:
        err = np.abs(sharded[k] - full[k]).max() / (np.abs(full[k]).max() + 1e-12)
        worst = max(worst, err)
    ok = worst < 2e-4 and abs(loss_s - loss_f) < 1e-4 * abs(loss_f)
    # every rank must hold identical parameters
    for k in sorted(sharded)[:6]:
        t = torch.from_numpy(sharded[k]).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and bool(torch.equal(lo, hi))
    if dist.get_rank() == 0:
        print(f'precision={npm_b200.get_precision()} world={dist.get_world_size()} overlap={"off" if os.environ.get("NPM_DP_NO_OVERLAP") else "on"} '
              f'max rel param diff sharded vs full-batch = {worst:.3e}, loss {loss_s:.6f} vs {loss_f:.6f}')
        print('DP_CHECK_OK' if ok else 'DP_CHECK_FAILED', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
