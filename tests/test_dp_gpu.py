"""Data-parallel Trainer on 2 GPUs (NCCL): batch-sharded training with bucketed, backward-overlapped gradient all-reduce
equals training on the full batch (tools/dp_check.py).  Skipped on boxes with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('overlap', ['on', 'off'])
@pytest.mark.timeout(600)
def test_sharded_training_equals_full_batch_on_two_gpus(overlap):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ)
    if overlap == 'off':
        env['NPM_DP_NO_OVERLAP'] = '1'
    port = 29500 + (os.getpid() % 2000)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'tools', 'dp_check.py')], env=env, capture_output=True, text=True, timeout=540)
    assert 'DP_CHECK_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
