"""Data-parallel PPOLearner on 2 GPUs (NCCL): spawned with torch.distributed.run; skipped with < 2 devices."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('env', [{'DP_MODE': 'clip'}, {'DP_MODE': 'adapt', 'DP_BIGLR': '1'}, {'DP_MODE': 'clip', 'SB200_DP_GRAPH': '0'}])
def test_data_parallel_learner_matches_global_batch_oracle(env):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    e = dict(os.environ)
    e.update(env)
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dp_check.py')],
                       env=e, capture_output=True, text=True, timeout=240)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and out.count('DP_OK') == 2, out[-3000:]
