"""Multi-GPU parity (element-partitioned mesh, NCCL all-reduce + interface exchange).
Needs >= 2 GPUs on the box; on a single-GPU box the N > 1 logic is covered by the gloo tests."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize('p2p', ['nccl', 'nvlink_p2p'])
@pytest.mark.parametrize('world', [2, 4])
def test_multirank_parity(world, p2p, lib):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs, box has {torch.cuda.device_count()}')
    env = dict(os.environ, NSB_TEST_P2P='1' if p2p == 'nvlink_p2p' else '0')
    port = 29600 + world + (10 if p2p == 'nvlink_p2p' else 0)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(port), str(ROOT / 'tests' / 'multirank_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
