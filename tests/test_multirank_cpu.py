"""N > 1 host-side logic on CPU: world_size-2 and -3 gloo runs of the partition, the library's
gather-scatter / exchange plans and the all-reduced inner product (no GPU needed)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize('world', [2, 3])
def test_gloo_partitioned_dssum_and_dot(world, lib):
    env = dict(os.environ, OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29700 + world),
           str(ROOT / 'tests' / 'gloo_dssum_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('dssum err=') == world
