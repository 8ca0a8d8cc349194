"""bench.py's reference arm on the CPU: the JSON contract the driver reads (one line on stdout, same metric /
unit / config as the GPU arm, impl = reference, cpu_baseline describing the run, zero-copy e2e)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'arnoldi_steps_per_s' and d['unit'] == 'Arnoldi steps/s'
    assert d['higher_is_better'] is True and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert 'workload' in d['config'] and 'k_dim=100' in d['config']['workload']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d['gpu_launches'] == 0
