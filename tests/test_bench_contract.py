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


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (final code of round 2, 1/2/4/8 GPUs) carry every key of the driver's
    contract, a green parity block, and a roofline whose numbers are consistent with each other."""
    import json
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    peaks = json.loads((root / 'MEASURED_PEAKS.json').read_text()) if (root / 'MEASURED_PEAKS.json').exists() else None
    for n in (1, 2, 4, 8):
        d = json.loads((root / 'profiles' / f'bench_r02_final_{n}gpu.json').read_text())
        for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                  'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'clocks', 'e2e', 'gpu_launches', 'parity', 'build_id'):
            assert k in d, (n, k)
        assert d['n_gpus'] == n and d['metric'] == 'arnoldi_steps_per_s' and d['dtype'] == 'f64' and d['higher_is_better']
        assert 'workload' in d['config'] and 'model' not in d['config']
        assert d['vs_baseline'] is None                      # BASELINE.md publishes no number for this metric
        r = d['roofline']
        assert r['bound'] == 'hbm' and r['unit'] == 'GB/s'
        assert abs(r['frac'] - r['achieved'] / r['peak']) <= 1e-3
        if peaks:
            assert abs(r['peak'] - peaks['hbm_gbs']) <= 1e-6 * r['peak']
        assert abs(r['achieved'] * 1e9 * r['avg_launch_ms'] * 1e-3 - r['algorithmic_bytes_per_launch']) <= 2e-3 * r['algorithmic_bytes_per_launch']
        e = d['e2e']
        assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and e['value'] < d['value']
        assert d['parity']['ok'] and d['parity']['orth'] < 1e-10 and d['parity']['H_replicated'] == 0.0
        assert d['gpu_launches'] > 0 and d['clocks']['sm_mhz'] > 0
        assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
        if n == 1:
            c = d['cpu_baseline']
            assert c['kind'] == 'port' and c['cores'] >= 1 and c['value'] > 0 and 'full mesh' in c['sample']
            assert r['traffic'] is not None and abs(r['traffic'] / r['algorithmic_bytes_per_launch'] - 1.0) < 0.05
