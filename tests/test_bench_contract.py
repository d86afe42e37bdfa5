"""bench.py's JSON contract: the committed line of the last GPU run carries every key the driver reads, and the
reference arm (CPU, runs anywhere) prints exactly one JSON line on stdout with the same metric / unit / config."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
             'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _last_line(path):
    return json.loads(path.read_text().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _last_line(sorted((ROOT / 'profiles').glob('r*_bench_cfg1.json'))[-1])
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline'} <= set(d)
    assert d['metric'] == 'train sessions/sec' and d['unit'] == 'sessions/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['scaling'] == 'weak' and d['data'] == 'synthetic' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'configs[1]' in d['config']['workload'] and 'l2' in d['config']
    assert abs(d['value'] - d['config']['B'] * 1e3 / d['ms_per_step']) <= 2e-3 * d['value']      # value = B / step time
    assert d['warmup'] >= 3 and d['gpu_launches'] > 0
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    e = d['e2e']
    assert {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'} <= set(e)
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and e['value'] < d['value']
    r = d['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r) and r['bound'] in ('hbm', 'tensor')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-3
    c = d['cpu_baseline']
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(c) and c['kind'] in ('port', 'reference') and c['cores'] >= 1


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, 'stdout must hold exactly the JSON line'
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d)
    assert d['metric'] == 'train sessions/sec' and d['unit'] == 'sessions/s' and d['value'] > 0
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['kind'] in ('port', 'reference')
    gpu = _last_line(sorted((ROOT / 'profiles').glob('r*_bench_cfg1.json'))[-1])
    assert d['config']['workload'] == gpu['config']['workload']                                  # both arms: same workload


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches both arms the same way for N > 1: rank 0 alone times the CPU path, the other ranks exit 0."""
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29617', str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
