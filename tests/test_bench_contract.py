"""bench.py's one-line JSON contract: the reference arm on the CPU (no GPU needed), the native arm on a small batch (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

from oracle_lib import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
             'cpu_baseline', 'e2e', 'gpu_launches']


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, 'exactly one JSON line on stdout'
    return json.loads(lines[0])


@pytest.mark.skipif(not have_ref(), reason='oracle/_ref not built')
def test_reference_arm_prints_the_contract_line():
    d = _run(['--impl', 'reference', '--steps', '2', '--warmup', '1'], 300)
    for k in BASE_KEYS + ['impl']:
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'mobiclip_frames_per_sec_400x240' and d['unit'] == 'frames/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['value'] > 0 and d['steps'] == 2
    assert d['config']['workload'] == 'moflex_400x240' and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


@pytest.mark.gpu
def test_native_arm_prints_the_contract_line():
    d = _run(['--steps', '3', '--warmup', '3', '--streams', '64', '--repeats', '3', '--cpu-seconds', '0.3'], 600)
    for k in BASE_KEYS + ['roofline', 'roofline_bgra', 'clocks', 'extra', 'e2e_variants']:
        assert k in d, k
    assert d['metric'] == 'mobiclip_frames_per_sec_400x240' and d['n_gpus'] == 1 and d['steps'] == 3 and d['scaling'] == 'weak'
    assert d['value'] > 0 and abs(d['value'] - 64 / (d['ms_per_step'] * 1e-3)) / d['value'] < 1e-6
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and r['peak'] > 0 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['kernel'] == 'k_inter_chunk' and (r['traffic'] is None or r['traffic'] > 0)
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] == 64 * 400 * 240 * 4
    assert e['value'] < d['value']            # copies and the host parse are inside its timed region
    assert d['gpu_launches'] > 0 and d['cpu_baseline']['cores'] >= 1
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    x = d['extra']
    assert x['config2_pframes_256x192']['frames_per_s'] > 0 and x['config3_single_stream_400x240']['latency_ms']['p50'] > 0
