"""bench.py contract on the CPU arm (no GPU needed): `--impl reference` times the reference's own CPU implementation
(oracle/_ref, placed by oracle/make_ref.sh when /root/reference is present; else the oracle port) on a bounded sample of
the cfg5 workload and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--seq', '64', '--cpu-seq', '32', '--d-model', '128', '--heads', '2', '--hidden', '256', '--layers', '2'],
                         capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'].startswith('TransformerDecoder') and d['unit'] == 'tokens/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['vs_baseline'] is None
    have_ref = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'layers', 'transformer.py'))
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample'] and d['steps'] == 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_make_ref_recipe_copies_only_into_the_ignored_directory():
    """oracle/make_ref.sh is the committed recipe for oracle/_ref; reference sources never enter the history."""
    with open(os.path.join(ROOT, '.gitignore')) as f:
        assert 'oracle/_ref/' in f.read()
    tracked = subprocess.run(['git', 'ls-files', 'oracle/_ref'], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    assert tracked == ''
    assert os.access(os.path.join(ROOT, 'oracle', 'make_ref.sh'), os.X_OK)
