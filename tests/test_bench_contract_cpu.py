"""bench.py's reference arm (the UNMODIFIED reference decoder staged under oracle/_ref, on the host cores) runs without a GPU: hold
its one JSON line to the driver's contract -- one line on stdout, the keys the driver reads, `impl`, a `cpu_baseline` describing
the run and a zero-copy `e2e`.  The default metric is the training step (BASELINE.json: "decode & train step"), like the GPU arm's."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    cmd = [sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-batch', '1', *extra]
    out = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # exactly ONE line on stdout, and it is JSON
    return json.loads(lines[0])


def test_reference_arm_json_line():
    j = _run()
    assert j['impl'] == 'reference'
    assert j['metric'] == 'images/sec (train)' and j['unit'] == 'images/s' and j['higher_is_better'] is True
    assert j['n_gpus'] == 1 and j['steps'] == 1 and j['scaling'] == 'weak' and j['data'] == 'synthetic'
    assert j['vs_baseline'] is None                                   # BASELINE.md holds no published number for this metric
    assert j['value'] > 0 and abs(j['value'] - 1e3 / j['ms_per_step']) <= 1e-6 * j['value']      # 1 image per step here
    assert 'workload' in j['config'] and 'D-legacy' in j['config']['workload'] and 'model' not in j['config']
    cb = j['cpu_baseline']
    staged = os.path.isfile(os.path.join(REPO, 'oracle', '_ref', 'MANIFEST.json'))
    assert cb['kind'] == ('reference' if staged else 'port') and cb['cores'] >= 1 and cb['unit'] == j['unit'] and cb['value'] == j['value'] and cb['sample']
    e = j['e2e']
    assert e['value'] == j['value'] and e['unit'] == j['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_decode_mode_names_the_decode_metric():
    j = _run('--mode', 'decode')
    assert j['impl'] == 'reference' and j['metric'] == 'images/sec (decode)' and j['value'] > 0


def test_reference_arm_runs_the_staged_reference_not_the_mirror():
    """With oracle/_ref staged, the arm must import the reference's own networks.generator and nothing of vfm_vae_b200.decoder."""
    if not os.path.isfile(os.path.join(REPO, 'oracle', '_ref', 'MANIFEST.json')):
        import pytest
        pytest.skip('reference not staged')
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-batch', '1', '--mode', 'decode'];"
            "import bench; a = bench.parse(); bench._REAL_STDOUT = 2; bench.run_reference_arm(a);"
            "import networks.generator as g; assert '/oracle/_ref/' in g.__file__, g.__file__;"
            "assert 'vfm_vae_b200.decoder' not in sys.modules and 'vfm_vae_b200.plugins' not in sys.modules")
    out = subprocess.run([sys.executable, '-c', code], cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints the reference arm; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    cmd = [sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--cpu-batch', '1']
    out = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
