"""CPU checks of the reference arm: oracle/build_ref.py's sourceless copy of the unmodified reference imports and
computes, and `bench.py --impl reference` prints the contract line (same `config` object as our arm, cpu_baseline.kind
"reference" when oracle/_ref exists, "port" otherwise)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import build_ref  # noqa: E402
import ref_shim  # noqa: E402


def test_compiled_reference_imports_and_steps():
    build_ref.build()
    if not build_ref.available():
        import pytest
        pytest.skip('neither /root/reference nor a prebuilt oracle/_ref')
    models = ref_shim.import_reference_models()
    torch.manual_seed(0)
    m = models.MultiDMM(['a', 'b'], [2, 3], h_dim=8, z_dim=4, device=torch.device('cpu'))
    x = {'a': torch.randn(5, 3, 2), 'b': torch.randn(5, 3, 3)}
    x['b'][1, 0] = float('nan')
    mask = torch.ones(5, 3, 1, dtype=torch.bool)
    loss = m.step(x, mask, 1.0, {'a': 1.0, 'b': 1.0}, targets=x, lengths=[5, 5, 5], train_particles=3, match_particles=4)
    loss.backward()
    assert torch.isfinite(loss) and m.trans['fwd'].z_lin.weight.grad is not None


def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c1',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['metric'] == 'bfvi_elbo_fwd_bwd_seq_timesteps_per_sec' and d['value'] > 0
    assert d['unit'] == 'seq-timesteps/s' and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == ('reference' if build_ref.available() else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['e2e']['h2d_bytes_per_step'] == 0
    assert d['config']['workload'].startswith('C1')
