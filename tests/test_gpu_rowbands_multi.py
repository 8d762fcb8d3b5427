"""One field split into row bands over ALL the GPUs of the box (2, 4 or 8), halos read over NVLink peer memory:
forward and backward of the 't' warp and the 's' splat against the single-GPU result of the whole field (SURVEY.md 8e,
"single oversized field").  Self-skips on a box with one GPU; the host logic of the same code is covered by the
world-size-2/3 gloo tests, the kernels by the band-emulation tests."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least 2 GPUs')
def test_row_bands_over_all_gpus_match_the_single_gpu_result():
    g = max(k for k in (2, 4, 8) if k <= torch.cuda.device_count())
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(g),
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tools', 'rowband_bench.py'),
           '2160', '3840']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout[out.stdout.index('{'):])
    assert res['n_gpus'] == g
    assert all(res['parity'].values()), res['parity']
