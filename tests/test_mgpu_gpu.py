'''Multi-GPU (z-slab) parity under pytest: launches tests/mgpu_worker.py with torch.distributed.run on 2, 4 and 8
ranks (one per GPU) and asserts that every rank's owned CSR rows / RHS are bit-identical to the oracle and that the
distributed Newton updates agree with the pinned SuperLU solve to 1e-8 -- lid-driven cavity and Rayleigh-Benard, ragged
slabs with at least two planes per rank (the reference's pattern: /root/reference/tests/test_PETSc.py:195-286).
Skipped when fewer devices are visible.'''
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    sys.path.insert(0, ROOT)
    from transiflow_b200 import _lib
    return _lib.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 4, 8])
def test_z_slab_ranks_match_the_oracle(world):
    have = _device_count()
    if have < world:
        pytest.skip('%d GPUs visible, %d needed' % (have, world))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'mgpu_worker.py')]
    env = dict(os.environ, OMP_NUM_THREADS='2')
    run = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (run.stdout + '\n' + run.stderr)[-6000:]
    assert run.returncode == 0, tail
    assert run.stdout.count('ALL OK') == world, tail
    assert 'MISMATCH' not in run.stdout, tail
