'''Host-side logic of the z-slab partition, including a world_size-2 gloo run on CPU.'''
import os
import socket
import subprocess
import sys

import pytest

from transiflow_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('nz,world', [(128, 1), (128, 8), (13, 4), (7, 7), (10, 3)])
def test_slabs_tile_the_domain(nz, world):
    planes = []
    for r in range(world):
        k0, k1 = parallel.slab_range(nz, world, r)
        assert k1 > k0
        planes += list(range(k0, k1))
    assert planes == list(range(nz))
    sizes = [parallel.slab_range(nz, world, r)[1] - parallel.slab_range(nz, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_slab_errors():
    with pytest.raises(ValueError):
        parallel.slab_range(3, 4, 0)
    with pytest.raises(ValueError):
        parallel.slab_range(8, 2, 2)


def test_owned_rows_are_contiguous():
    r = [parallel.owned_rows(5, 4, 4, *parallel.slab_range(9, 3, k)) for k in range(3)]
    assert r[0][0] == 0 and r[-1][1] == 5 * 4 * 9 * 4
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from transiflow_b200 import parallel
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
uid = parallel.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
assert uid == bytes(range(128)), uid
k0, k1 = parallel.slab_range(10, world, rank)
import torch
t = torch.tensor([float(k1 - k0)])
dist.all_reduce(t)
assert t.item() == 10.0
print('rank', rank, 'ok', k0, k1)
"""


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('ok') == 2
