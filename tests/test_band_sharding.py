"""Host-side logic of the N > 1 path (band sharding + final flux reduction) with world_size-2 gloo on CPU.
The solver is a deterministic stand-in with the interface of tenstream_b200.pprts (no GPU here); the GPU version of
this test is tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tenstream_b200.parallel import BandShardedIntegrator, shard_bands


def test_shard_bands_partition():
    for nb in (1, 2, 7, 112, 252):
        for world in (1, 2, 3, 4, 8):
            parts = [list(shard_bands(nb, world, r)) for r in range(world)]
            flat = [b for p in parts for b in p]
            assert flat == list(range(nb))                       # disjoint, ordered, complete
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


class FakeSolver:
    def __init__(self, shape):
        self.shape = shape
        self.acc = None
        self.cur = None

    def reset_accumulator(self):
        self.acc = [np.zeros(self.shape) for _ in range(4)]

    def accumulate_result(self, uid, weight):
        for q in range(4):
            self.acc[q] += weight * self.cur[q]

    def get_accumulated(self):
        return tuple(a.copy() for a in self.acc)


class FakeApi:
    @staticmethod
    def set_optical_properties(s, albedo, kabs, ksca, g, planck=None, planck_srfc=None):
        s.k = kabs

    @staticmethod
    def solve_pprts(s, lthermal, lsolar, edirTOA, uid):
        s.cur = [np.full(s.shape, (q + 1) * uid + edirTOA) * s.k for q in range(4)]


def make_bands(n):
    rng = np.random.default_rng(0)
    return [dict(albedo=0.1, kabs=rng.random((3, 2, 2)), ksca=None, g=None, lsolar=b % 2 == 0, edirTOA=float(b),
                 uid=b + 1, weight=1.0 / (b + 1)) for b in range(n)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bands = make_bands(7)
    integ = BandShardedIntegrator(FakeApi, FakeSolver((3, 2, 2)), world, rank, reduce="torch")
    out = integ.integrate(bands)
    q.put((rank, [o.copy() for o in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_band_sharding_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = BandShardedIntegrator(FakeApi, FakeSolver((3, 2, 2)), 1, 0).integrate(make_bands(7))
    for r in (0, 1):
        for a, b in zip(res[r], serial):
            assert np.allclose(a, b, rtol=1e-14, atol=0)
