"""Band sharding over 2 GPUs with the library's own NCCL all-reduce (run with `gpurun --gpus 2 -- pytest -m gpu`).
Skipped when fewer than 2 devices are visible."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, paths, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    from tenstream_b200 import scenarios
    from tenstream_b200.parallel import BandShardedIntegrator
    cfg = L.small_test_config(310)
    lut = P.Lut(cfg, paths)
    sc = scenarios.cloud_field(Nx=16, Ny=12, Nz=10, dx=100., dz=60., seed=3)
    s = P.init_pprts(0, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                     dz1d=sc["dz1d"], lut=lut, device=rank, rtol=1e-8, atol=1e-30, keep_solutions=False)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    bands = []
    for b in range(5):
        sca = 10.0 ** (-1 + 0.5 * b)
        if b % 2 == 0:
            bands.append(dict(albedo=sc["albedo"], kabs=sc["kabs"] * sca, ksca=sc["ksca"], g=sc["g"], lsolar=True,
                              edirTOA=100.0 * (b + 1), uid=b + 1, weight=1.0))
        else:
            bands.append(dict(albedo=sc["albedo_th"], kabs=sc["kabs_th"] * sca, ksca=sc["ksca_th"], g=sc["g_th"],
                              planck=sc["planck"], planck_srfc=sc["planck_srfc"], lsolar=False, edirTOA=0.0,
                              uid=500 + b, weight=0.5))
    out = BandShardedIntegrator(P, s, world, rank, reduce="nccl").integrate(bands)
    serial = BandShardedIntegrator(P, s, 1, 0).integrate(bands) if rank == 0 else None
    q.put((rank, [np.array(o) for o in out], [np.array(o) for o in serial] if serial else None))
    dist.barrier()
    P.destroy_pprts(s)
    dist.destroy_process_group()


def test_band_sharding_two_gpus(small_lut):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _, paths = small_lut
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, paths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, out, serial = q.get(timeout=600)
        res[r] = (out, serial)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    serial = res[0][1]
    for r in (0, 1):
        for a, b in zip(res[r][0], serial):
            assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()       # same solves, sum order differs by one add
