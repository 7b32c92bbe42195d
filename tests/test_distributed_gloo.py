"""The multi-slab host layer (SlabDriver) on CPU: world sizes 2 and 3 under torch.distributed/gloo
with the numpy test double of the per-slab backend.  Results must equal the single-domain CPU
checker bit for bit (halo pairing, neighbour wrap with two ranks, uneven slabs, all-to-all splits)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
NAMES = ("ux_e", "uy_e", "ux_i", "uy_i", "ux_n", "uy_n", "T_e", "T_i", "T_n", "rho_e", "rho_i", "rho_n", "rho_q", "Ex", "Ey", "phi")


def _worker(rank, world, port, NX, NY, poisson, steps, out_path):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    import torch.distributed as dist
    import plbm_b200 as P
    from slab_double import SlabDouble
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert P.slab_of(NY, rank, world) == (SlabDouble(NX, NY, rank, world, poisson).y0, SlabDouble(NX, NY, rank, world, poisson).nyl)
        b = SlabDouble(NX, NY, rank, world, poisson)
        drv = P.SlabDriver(b)
        drv.step(steps, want_fields=True)
        full = drv.gather_fields(NAMES)
        if rank == 0:
            np.savez(out_path, **full)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,NX,NY,poisson,steps", [(2, 24, 24, "fft", 6), (2, 20, 28, "none", 5), (3, 20, 20, "fft", 5)])
def test_slab_driver_matches_single_domain(tmp_path, oracle, plbm, world, NX, NY, poisson, steps):
    out = tmp_path / "gathered.npz"
    mp.spawn(_worker, args=(world, _free_port(), NX, NY, poisson, steps, str(out)), nprocs=world, join=True)
    got = np.load(out)
    o = oracle.PortOracle(NX, NY, poisson=poisson)
    o.step(steps)
    want = o.fields()
    for n in NAMES:
        eq = (got[n] == want[n]) | (np.isnan(got[n]) & np.isnan(want[n]))
        assert eq.all(), f"{n}: {int((~eq).sum())} values differ (world={world})"
