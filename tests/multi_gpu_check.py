"""Run under torchrun with N >= 2 GPUs: slab-decomposed run on CUDA (NCCL) against the single-domain
CPU checker.  Exit code 0 = bit-identical.  Used by tests/test_gpu_multi.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    import plbm_b200 as P
    from oracle import oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bad = 0
    require_peer = os.environ.get("PLBM_REQUIRE_PEER", "0") == "1"
    # both transposes: all-to-alls through NCCL, and the column pass working in peer memory (auto: falls back when
    # CUDA IPC / peer access is unavailable, unless PLBM_REQUIRE_PEER=1)
    for NX, poisson, steps, peer in ((64, "fft", 10, False), (64, "fft", 10, None), (60, "fft", 6, None), (256, "fft", 12, None), (1536, "fft", 3, None),
                                     (48, "none", 5, None)):
        b = P.CudaSlabBackend(NX, NX, rank, world, poisson=poisson, device=local)
        drv = P.SlabDriver(b, peer_memory=(True if (require_peer and peer is None and poisson == "fft") else peer))
        drv.step(steps, want_fields=True)
        if drv.peer:
            b.peer_check()
        b.sync()
        full = drv.gather_fields(P.FIELD_NAMES)
        if rank == 0:
            o = O.PortOracle(NX, NX, poisson=poisson)
            o.step(steps)
            want = o.fields()
            for n in P.FIELD_NAMES:
                if not O.same_bits(full[n], want[n]):
                    bad += 1
                    print(f"MISMATCH world={world} {NX}x{NX}/{poisson}: {n} max|diff|={np.nanmax(np.abs(full[n] - want[n])):.3e}", flush=True)
            print(f"checked {NX}x{NX}/{poisson}, {steps} steps on {world} GPUs, transposes through "
                  f"{'peer memory' if drv.peer else 'all-to-all'}: {'ok' if not bad else 'FAILED'}", flush=True)
        drv.close()
        b.close()
    # the pipelined peer step (Poisson solve of step t on a second stream beside K1 of step t, plbm_step_peer with PLBM_PEER_PIPELINE=1)
    os.environ["PLBM_PEER_PIPELINE"] = "1"
    for NX, steps in ((64, 9), (256, 7)):
        b = P.CudaSlabBackend(NX, NX, rank, world, poisson="fft", device=local)
        drv = P.SlabDriver(b, peer_memory=(True if require_peer else None))
        if drv.peer:
            drv.step(steps - 2)
            drv.step(2, want_fields=True)
            b.sync()
            full = drv.gather_fields(P.FIELD_NAMES)
            if rank == 0:
                o = O.PortOracle(NX, NX, poisson="fft")
                o.step(steps)
                want = o.fields()
                nb = sum(not O.same_bits(full[n], want[n]) for n in P.FIELD_NAMES)
                bad += nb
                print(f"checked pipelined peer step {NX}x{NX}/fft, {steps} steps on {world} GPUs: {'ok' if not nb else 'FAILED'}", flush=True)
        drv.close()
        b.close()
    os.environ["PLBM_PEER_PIPELINE"] = "0"
    # an uploaded (not initialised) state on slabs: every rank uploads its rows of one global random state; the halo rows come
    # from the upload itself and refresh_halos() must leave them alone (it used to overwrite them with stale rows)
    for peer in (None, False):
        NX, steps = 96, 3
        rng = np.random.default_rng(4321)
        f = rng.uniform(0.05, 1.0, size=(3, NX, NX, 9)); g = rng.uniform(0.01, 0.5, size=(3, NX, NX, 9))
        f[1] *= 1800.0; f[2] *= 1e9
        b = P.CudaSlabBackend(NX, NX, rank, world, poisson="fft", device=local, initialize=False)
        drv = P.SlabDriver(b, peer_memory=(True if (require_peer and peer is None) else peer))
        y0, y1 = b.slab_y0[rank], b.slab_y0[rank + 1]
        b.sim.upload_state(f[:, y0:y1], g[:, y0:y1])
        units = O.units_from_si()
        b.sim.set_efield(np.full((y1 - y0, NX), units.Ex_ext), np.full((y1 - y0, NX), units.Ey_ext))
        drv.refresh_halos()
        drv.step(steps, want_fields=True)
        b.sync()
        full = drv.gather_fields(P.FIELD_NAMES)
        if rank == 0:
            o = O.PortOracle(NX, NX, poisson="fft", initialize=False)
            for s in range(3):
                o.f(s)[...] = f[s]; o.g(s)[...] = g[s]
            o.scalar(O.PO_EX)[...] = o.units.Ex_ext; o.scalar(O.PO_EY)[...] = o.units.Ey_ext
            o.step(steps)
            want = o.fields()
            nb = sum(not O.same_bits(full[n], want[n]) for n in P.FIELD_NAMES)
            bad += nb
            print(f"checked uploaded state {NX}x{NX}/fft, {steps} steps on {world} GPUs, transposes through "
                  f"{'peer memory' if drv.peer else 'all-to-all'}: {'ok' if not nb else 'FAILED'}", flush=True)
        drv.close()
        b.close()
    flag = torch.tensor([bad], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
