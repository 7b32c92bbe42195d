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
    flag = torch.tensor([bad], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
