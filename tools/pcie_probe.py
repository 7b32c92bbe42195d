#!/usr/bin/env python
"""Host-link characterisation for the end-to-end number (bench.py `e2e`): device-to-host copy rate into pinned host memory of
every GPU alone and of all GPUs at once.  The reference's API hands 15 FP64 fields per step to the host, so at N GPUs the e2e
metric is bounded by the AGGREGATE host-memory / PCIe rate of the box, not by the GPUs.

    python tools/pcie_probe.py [--mb 512] > profiles/rN_pcie.json
"""
import argparse
import json
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    nbytes = args.mb << 20
    dev, host, streams = [], [], []
    for d in range(n):
        torch.cuda.set_device(d)
        dev.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"))
        host.append(torch.empty(nbytes, dtype=torch.uint8).pin_memory())
        streams.append(torch.cuda.Stream(device=d))

    def run(devices):
        best = 0.0
        for _ in range(args.reps):
            for d in devices:
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for d in devices:
                with torch.cuda.stream(streams[d]):
                    host[d].copy_(dev[d], non_blocking=True)
            for d in devices:
                streams[d].synchronize()
            dt = time.perf_counter() - t0
            best = max(best, len(devices) * nbytes / dt / 1e9)
        return best

    out = {"gpus": n, "mb_per_copy": args.mb, "cpus": os.cpu_count(),
           "d2h_alone_GBs": [round(run([d]), 2) for d in range(n)],
           "d2h_all_at_once_GBs_aggregate": round(run(list(range(n))), 2)}
    for k in (2, 4):
        if n > k:
            out[f"d2h_{k}_at_once_GBs_aggregate"] = round(run(list(range(k))), 2)
    try:
        out["numa_nodes"] = len([x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")])
    except OSError:
        out["numa_nodes"] = None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
