#!/usr/bin/env python
"""bench.py -- throughput of the three-species plasma LBM time step (BASELINE.json metric: MLUPS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--nx NX]

One "step" = one pass of the hot path (fused collide-stream K1 + Poisson solve + field) over the
whole lattice.  One lattice update = one cell advanced one step for all three species, f and g.
Workload at N=1: BASELINE.json configs[2], 2048x2048, FFT Poisson, periodic -- the configuration
the single-GPU roofline number is quoted on (the state, 2 x 1.8 GB, is far larger than L2).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      K1 (fused collide-stream) against the measured HBM copy peak; 880 algorithmic B/update
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/ref_plasma_timing) on the host cores, bounded sample
  e2e           same metric through the public API with HOST buffers: initial state uploaded from pinned
                memory, the 15 visualised fields fetched to pinned memory every step (copy of step t overlapped with step t+1)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K1_BYTES_PER_UPDATE = 880          # 54*8 read + 54*8 write + phi 8 (E = -grad phi is formed in the kernel) + rho_q 8   
STEP_BYTES_PER_UPDATE = 928        # + Poisson passes 3 x (8 + 8); SURVEY.md 8d's 960 minus the E-field sweep that no longer exists
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=0, help="lattice side (default: 2048 at N=1)")
    ap.add_argument("--poisson", default="fft")
    ap.add_argument("--cpu-steps", type=int, default=4, help="time steps of the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (tuning sweeps only)")
    ap.add_argument("--no-dense", action="store_true", help="skip the K1 measurement on an all-dense lattice")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong_8192 / weak_4096 configurations")
    ap.add_argument("--weak-sides", default="friendly", choices=["friendly", "literal"],
                    help="weak_4096 lattice sides: 4096/6144/8192/12288 (default) or the literal roundings 4096/5760/8192/11520")
    ap.add_argument("--parity-steps", type=int, default=8, help="time steps of the parity gate")
    return ap.parse_args()


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs.

    nvidia-smi needs 0.1-0.3 s before its first line, as long as the whole timed region of a default run, so the
    sampling loop is started BEFORE the warm-up (start()); the `with` block only marks the timed window, and the
    summary is taken from the lines that arrived inside it (each line is stamped on arrival; one sampling period
    of slack on either side, and if the window was shorter than that, the nearest line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 20

    def __init__(self, index: int = 0):
        self.index = index
        self.samples = []               # (arrival time, line)
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.PERIOD_MS)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def wait_first(self, timeout: float = 3.0):
        """Block until the sampling loop delivers lines (called after the warm-up, before the timed region)."""
        t0 = time.perf_counter()
        while self.proc and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def __enter__(self):
        if self.proc is None:
            self.start()
        self.wait_first()
        self.t_begin = time.perf_counter()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def __exit__(self, *exc):
        self.t_end = time.perf_counter()
        time.sleep(1.5 * self.PERIOD_MS / 1e3)          # let the line that covers the end of the window arrive
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        slack = self.PERIOD_MS / 1e3
        lo, hi = (self.t_begin or 0.0) - slack, (self.t_end or float("inf")) + slack
        inside = [ln for (ts, ln) in self.samples if lo <= ts <= hi]
        if not inside and self.samples and self.t_begin is not None:
            mid = 0.5 * (self.t_begin + self.t_end)
            inside = [min(self.samples, key=lambda s: abs(s[0] - mid))[1]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def ref_binary():
    return ROOT / "oracle" / "_ref" / "ref_plasma_timing"


def run_cpu_reference(nx: int, steps: int, poisson: str, threads: int):
    """The reference's own CPU implementation of the path (unmodified sources, FFTW replaced by the
    oracle FFT, visualisation disabled).  Returns (MLUPS over the time loop, info)."""
    from oracle import oracle as O
    exe = ref_binary()
    if exe.exists():
        info, _, _ = O.run_reference(nx, nx, steps, poisson=poisson, threads=threads, kind="timing")
        return info["mlups"], "reference", info
    # the reference binary is built where /root/reference is mounted; otherwise time the C restatement
    o = O.PortOracle(nx, nx, poisson=poisson)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    return nx * nx * steps / dt / 1e6, "port", {"loop_s": dt}


def cpu_fields(nx: int, steps: int, poisson: str, threads: int):
    """The 15 visualised fields + phi after `steps` time steps from the reference's initial condition, computed on the host:
    by the UNMODIFIED reference (parity build: reference sources, -ffp-contract=off) where oracle/_ref holds it, else by
    the C restatement.  Checker only -- nothing here is timed as, or shipped in, the product."""
    from oracle import oracle as O
    if O.have_reference("parity"):
        _, dumps, _ = O.run_reference(nx, nx, steps, poisson=poisson, threads=threads, kind="parity", dump_steps=[steps - 1])
        return dumps[steps - 1], "unmodified reference, parity build (oracle/_ref/ref_plasma_parity)"
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    o = O.PortOracle(nx, nx, poisson=poisson)
    o.step(steps)
    want = o.fields()
    o.close()
    return want, "C restatement (oracle/plasma_oracle.c)"


def compare_fields(got: dict, want: dict):
    """Number of fields that are not bit-identical (up to the sign of zero; NaN matches NaN), and the worst field-normalised error."""
    import numpy as np
    from oracle import oracle as O
    bad, worst = [], 0.0
    for n, w in want.items():
        if n not in got:
            continue
        if not O.same_bits(got[n], w):
            bad.append(n)
            with np.errstate(all="ignore"):
                den = float(np.nanmax(np.abs(w))) or 1.0
                worst = max(worst, float(np.nanmax(np.abs(got[n] - w))) / den)
    return bad, worst


def parity_single(sim, nx: int, poisson: str, steps: int):
    """N=1 gate: the benchmark's own workload, `steps` steps from the initial condition, GPU against the host checker."""
    threads = os.cpu_count() or 1
    want, checker = cpu_fields(nx, steps, poisson, threads)
    sim.initialize()
    sim.step(steps, want_fields=True)
    got = sim.fields()
    bad, worst = compare_fields(got, want)
    return {"config": f"{nx}x{nx}, {poisson.upper()} Poisson, periodic (the timed workload)", "steps": steps, "fields": len(want),
            "cells_compared": nx * nx * len(want), "mismatches": len(bad), "mismatched_fields": bad, "worst_normalised_error": worst,
            "bar": "bit-identical (sign of zero aside)", "checker": checker}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    arm_nx = args.nx or (WEAK_SIDES.get(world, 2048) if world > 1 else 2048)      # the lattice the b200 arm runs at this N
    nx = min(arm_nx, 2048)      # bounded sample: the reference keeps ~2.15 kB of host memory per cell (81 GB at 6144^2)
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 32))        # ~1 s per 2048^2 step on 16 host threads: the driver's K (20) is honoured as it is
    warm = 1 if args.warmup > 0 else 0
    if warm:
        run_cpu_reference(min(nx, 256), 2, args.poisson, threads)
    t0 = time.perf_counter()
    mlups, kind, info = run_cpu_reference(nx, steps, args.poisson, threads)
    wall = time.perf_counter() - t0
    sample = (f"{nx}x{nx} lattice, {steps} time steps of the unmodified reference loop (time loop only; constructor "
              f"{info.get('ctor_s', 0):.1f}s excluded), FFTW replaced by oracle/fft_oracle.c, visualisation disabled"
              + ("" if nx == arm_nx else f"; bounded sample of the {arm_nx}x{arm_nx} workload (MLUPS is per cell update; the full lattice "
                                         f"would need {arm_nx * arm_nx * 2150 / 1e9:.0f} GB of host memory in the reference's layout)"))
    line = {"impl": "reference", "metric": "MLUPS (all species)", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": info.get("loop_s", wall) / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(arm_nx, args.poisson, world),
                       "sample_lattice": f"{nx}x{nx}" + ("" if nx == arm_nx else " (bounded sample of the workload: MLUPS is per cell update)")},
            "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        # started as a plain `python bench.py --gpus N`: become the launch the contract describes (one rank per GPU)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import plbm_b200 as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)
    peak, peak_src = hbm_peak()
    if world > 1:
        multi_gpu(args, P, torch, world, rank, local_rank, K, W, peak, peak_src)
        return
    nx = args.nx or 2048
    sampler = ClockSampler(local_rank).start()          # running before the warm-up, see ClockSampler
    sim = P.PlasmaLBM(nx, nx, poisson=args.poisson, device=local_rank)
    sim.step(W)
    sim.sync()
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = {"ms_k1": 0.0, "ms_poisson": 0.0, "launches": 0}
    with sampler as clocks:
        with torch.cuda.stream(ext):
            ev0.record()
        for n in segments(K):
            sim.initialize()                      # restart from the reference's initial condition (inside the timed region)
            seg = sim.step_timed(n)
            for k in t:
                t[k] += seg[k]
            t["launches"] += 5                    # initialise: 1 kernel + 2 fills + 2 memsets
        with torch.cuda.stream(ext):
            ev1.record()
        sim.sync()
        torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    cells = nx * nx
    mlups = cells * K / (ms_total * 1e-3) / 1e6
    k1_ms = t["ms_k1"] / K
    achieved = K1_BYTES_PER_UPDATE * cells / (k1_ms * 1e-3) / 1e9
    line = base_line(args, nx, world, K, W, ms_total, mlups, "configs[2]")
    line["roofline"] = roofline(nx, cells, achieved, peak, peak_src, k1_ms, t["ms_k1"] / ms_total)
    line["clocks"] = clocks.summary()
    line["gpu_launches"] = t["launches"]

    # ---- the lattices BASELINE.json states its scaling targets on, single-GPU points of those curves ----
    if not args.no_extra:
        line["strong_8192"] = single_extra(P, args, local_rank, 8192, "8192x8192 strong scaling (BASELINE.json configs[3])", peak)
        line["weak_4096"] = single_extra(P, args, local_rank, 4096, "4096x4096 = 4096^2 cells per GPU, weak scaling (BASELINE.json configs[4])", peak)

    # ---- K1 on a lattice where all three species are dense in every cell -----------------------------------------------------
    if not args.no_dense:
        line["roofline"]["dense_lattice"] = dense_k1(sim, nx, peak)
    line["config"]["k1_cell_paths"] = ("K1's time per cell depends on the data: in this workload (the reference's own initial condition) electrons and "
                                       "ions fill the central quarter of the lattice, and the cells where both are empty take a shorter, bit-identical copy "
                                       "of the cell code; roofline.frac is the workload's figure, roofline.dense_lattice the figure with every species in every cell")

    # ---- parity gate: the timed workload against the host checker ---------------------------
    if not args.no_parity:
        line["parity"] = parity_single(sim, nx, args.poisson, args.parity_steps)

    # ---- e2e: public API with host buffers -------------------------------------------------
    if not args.no_e2e:
        Ke = max(1, min(args.e2e_steps, K))
        f_host = torch.empty((3, nx, nx, 9), dtype=torch.float64).pin_memory()
        g_host = torch.empty((3, nx, nx, 9), dtype=torch.float64).pin_memory()
        NF = 15                                   # what visualize::UpdateVisualization takes (phi is not part of it)
        out_host = [torch.empty((NF, nx, nx), dtype=torch.float64).pin_memory() for _ in range(2)]
        sim.initialize()                          # e2e starts from the reference's initial condition as well
        f0, g0 = sim.download_state()
        f_host.numpy()[...] = f0
        g_host.numpy()[...] = g0
        del f0, g0
        import ctypes as C
        dp = C.POINTER(C.c_double)
        fa = (dp * 3)(*[C.cast(f_host[s].data_ptr(), dp) for s in range(3)])
        ga = (dp * 3)(*[C.cast(g_host[s].data_ptr(), dp) for s in range(3)])
        outs = [(dp * len(P.FIELD_NAMES))(*([C.cast(o[k].data_ptr(), dp) for k in range(NF)] + [dp()] * (len(P.FIELD_NAMES) - NF)))
                for o in out_host]
        lib = sim.lib
        # host-memory bandwidth of the (virtualised) box fluctuates by 2x between runs: best of three identical passes
        dts = []
        for _rep in range(3):
            sim.sync()
            t0 = time.perf_counter()
            if lib.plbm_upload_state(sim._h, fa, ga) != 0:
                raise SystemExit(lib.plbm_last_error().decode())
            # LBmethod::Run_simulation's loop (12-lb-12-lb_b200/host/plasma.cpp): step t+1 is issued while step t's fields
            # are still crossing PCIe; every step's 15 fields are complete in host memory before the next fetch is started
            for k in range(Ke):
                if lib.plbm_step(sim._h, 1, 1) != 0 or lib.plbm_fetch_wait(sim._h) != 0 or lib.plbm_fetch_begin(sim._h, outs[k & 1]) != 0:
                    raise SystemExit(lib.plbm_last_error().decode())
            if lib.plbm_fetch_wait(sim._h) != 0:
                raise SystemExit(lib.plbm_last_error().decode())
            sim.sync()
            dts.append(time.perf_counter() - t0)
        dt = min(dts)
        line["e2e"] = {"value": cells * Ke / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 6 * 9 * cells * 8 / Ke,
                       "d2h_bytes_per_step": NF * cells * 8, "steps": Ke, "what": E2E_WHAT,
                       "passes_ms_per_step": [round(x / Ke * 1e3, 3) for x in dts]}
        # alternate output path (plbm_frames_begin): the visualiser's 12 CV_32F matrices + 19x9 series are
        # formed on the device, 48 B per cell cross PCIe instead of 120.  Reported beside the headline, never as it.
        fr_host = [torch.empty((12, nx, nx), dtype=torch.float32).pin_memory() for _ in range(2)]
        se_host = [torch.empty((19, 9), dtype=torch.float64).pin_memory() for _ in range(2)]
        fpt = C.POINTER(C.c_float)
        frs = [(fpt * 12)(*[C.cast(h[k].data_ptr(), fpt) for k in range(12)]) for h in fr_host]
        ses = [C.cast(h.data_ptr(), dp) for h in se_host]
        lib.plbm_frames_begin.argtypes = [C.c_void_p, C.POINTER(fpt), dp]
        dtf = []
        for _rep in range(3):
            sim.sync()
            t0 = time.perf_counter()
            if lib.plbm_upload_state(sim._h, fa, ga) != 0:
                raise SystemExit(lib.plbm_last_error().decode())
            for k in range(Ke):
                if lib.plbm_step(sim._h, 1, 1) != 0 or lib.plbm_fetch_wait(sim._h) != 0 or lib.plbm_frames_begin(sim._h, frs[k & 1], ses[k & 1]) != 0:
                    raise SystemExit(lib.plbm_last_error().decode())
            if lib.plbm_fetch_wait(sim._h) != 0:
                raise SystemExit(lib.plbm_last_error().decode())
            sim.sync()
            dtf.append(time.perf_counter() - t0)
        line["e2e_frames"] = {"value": cells * Ke / min(dtf) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 6 * 9 * cells * 8 / Ke,
                              "d2h_bytes_per_step": 12 * cells * 4 + 19 * 9 * 8, "steps": Ke,
                              "what": "same loop through the alternate output path (LBmethod::Run_simulation_frames): per step the 12 float32 matrices "
                                      "and 19x9 sample values the visualiser derives from the fields, formed on the device (bit-identical to the "
                                      "visualiser's own narrowing); needs visualize::UpdateVisualizationFrames, so it is NOT the drop-in headline",
                              "passes_ms_per_step": [round(x / Ke * 1e3, 3) for x in dtf]}
        del f_host, g_host, out_host, fr_host, se_host

    # ---- cpu baseline (bounded sample of the same workload) -------------------------------
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        try:
            v, kind, info = run_cpu_reference(nx, args.cpu_steps, args.poisson, threads)
            line["cpu_baseline"] = {"value": v, "unit": "MLUPS", "cores": threads, "kind": kind,
                                    "sample": f"{nx}x{nx}, {args.cpu_steps} time steps (time loop only), unmodified reference sources, "
                                              f"FFTW replaced by oracle/fft_oracle.c, visualisation disabled"}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": threads, "kind": "unavailable", "sample": str(e)[:200]}
    sim.close()
    print(json.dumps(line), flush=True)
    if line.get("parity", {}).get("mismatches"):
        raise SystemExit("bench.py: the timed path is NOT bit-identical to the host checker (see the line's parity object)")


def dense_k1(sim, nx, peak, steps=12):
    """K1 with electrons and ions dense in EVERY cell.  In the reference's initial condition (the timed workload) they fill only the
    central quarter of the lattice (src/plasma.cpp:128-143) and an empty charged species takes K1's short path (k1_species_empty,
    bit-identical); this is the kernel's figure when no cell can take it: the populations of the central cell copied to all cells."""
    sim.initialize()
    f, g = sim.download_state()
    c = nx // 2
    for s in (0, 1):
        f[s][...] = f[s][c, c, :]
        g[s][...] = g[s][c, c, :]
    sim.upload_state(f, g)
    del f, g
    sim.step(3)
    seg = sim.step_timed(steps)
    sim.sync()
    k1_ms = seg["ms_k1"] / steps
    achieved = K1_BYTES_PER_UPDATE * nx * nx / (k1_ms * 1e-3) / 1e9
    return {"what": "electrons and ions dense in every cell (central cell of the reference's initial condition copied everywhere): "
                    "no cell takes the empty-species path", "steps": steps, "k1_ms_per_launch": k1_ms, "achieved": achieved,
            "frac": achieved / peak, "ms_per_step": seg["ms_total"] / steps}


def single_extra(P, args, device, nx, label, peak):
    """One more lattice on this GPU: 3 untimed + 12 timed steps from the initial condition, device time from the library's events."""
    K, W = 12, 3
    try:
        sim = P.PlasmaLBM(nx, nx, poisson=args.poisson, device=device)
    except P.PlbmError as e:
        return {"workload": label, "unavailable": str(e)[:200]}
    sim.step(W)
    sim.sync()
    sim.initialize()
    seg = sim.step_timed(K)
    sim.sync()
    sim.close()
    cells = nx * nx
    k1_ms = seg["ms_k1"] / K
    return {"workload": label, "value": cells * K / (seg["ms_total"] * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": seg["ms_total"] / K,
            "steps": K, "warmup": W, "k1_ms": k1_ms, "k1_frac_of_hbm_peak": K1_BYTES_PER_UPDATE * cells / (k1_ms * 1e-3) / 1e9 / peak,
            "poisson_ms": seg["ms_poisson"] / K}


SEGMENT = 48


def segments(K):
    """The reference's dynamics overflow at large lattices (its own CPU build reaches NaN at step ~100 at 1024^2 and
    ~85 at 2048^2, DESIGN.md "Divergence of the reference"), and non-finite cells take the kernel's IEEE fallback.  The
    timed steps therefore run in segments that each restart from the reference's initial condition, which keeps every
    timed step in the finite regime; the restart is part of the timed region."""
    out = []
    while K > 0:
        out.append(min(SEGMENT, K))
        K -= out[-1]
    return out


E2E_WHAT = ("plbm_upload_state of the 6 AoS population arrays from pinned host memory (once, amortised over the steps), then per "
            "step the time step + plbm_fetch_begin/plbm_fetch_wait of the 15 visualised fields into pinned host memory, as "
            "LBmethod::Run_simulation hands them to the visualiser (the copy of step t overlaps the kernels of step t+1)")

# weak scaling: constant cells per GPU (2048^2) on square lattices of side ~ sqrt(N), like the reference's own weak-scaling
# runs (build/weak_scalability.py); sides of the form 2^k or 3*2^k keep the spectral solve on radix-4/2 stages (+ one radix-3),
# so cells per GPU are 2048^2 (N=1,4) or 1.125 x 2048^2 (N=2,8)
WEAK_SIDES = {1: 2048, 2: 3072, 4: 4096, 8: 6144}


def workload_name(nx, poisson, world):
    """The same string in both arms, so that the driver can tell they ran the same configuration."""
    cfg = "BASELINE.json configs[2]" if world == 1 else "BASELINE.json configs[4]-style weak scaling, 2048^2 cells per GPU"
    return f"{nx}x{nx} plasma D2Q9 3 species + DDF thermal, {poisson.upper()} Poisson, periodic ({cfg})"


def base_line(args, nx, world, K, W, ms_total, mlups, cfg_name):
    cells = nx * nx
    return {"metric": "MLUPS (all species)", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(nx, args.poisson, world),
                       "initial_condition": "reference Initialize(): e/i block in the central square, neutrals uniform; restarted every "
                                            f"{SEGMENT} timed steps (restart inside the timed region) because the reference's dynamics overflow "
                                            "to NaN after ~85 steps at this size on the CPU as well",
                       "l2": f"state 2 x {54 * 8 * cells / world / 1e9:.2f} GB per GPU streams through HBM every step (inputs larger than L2, no flush needed)",
                       "bytes_per_update_k1": K1_BYTES_PER_UPDATE, "bytes_per_update_step": STEP_BYTES_PER_UPDATE,
                       "hbm_gbs_whole_step": STEP_BYTES_PER_UPDATE * mlups * 1e-3}}


def roofline(nx, cells, achieved, peak, peak_src, k1_ms, share):
    traffic = None
    tfile = ROOT / "profiles" / "k1_traffic.json"       # dram bytes per launch from the committed ncu capture
    if tfile.exists():
        try:
            rec = json.loads(tfile.read_text())
            if rec.get("nx") == nx:
                traffic = rec.get("dram_bytes_per_launch")
        except Exception:
            pass
    return {"bound": "hbm", "kernel": "k1_tma_kernel<false,true>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic,
            "traffic_source": ("committed ncu capture (profiles/k1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch), "
                               "not measured in this run" if traffic is not None else None),
            "peak_source": peak_src, "k1_ms_per_launch": k1_ms, "k1_share_of_step": share}


def parity_multi(args, P, torch, dist, world, rank, local_rank):
    """N>1 gate, run BEFORE the timed region on the same ranks: slab-decomposed runs through BOTH exchange paths (peer memory over
    NVLink, and NCCL messages) against the single-domain host checker, all 15 fields + phi bit for bit.  Cases: a lattice the
    host checker finishes in seconds at the benchmark's slab height class (2048^2), and one whose rows do not divide evenly."""
    cases = [(2048, args.parity_steps // 2 or 1), (200, 6)]
    report, total_bad = [], 0
    for nx, steps in cases:
        want = checker = None
        if rank == 0:
            want, checker = cpu_fields(nx, steps, args.poisson, os.cpu_count() or 1)
        for path in ("peer", "nccl"):
            b = P.CudaSlabBackend(nx, nx, rank, world, poisson=args.poisson, device=local_rank)
            try:
                drv = P.SlabDriver(b, peer_memory=(path == "peer"))
            except P.PlbmError as e:
                b.close()
                report.append({"lattice": f"{nx}x{nx}", "steps": steps, "path": path, "skipped": str(e)[:120]})
                continue
            drv.step(steps, want_fields=True)
            b.sync()
            full = drv.gather_fields(P.FIELD_NAMES)
            rows = [b.slab_y0[r + 1] - b.slab_y0[r] for r in range(world)]
            bad, worst = ([], 0.0)
            if rank == 0:
                bad, worst = compare_fields(full, want)
            report.append({"lattice": f"{nx}x{nx}", "steps": steps, "path": path, "slab_rows": rows, "mismatches": len(bad),
                           "mismatched_fields": bad, "worst_normalised_error": worst})
            total_bad += len(bad)
            drv.close()
            b.close()
    flag = torch.tensor([total_bad], device=f"cuda:{local_rank}")
    dist.broadcast(flag, 0)
    return {"config": f"{world} y-slabs, {args.poisson.upper()} Poisson, periodic; single-domain host checker", "fields": len(P.FIELD_NAMES),
            "mismatches": int(flag.item()), "cases": report, "bar": "bit-identical (sign of zero aside)", "checker": checker}


def timed_slabs(args, P, torch, dist, world, rank, local_rank, nx, K, W, sampler=None):
    """One timed run of nx x nx on `world` y-slabs: W untimed steps, then K steps (restarted from the initial condition every
    SEGMENT steps) between barriers; device time by CUDA events on the library's stream, max over ranks.  Returns a dict with the
    whole-job MLUPS, ms per step, K1's time per launch, the per-stage device times of the slowest-K1 rank and the open driver."""
    b = P.CudaSlabBackend(nx, nx, rank, world, poisson=args.poisson, device=local_rank)
    drv = P.SlabDriver(b)
    drv.step(min(W, SEGMENT))
    b.sync(); torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_events, stage_ms = [], {}

    def timed_region():
        with b.stream_context():
            ev0.record()
        for n in segments(K):
            b.sim.initialize()
            drv.refresh_halos()
            if drv.peer:
                drv.step(n, stage_ms=stage_ms)
            else:
                drv.step(n, timing=k1_events)
        with b.stream_context():
            ev1.record()
        b.sync(); torch.cuda.synchronize()

    if sampler is not None:
        with sampler:
            timed_region()
    else:
        timed_region()
    dev = f"cuda:{local_rank}"
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if drv.peer:
        timed = max(stage_ms.get("timed_steps", 0), 1)
        mine = [stage_ms.get(n, 0.0) / timed for n in b.STAGES]
    else:
        mine = [sum(a.elapsed_time(c) for a, c in k1_events) / max(len(k1_events), 1)] + [0.0] * (len(b.STAGES) - 1)
    st = torch.tensor(mine, dtype=torch.float64, device=dev)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)              # per stage: the slowest rank
    ms_total = float(ms.item())
    cells = nx * nx
    out = {"nx": nx, "ms_total": ms_total, "value": cells * K / (ms_total * 1e-3) / 1e6, "ms_per_step": ms_total / K,
           "k1_ms": float(st[0].item()), "local_cells": nx * (b.slab_y0[rank + 1] - b.slab_y0[rank]), "peer": bool(drv.peer),
           "stage_us": ({n: round(float(st[k].item()) * 1e3, 1) for k, n in enumerate(b.STAGES)} if drv.peer else None),
           "drv": drv, "b": b}
    return out


def extra_config(args, P, torch, dist, world, rank, local_rank, nx, label, peak):
    """A second lattice on the same ranks (BASELINE.json configs[3] / configs[4]): fewer steps, same timing rules."""
    K, W = 12, 3
    try:
        r = timed_slabs(args, P, torch, dist, world, rank, local_rank, nx, K, W)
    except P.PlbmError as e:
        return {"workload": label, "unavailable": str(e)[:200]}
    r["drv"].close(); r["b"].close()
    ach = K1_BYTES_PER_UPDATE * r["local_cells"] / (r["k1_ms"] * 1e-3) / 1e9
    return {"workload": label, "value": r["value"], "unit": "MLUPS", "ms_per_step": r["ms_per_step"], "steps": K, "warmup": W,
            "k1_ms": r["k1_ms"], "k1_frac_of_hbm_peak": ach / peak, "stage_us": r["stage_us"],
            "transposes": "peer memory" if r["peer"] else "NCCL all-to-all"}


# BASELINE.json configs[4]: weak scaling at 4096^2 cells per GPU on square lattices of side ~ sqrt(N) (build/weak_scalability.py).
# Sides 2^k and 3*2^k (as for the headline: 1 or 1.125 x 4096^2 cells per GPU) keep the spectral solve on radix-4/2 passes plus one
# radix-3 pass.  The literal sqrt(N)*4096 roundings 5760 = 2^7*45 and 11520 = 2^8*45 run and are bit-exact too (--weak-sides literal),
# but their radix-3/3/5 passes make the solve 2.6x slower per cell (profiles/r2_summary.md: N=8, 11520^2: 5.79 ms/step, 2.1 ms of it Poisson).
WEAK4096_SIDES = {1: 4096, 2: 6144, 4: 8192, 8: 12288}
WEAK4096_SIDES_LITERAL = {1: 4096, 2: 5760, 4: 8192, 8: 11520}


def multi_gpu(args, P, torch, world, rank, local_rank, K, W, peak, peak_src):
    """Slab decomposition over `world` GPUs: halo rows, the spectral transposes and the phi rows through peer memory (NVLink),
    or NCCL send/recv + two all-to-alls per step where peer memory is unavailable."""
    import torch.distributed as dist
    nx = args.nx or WEAK_SIDES.get(world) or (int(2048 * world ** 0.5) // 64 * 64)
    parity = None if args.no_parity else parity_multi(args, P, torch, dist, world, rank, local_rank)
    sampler = ClockSampler(local_rank).start()
    r = timed_slabs(args, P, torch, dist, world, rank, local_rank, nx, K, W, sampler)
    drv, b = r["drv"], r["b"]
    cells = nx * nx
    line = base_line(args, nx, world, K, W, r["ms_total"], r["value"], "configs[4]-style weak scaling, 2048^2 cells per GPU")
    achieved = K1_BYTES_PER_UPDATE * r["local_cells"] / (r["k1_ms"] * 1e-3) / 1e9
    line["roofline"] = roofline(nx, cells, achieved, peak, peak_src, r["k1_ms"], r["k1_ms"] * K / r["ms_total"])
    line["roofline"]["note"] = "per GPU (slowest rank)"
    if r["stage_us"]:
        line["stage_us"] = r["stage_us"]
    transposes = ("column pass of the spectral solve reads/writes every slab's half spectrum in place through peer memory (NVLink), 3 flag barriers, "
                  "the whole sequence issued by one library call per batch of steps (plbm_step_peer)"
                  if drv.peer else "2 all-to-all transposes of the half spectrum (NCCL)")
    line["config"]["decomposition"] = f"{world} y-slabs; per step 18 halo rows per side + {transposes} + 1 phi row per side"
    line["clocks"] = sampler.summary()
    # launches inside the timed region, counted from the sequence the library issues per step (plbm_step_peer: K1, halo push, P1,
    # barrier, unpack, P2, barrier, P3, barrier; pipelined: + charge pull and a fourth barrier) plus the 5 of every restart;
    # without peer memory K1, pack, P1, P2, P3, unpack
    pipelined = os.environ.get("PLBM_PEER_PIPELINE", "0") == "1"
    line["gpu_launches"] = K * ((11 if pipelined else 9) if drv.peer else 6) + 5 * len(segments(K))
    if parity is not None:
        line["parity"] = parity
    if not args.no_e2e:
        import ctypes as C
        Ke = max(1, min(args.e2e_steps, K))
        nyl = b.slab_y0[rank + 1] - b.slab_y0[rank]
        NF = 15
        out_host = [torch.empty((NF, nyl, nx), dtype=torch.float64).pin_memory() for _ in range(2)]
        dp = C.POINTER(C.c_double)
        outs = [(dp * len(P.FIELD_NAMES))(*([C.cast(o[k].data_ptr(), dp) for k in range(NF)] + [dp()] * (len(P.FIELD_NAMES) - NF)))
                for o in out_host]
        b.sync(); dist.barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            drv.step(1, want_fields=True)
            if b.lib.plbm_fetch_wait(b.sim._h) != 0 or b.lib.plbm_fetch_begin(b.sim._h, outs[k & 1]) != 0:
                raise SystemExit(b.lib.plbm_last_error().decode())
        if b.lib.plbm_fetch_wait(b.sim._h) != 0:
            raise SystemExit(b.lib.plbm_last_error().decode())
        b.sync(); dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": cells * Ke / float(dt.item()) / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": NF * cells * 8, "steps": Ke,
                       "what": "initial state built on the device (plbm_initialize, as LBmethod's constructor does); per step the time step + "
                               "download of every slab's 15 visualised fields into pinned host memory (all ranks in parallel; the copy of step t "
                               "overlaps the kernels of step t+1)",
                       "host_topology": host_topology()}
        del out_host
    drv.close()
    b.close()
    if not args.no_extra:
        # the configurations BASELINE.json states its scaling targets on, on the same ranks in the same run
        line["strong_8192"] = extra_config(args, P, torch, dist, world, rank, local_rank, 8192,
                                           "8192x8192 strong scaling (BASELINE.json configs[3])", peak)
        wk = (WEAK4096_SIDES_LITERAL if args.weak_sides == "literal" else WEAK4096_SIDES).get(world)
        if wk:
            per = wk * wk / world / 4096 ** 2
            line["weak_4096"] = extra_config(args, P, torch, dist, world, rank, local_rank, wk,
                                             f"{wk}x{wk} = {per:.3f} x 4096^2 cells per GPU, weak scaling (BASELINE.json configs[4])", peak)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if parity is not None and parity["mismatches"]:
        raise SystemExit("bench.py: the slab-decomposed path is NOT bit-identical to the host checker (see the line's parity object)")


def host_topology():
    """Where the host side of the copies runs: NUMA nodes and CPU count (the per-GPU D2H rate at N > 1 depends on it)."""
    info = {"cpus": os.cpu_count()}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["numa_nodes"] = len(nodes)
    except OSError:
        info["numa_nodes"] = None
    return info


if __name__ == "__main__":
    main()
