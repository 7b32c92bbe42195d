#!/usr/bin/env python
"""bench_csv.py -- bench.py JSON lines as rows of the reference's timing CSV.

The reference appends one row per run to build/simulation_time_plasma_details.csv
(src/main_plasma.cpp:80-94): Grid_Dimension,Number_of_Steps,Number_of_Cores,Poisson,BC,Total_Computation_Time(ms);
build/Scalability_analysis.py and build/weak_scalability.py plot that file.  This tool writes the same six
columns (so those scripts read it unchanged) followed by GPU columns in the same row:
GPUs,MLUPS,K1_GBps,Roofline_Fraction,E2E_MLUPS.

    python tools/bench_csv.py profiles/r1_bench_default_n1.json profiles/r1_scaling_8gpu.jsonl -o build/simulation_time_plasma_details_b200.csv
"""
from __future__ import annotations

import argparse
import json
import re
import sys

HEADER = "Grid_Dimension,Number_of_Steps,Number_of_Cores,Poisson,BC,Total_Computation_Time(ms)"
EXTRA = "GPUs,MLUPS,K1_GBps,Roofline_Fraction,E2E_MLUPS"
POISSON = {"none": 0, "gs": 1, "sor": 2, "fft": 3, "nps": 4}       # enumerator order of include/poisson.hpp
BC = {"periodic": 0, "bounceback": 1}                               # include/streaming.hpp


def row(line: dict) -> str:
    wl = line.get("config", {}).get("workload", "")
    m = re.search(r"(\d+)x(\d+)", wl)
    if not m:
        raise ValueError(f"no lattice size in workload {wl!r}")
    poisson = next((v for k, v in POISSON.items() if re.search(rf"\b{k}\b", wl, re.I)), POISSON["fft"])
    bc = BC["bounceback"] if re.search(r"bounce", wl, re.I) else BC["periodic"]
    steps = int(line["steps"])
    total_ms = float(line["ms_per_step"]) * steps
    cores = line.get("cpu_baseline", {}).get("cores", 0) if line.get("impl") == "reference" else 0
    roof, e2e = line.get("roofline") or {}, line.get("e2e") or {}
    fmt = lambda v: "" if v is None else f"{v:.6g}"
    return (f"{m.group(1)}x{m.group(2)},{steps},{cores},{poisson},{bc},{total_ms:.3f},"
            f"{0 if line.get('impl') == 'reference' else line.get('n_gpus', 1)},{fmt(line.get('value'))},{fmt(roof.get('achieved'))},"
            f"{fmt(roof.get('frac'))},{fmt(e2e.get('value'))}")


def read_lines(paths):
    for p in paths:
        with open(p) as fh:
            for raw in fh:
                raw = raw.strip()
                if raw.startswith("{"):
                    yield json.loads(raw)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("inputs", nargs="+", help="files holding bench.py JSON lines (.json / .jsonl)")
    ap.add_argument("-o", "--output", default="-")
    args = ap.parse_args(argv)
    rows = [row(l) for l in read_lines(args.inputs) if "unavailable" not in l]
    text = HEADER + "," + EXTRA + "\n" + "\n".join(rows) + "\n"
    if args.output == "-":
        sys.stdout.write(text)
    else:
        with open(args.output, "w") as fh:
            fh.write(text)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
