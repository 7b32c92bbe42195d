"""tools/bench_csv.py: bench lines in the reference's timing-CSV format."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF_HEADER = "Grid_Dimension,Number_of_Steps,Number_of_Cores,Poisson,BC,Total_Computation_Time(ms)"   # src/main_plasma.cpp:88


def test_bench_csv_keeps_the_reference_columns(tmp_path):
    line = {"metric": "MLUPS (all species)", "value": 3656.5, "unit": "MLUPS", "n_gpus": 1, "steps": 192, "warmup": 5, "ms_per_step": 1.147,
            "config": {"workload": "2048x2048 plasma D2Q9 3 species + DDF thermal, FFT Poisson, periodic (BASELINE.json configs[2])"},
            "roofline": {"achieved": 3590.0, "frac": 0.548}, "e2e": {"value": 434.2}}
    ref = dict(line, impl="reference", value=4.4, ms_per_step=950.0, steps=12, roofline=None, e2e={"value": 4.4}, cpu_baseline={"cores": 16})
    src = tmp_path / "lines.jsonl"
    src.write_text(json.dumps(line) + "\n" + json.dumps(ref) + "\n" + json.dumps({"impl": "reference", "unavailable": "x"}) + "\n")
    out = tmp_path / "out.csv"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_csv.py"), str(src), "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = out.read_text().strip().splitlines()
    assert rows[0].startswith(REF_HEADER + ",")
    assert len(rows) == 3
    gpu, cpu = rows[1].split(","), rows[2].split(",")
    assert gpu[:5] == ["2048x2048", "192", "0", "3", "0"] and abs(float(gpu[5]) - 1.147 * 192) < 1e-6
    assert gpu[6] == "1" and float(gpu[7]) == 3656.5 and float(gpu[9]) == 0.548
    assert cpu[:5] == ["2048x2048", "12", "16", "3", "0"] and cpu[6] == "0"


def test_bench_clock_sampler_windows_its_samples():
    """bench.py's clock sampler keeps only the nvidia-smi lines that arrived inside the timed window (plus one period)."""
    sys.path.insert(0, str(ROOT))
    import bench
    s = bench.ClockSampler(0)
    line = "1965, 1965, 612.3, Not Active, Not Active, Not Active, {}"
    s.samples = [(9.90, line.format("Not Active")), (10.01, line.format("Active")), (10.05, line.format("Not Active")),
                 (10.30, "1200, 1965, 100.0, Active, Not Active, Not Active, Not Active")]
    s.t_begin, s.t_end = 10.0, 10.1
    out = s.summary()
    assert out["samples"] == 2 and out["sm_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    s.t_begin, s.t_end = 20.0, 20.001                      # nothing inside: the nearest line is used
    assert s.summary()["samples"] == 1
    s.samples = []
    assert s.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_bench_segments_cover_exactly_k_steps():
    sys.path.insert(0, str(ROOT))
    import bench
    for k in (1, 47, 48, 49, 192, 200):
        seg = list(bench.segments(k))
        assert sum(seg) == k and max(seg) <= bench.SEGMENT and min(seg) >= 1
