"""The JSON line of bench.py (the driver's contract).  The reference arm is CPU-only (it times the reference's own programs in
oracle/_ref, or the oracle port), so its line can be checked here; the line of our arm is checked on the committed run of the
round (profiles/), which the GPU box produced with the same script."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float, "higher_is_better": bool,
          "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict}


def _check_common(d):
    for k, t in COMMON.items():
        assert k in d, k
        assert isinstance(d[k], t), (k, type(d[k]))
    assert d["metric"] == "ppm_1080p_1Mphotons_rays_per_second" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k


def test_reference_arm_line(oracle):
    """`bench.py --impl reference` on a tiny bounded sample: K steps and W warm-up steps as asked, all host threads, no GPU."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample", "96x54:32", "--detail", "0.05"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "96x54" in cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_committed_line_of_our_arm():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1*_bench_default.json")))
    assert files, "no committed default bench line"
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    _check_common(d)
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["dtype"] == "f32"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in cb, k
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["e2e"]["d2h_bytes_per_step"] == 1920 * 1080 * 12 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.05
