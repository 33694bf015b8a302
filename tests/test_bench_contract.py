"""bench.py's CPU-runnable arm (`--impl reference`) and the JSON contract of its one output line — no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "2", "--points", "1024", *args], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "chamfer_fwd_bwd_point_pairs_per_s" and d["unit"] == "point-pairs/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert out.strip() == ""


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still use the cores it is allowed."""
    d = json.loads(_run({"OMP_NUM_THREADS": "1"}).strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_committed_gpu_bench_lines_carry_the_contract():
    """The B200 lines committed under profiles/ (what the round's numbers are quoted from): every key of the bench
    contract, the roofline and cpu_baseline objects, clocks sampled during the timed region, and weak scaling of
    `value` over the committed GPU counts."""
    base = None
    for n in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", f"r1_bench_n{n}.json")
        if not os.path.isfile(path):
            continue
        d = json.load(open(path))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert key in d, (n, key)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["dtype"] == "f32" and d["gpu_launches"] > 0
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            base = d["value"]
            c = d["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        elif base:
            assert d["value"] / (n * base) > 0.9, (n, d["value"], base)
