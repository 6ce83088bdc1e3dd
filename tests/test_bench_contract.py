"""bench.py contract on CPU: the reference arm (`--impl reference`, the oracle port timed on the
host cores) prints exactly one JSON line with the keys the driver reads; the b200 arm refuses to
run without a GPU instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                          capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("DOF-updates/sec")
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the bounded sample the CPU legs time is part of the config (same in both arms)
    assert d["config"]["cpu_sample"]["elements"] == [32, 32] and d["config"]["cpu_sample"]["ndofs"] == 16384
    assert "32x32 elements" in d["cpu_baseline"]["sample"]


def test_reference_arm_uses_every_host_thread_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the arm must not silently become a
    single-threaded baseline, and `cores` must be what OpenMP really uses."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    # ranks other than 0 exit without work and without output
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = _run("--workload", "cfg1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-e2e")
    assert out.returncode != 0
    assert "no usable CUDA device" in (out.stderr + out.stdout) or "CUDA" in (out.stderr + out.stdout)


def test_clock_sampler_parses_clocks_reasons_and_power(tmp_path):
    """The `clocks` record of the bench line: median SM clock, throttle reasons and (when nvidia-smi
    reports them) board power against its enforced limit -- what `sw_power_cap` refers to."""
    sys.path.insert(0, ROOT)
    import bench

    class _Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    p = tmp_path / "smi.csv"
    p.write_text(
        "0, 1710, 1965, 981.20, 0x0000000000000004, Not Active, Not Active, Not Active, Active, 1000.00\n"
        "0, 1725, 1965, 990.10, 0x0000000000000004, Not Active, Not Active, Not Active, Active, 1000.00\n"
        "0, 1740, 1965, [N/A], 0x0000000000000000, Not Active, Not Active, Not Active, Not Active, [N/A]\n"
        "garbage line\n")
    s = bench.ClockSampler(0)
    s.proc, s.path = _Done(), str(p)
    out = s.stop()
    assert out["sm_mhz"] == 1725.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]
    assert abs(out["power_w"] - 985.65) < 1e-9 and out["power_limit_w"] == 1000.0
