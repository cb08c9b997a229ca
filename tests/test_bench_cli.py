"""bench.py's CPU arm (`--impl reference`) runs without a GPU: one JSON line on stdout with the keys the
driver reads, same metric / unit / config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-images", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "masks/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("topo-loss fwd+bwd masks/sec")
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--ref-images", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_nvml_clock_sampler_with_a_fake_nvml(monkeypatch):
    """The in-process NVML sampler: samples between mark() and mark_end() only, median clock, reason bits -> names;
    without NVML (this container) make_sampler falls back to the nvidia-smi loop."""
    import sys
    import time
    import types
    import bench
    assert type(bench.make_sampler(0)).__name__ == "ClockSampler"  # no libnvidia-ml here
    fake = types.ModuleType("pynvml")
    state = {"why": 0}
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByUUID = lambda u: (_ for _ in ()).throw(RuntimeError("no such uuid"))
    fake.nvmlDeviceGetHandleByIndex = lambda i: ("handle", i)
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: 1950
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: state["why"]
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    s = bench.make_sampler(0, "GPU-123")
    assert type(s).__name__ == "NvmlSampler" and s.h == ("handle", 0)
    s.start()
    time.sleep(0.02)
    state["why"] = 0x4 | 0x1  # sw_power_cap + gpu_idle (ignored)
    s.mark()
    time.sleep(0.03)
    s.mark_end()
    state["why"] = 0x8        # after the window: must not be reported
    time.sleep(0.01)
    out = s.stop()
    assert out["sm_mhz"] == 1950 and out["sm_max_mhz"] == 1965 and out["samples"] >= 5
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"
