"""bench.py contract, CPU side: the reference arm runs without a GPU, prints ONE JSON line with the required keys,
and non-zero ranks of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"}


def _run(env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--grid", "32", "--cpu-rays", "16"], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=REPO)


def test_reference_arm_json_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert p.returncode == 0 and not any(l.startswith("{") for l in p.stdout.splitlines())


def test_bench_cli_defaults_and_workloads():
    """Default invocation = BASELINE configs[1] (the metric's config), N = 1, a K / W that finish within minutes; the
    other BASELINE configs are selectable and named in `config.workload`."""
    import importlib
    import sys
    sys.argv = ["bench.py"]
    bench = importlib.import_module("bench")
    a = bench.parse()
    assert (a.gpus, a.config, a.scaling, a.impl) == (1, 2, "weak", "ours")
    assert a.warmup >= 3 and 5 <= a.steps <= 50
    assert (a.envmap_h, a.envmap_w) == (16, 32)
    sys.argv = ["bench.py", "--config", "3"]
    a3 = bench.parse()
    assert (a3.envmap_h, a3.envmap_w) == (8, 16)            # the north-star's 128 secondary directions
    cfg = bench.workload_config(a3, "dp1")
    assert "multi_light_rotated" in cfg["workload"] and "8x16 = 128" in cfg["workload"]
    assert set(bench.WORKLOADS) == {2, 3, 4, 5}
    env = bench.synthetic_hdr(64, 128)
    assert env.shape == (64, 128, 3) and float(env.max()) == 400.0 and float(env.min()) > 0


def test_configargparse_stand_in_reads_reference_configs(tmp_path):
    """tools/ref_stubs/configargparse.py (lets the unmodified reference scripts start in this image): `key = value`
    files with comments, `[a, b, c]` lists for action="append" options, store_true flags, command line overrides."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_cfgparse", os.path.join(root, "tools", "ref_stubs", "configargparse.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p = mod.ArgumentParser()
    p.add_argument("--config", is_config_file=True)
    p.add_argument("--n_iters", type=int, default=1)
    p.add_argument("--upsamp_list", type=int, action="append")
    p.add_argument("--light_rotation", type=str, action="append")
    p.add_argument("--white_bkgd", action="store_true")
    p.add_argument("--expname", type=str)
    cfg = tmp_path / "c.txt"
    cfg.write_text("expname = run   # comment\nn_iters = 80000\nupsamp_list = [10000, 20000]\nlight_rotation = [000]\n"
                   "white_bkgd = 1\nunknown_key = 3\n")
    a = p.parse_args(["--config", str(cfg), "--n_iters", "24"])
    assert (a.expname, a.n_iters, a.upsamp_list, a.light_rotation, a.white_bkgd) == ("run", 24, [10000, 20000], ["000"], True)


def test_clock_sampler_reports_rows_of_the_timed_region(tmp_path, monkeypatch):
    """ClockSampler against a stand-in nvidia-smi with a 200 ms start-up: started before the warm-up, only rows that arrive
    after mark() are reported ("window": "timed"); a throttle reason is picked up; without nvidia-smi it says so."""
    import importlib
    import stat
    import time
    stub = tmp_path / "nvidia-smi"
    stub.write_text("#!" + sys.executable + "\nimport sys, time\nper = int(sys.argv[sys.argv.index('-lms') + 1]) / 1000\n"
                    "time.sleep(0.2)\nk = 0\nwhile True:\n"
                    "    clk = 1200 if k < 2 else 1965\n    k += 1\n"
                    "    print(f'0, {clk}, 1965, 512.3, 0x4, Not Active, Not Active, Not Active, Active', flush=True)\n"
                    "    time.sleep(per)\n")
    stub.chmod(stub.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.argv = ["bench.py"]
    bench = importlib.import_module("bench")
    c = bench.ClockSampler(0)
    c.start()
    time.sleep(0.45)                      # "warm-up": the start-up and the two low-clock rows fall in here
    c.mark()
    time.sleep(0.2)                       # "timed region"
    out = c.stop()
    assert out["window"] == "timed" and out["samples"] >= 2
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    c = bench.ClockSampler(0)
    c.mark()
    assert c.stop()["reasons"] == ["nvidia-smi unavailable"]
