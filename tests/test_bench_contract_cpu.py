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
