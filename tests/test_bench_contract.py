"""bench.py's reference arm (runs on host cores: no GPU needed) prints the contract's JSON line, and its `config`
names the workload with exactly the keys the GPU arm's line uses (the driver compares the two arms' configs)."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line_on_the_tiny_workload():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "gcn_fwd_bwd_epoch_ms" and d["unit"] == "ms"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] == d["value"] and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "tiny" in cb["sample"]
    assert d["config"]["workload"].startswith("tiny-shaped synthetic, 3-layer highway GCN")
    # the GPU arm builds its `config` from the same keys (read from the source: that arm needs a GPU to run)
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    gpu_keys = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "run_gpu":
            for sub in ast.walk(node):
                if isinstance(sub, ast.Dict):
                    for k, v in zip(sub.keys, sub.values):
                        if isinstance(k, ast.Constant) and k.value == "config" and isinstance(v, ast.Call):
                            gpu_keys = {kw.arg for kw in v.keywords}
    assert gpu_keys is not None
    base = {"workload", "n_layers", "highway", "l2"}                  # bench_config(args), shared by both arms
    assert set(d["config"]) == gpu_keys | base, (sorted(d["config"]), sorted(gpu_keys | base))


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
