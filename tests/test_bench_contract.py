"""bench.py's JSON contract, exercised on the CPU through the reference arm (the B200 arm
needs a GPU and is run by the driver): exactly one JSON line on stdout with the keys the
driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "0", "--seqs", "40", "--ref-seconds", "0.3", "--ref-max-seqs", "24"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "dtw_gcups" and d["unit"] == "GCUPS"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _collective_calls_under_rank_branches(path):
    """Every rank must issue the same sequence of collectives: a call that contains one (NCCL
    all-gather inside align_all*, barrier, all_reduce, ...) inside an `if rank ...:` branch
    deadlocks the other ranks.  Static check of the source."""
    import ast
    collectives = {"align_all_device", "align_all", "barrier", "maxrank", "all_reduce", "all_gather",
                   "all_gather_into_tensor", "broadcast", "reduce_scatter", "measure_other_mode", "ShardedAligner",
                   "init_process_group", "destroy_process_group"}
    tree = ast.parse(open(path).read())
    bad = []

    def mentions_rank(node):
        return any(isinstance(n, ast.Name) and n.id in ("rank", "local") or
                   isinstance(n, ast.Attribute) and n.attr == "rank" for n in ast.walk(node))

    for node in ast.walk(tree):
        if isinstance(node, ast.If) and mentions_rank(node.test):
            for sub in node.body + node.orelse:
                for call in ast.walk(sub):
                    if isinstance(call, ast.Call):
                        f = call.func
                        name = f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", None)
                        on_oracle = isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "oracle"
                        if name in collectives and not on_oracle:  # oracle.align_all is the CPU checker, not a collective
                            bad.append((node.lineno, call.lineno, name))
    return bad


def test_no_collective_is_issued_by_a_subset_of_ranks():
    for rel in ("bench.py", os.path.join("audio_pattern_discovery_b200", "distributed.py"),
                os.path.join("tests", "multi_rank_check.py")):
        bad = _collective_calls_under_rank_branches(os.path.join(ROOT, rel))
        assert not bad, (rel, bad)
