"""bench.py's reference arm (the only arm that runs without a GPU): one JSON line with the contract's keys, and the
non-zero ranks of a torchrun launch exit without work."""
import json
import os
import subprocess
import sys


def _run(repo_root, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(repo_root, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--cpu-n", "8"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line(repo_root):
    r = _run(repo_root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "fas_mg_mhex_steps_per_sec" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "Mhex-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the time of the steps that were really timed (the cpu-n sample), value their throughput
    assert d["extrapolated"] is True and cb["sample_n"] == 8
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - 8 ** 3 / 1e6) < 1e-9


def test_reference_arm_sets_its_own_thread_count(repo_root):
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must use the host's cores anyway and say how many."""
    r = _run(repo_root, {"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly(repo_root):
    r = _run(repo_root, {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_of_the_stated_workload(repo_root):
    """The roofline numerators bench.py reports for the cube of 1 M hexahedra (SURVEY.md 8d formulas over the real V, Z, N_h of
    every level): the values the committed bench lines carry (profiles/r02d_bench_fas_n1.json), so the accounting cannot drift
    unnoticed."""
    import importlib.util
    import types
    from fem_simulation_b200 import hierarchy, mesh
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(repo_root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    H = hierarchy.build_hierarchy(*mesh.cube_mesh(100), 1.0 / 100)
    assert H.dims == [1030301, 132651, 17576, 17576, 17576] and H.pattern[0].nnz == 27270901
    obj = types.SimpleNamespace(hierarchy=H)
    assert 352 * 1000000 + 36 * H.pattern[0].nnz + 256 * 1000000 == 1589752436          # the assembly stage (roofline.achieved)
    assert bench.step_bytes(obj, "fasv3") == 4375273208.0
    assert bench.step_bytes(obj, "nmg") == 20650060952.0
    assert bench.gs_launch_bytes(obj) == 129713842.5
