"""bench.py's measurement contract, the parts that need no GPU: the reference arm prints ONE JSON line with the
same metric / unit / workload string our own arm uses (bench.workload_string is shared), and our own arm refuses to
run without a device instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line_on_the_small_workload():
    """`--impl reference` on c1 (the reference's own CPU-runnable size): the reference's code from oracle/_ref
    (the oracle port when the reference could not be compiled), nothing of the product loaded."""
    res = _bench("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["metric"] == "eals_nnzK_updates_per_s" and d["unit"] == "nnz*K updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    import re
    import bench
    m = re.match(r"c1: (\d+) users x (\d+) items, (\d+) interactions, K=(\d+),", d["config"]["workload"])
    from eals_cpp_b200 import datasets
    spec = datasets.WORKLOADS["c1"]
    assert m and (int(m.group(1)), int(m.group(2)), int(m.group(4))) == (spec["M"], spec["N"], spec["K"])
    assert d["config"]["workload"] == bench.workload_string("c1", spec, int(m.group(3)))   # the string our own arm prints
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "libeals_b200" not in res.stderr                                 # the product library is not on this path


def test_own_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                                              # on a GPU box the gpu-marked tests cover it
    res = _bench("--workload", "c1", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-configs")
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr
    assert not [l for l in res.stdout.split("\n") if l.startswith("{")]
