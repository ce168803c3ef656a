"""Epoch time of the C-ABI multi-GPU path (eals_group: ONE process, one host thread, N GPUs; the exchange runs
inside the library) on a bench workload.  `python tools/group_bench.py c4 8 [epochs]`; devices 0..N-1, or
`--virtual` to put all ranks on GPU 0.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from eals_cpp_b200.model import GroupMF_fastALS
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    virtual = "--virtual" in sys.argv
    spec, sm, test_items = bench.build_workload(name, 0)
    t0 = time.perf_counter()
    fals = GroupMF_fastALS(sm, None, topK=spec["topK"], factors=spec["K"], showLoss=False,
                           devices=[0] * n if virtual else list(range(n)))
    create_s = time.perf_counter() - t0
    for _ in range(3):
        fals.update_user(); fals.update_item()
    fals.sync()
    times = []
    for _ in range(epochs):
        t0 = time.perf_counter()
        fals.update_user(); fals.update_item()
        fals.sync()
        times.append(time.perf_counter() - t0)
    ms = float(np.median(times)) * 1e3
    out = {"path": "eals_group (single process, C ABI)", "workload": bench.workload_string(name, spec, sm.nnz), "ranks": n,
           "virtual": virtual, "epoch_ms_median": ms, "epoch_ms_min": float(np.min(times)) * 1e3,
           "value": 2.0 * sm.nnz * spec["K"] / (ms * 1e-3), "unit": bench.UNIT, "create_plus_init_s": create_s,
           "loss_after": fals.loss(), "replicas_consistent": fals.replicas_consistent(),
           "user_bounds": fals.user_bounds, "item_bounds": fals.item_bounds}
    if test_items is not None:
        gt = test_items.cpu().numpy().astype(np.int32)
        t0 = time.perf_counter()
        res = fals.evaluate(gt, spec["topK"])
        out["evaluate_first_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = fals.evaluate(gt, spec["topK"])
        out["evaluate_s"] = time.perf_counter() - t0
        out["hr_ndcg_rr"] = [float(x) for x in res]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
