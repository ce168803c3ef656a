"""C5 probe: full-catalogue leave-one-out evaluate (top-100) on the c4 shapes after a few training epochs.
Prints one JSON line: evaluate seconds (device + host), engine, candidate users, re-scored pairs, and a
sample cross-check of the tcgen05 engine against the exact fp64 single-user scan."""
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
    from eals_cpp_b200.model import MF_fastALS
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    topK = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    spec, sm, test_items = bench.build_workload(name, 0)
    M, N, K = spec["M"], spec["N"], spec["K"]
    fals = MF_fastALS(sm, None, topK=topK, factors=K, showLoss=False, init=False, device=0)
    U0, V0 = bench.random_factors(M, N, K, 0)
    fals.setUV(U0, V0)
    del U0, V0
    for _ in range(epochs):
        fals.update_user(); fals.update_item()
    fals.sync()
    gt = test_items.cpu().numpy().astype(np.int32) if hasattr(test_items, "cpu") else np.asarray(test_items, np.int32)
    out = {"workload": name, "epochs": epochs, "topK": topK, "loss": fals.loss()}
    for rep in range(2):
        t0 = time.perf_counter()
        res, hr, ndcg, prec, cnt = fals.evaluate(gt, topK, per_user=True)
        out[f"evaluate_s_rep{rep}"] = time.perf_counter() - t0
    out["device_ms"] = fals.timings()["evaluate"]
    out["metrics"] = res.tolist()
    out.update(fals.eval_stats())
    out["survivors"] = int((cnt <= topK).sum())
    if os.environ.get("EVAL_PROBE_NO_SAMPLE") == "1":
        print(json.dumps(out), flush=True)
        return
    rng = np.random.default_rng(0)
    surv = np.flatnonzero(cnt <= topK)
    sample = np.concatenate([rng.choice(M, 30, replace=False), rng.choice(surv, min(30, len(surv)), replace=False)]) if len(surv) else rng.choice(M, 30, replace=False)
    bad = 0
    for u in sample:
        one = fals.evaluate_for_user(int(u), int(gt[u]), topK)     # exact fp64 engine (single user)
        bad += one != [hr[u], ndcg[u], prec[u]]
    out["sample_checked"], out["sample_mismatches"] = len(sample), bad
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
