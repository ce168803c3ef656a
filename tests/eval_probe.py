"""Evaluate at scale: trains a couple of epochs on a workload and times evaluate() (both modes)."""
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, bench
from eals_cpp_b200.model import MF_fastALS
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec, sm, test_items = bench.build_workload(name, 0)
fals = MF_fastALS(sm, test_items.cpu().numpy(), topK=spec["topK"], factors=spec["K"], showLoss=False, init=False, device=0)
U, V = bench.random_factors(spec["M"], spec["N"], spec["K"], 0); fals.setUV(U, V); del U, V
for _ in range(epochs):
    fals.update_user(); fals.update_item()
fals.sync()
for exact in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = fals.evaluate(exact=exact)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    flops = 2.0 * spec["M"] * spec["N"] * spec["K"]
    print(f"{name} evaluate exact={exact}: {dt*1e3:.1f} ms  hr/ndcg/mrr {res.tolist()}  (dense-scan equivalent {flops/dt/1e12:.1f} TFLOP/s)", flush=True)
