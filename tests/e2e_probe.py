import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
os.environ["EALS_VERBOSE"] = "1"
import torch, bench
from eals_cpp_b200.model import MF_fastALS, SparseMat
spec, sm, _ = bench.build_workload("c4", 0)
fals = MF_fastALS(sm, None, factors=spec["K"], showLoss=False, init=False, device=0)
U, V = bench.random_factors(spec["M"], spec["N"], spec["K"], 0); fals.setUV(U, V); del U, V
def pinned(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h
keep = [pinned(x) for x in (sm.row_ptr, sm.col_idx, sm.col_ptr, sm.row_idx)]
smh = SparseMat(spec["M"], spec["N"], *[k.numpy() for k in keep])
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fals.setTrain(smh); torch.cuda.synchronize()
    print("setTrain total %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
