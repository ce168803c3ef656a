"""Worker for the multi-GPU parity test: run under torch.distributed.run with N ranks (one per GPU).
Every rank builds the same small matrix, trains 3 epochs with users/items sharded over the ranks and
checks its replica of U, V, the loss and the evaluation against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from conftest import random_csr
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    from oracle.bindings import PortModel
    M, N, K = 1500, 700, 32
    row_ptr, col_idx = random_csr(M, N, 20, seed=42, empty_frac=0.03)
    # a few heavy columns so that every kernel family runs on some rank
    rng = np.random.default_rng(1)
    rows = [set(col_idx[row_ptr[u]:row_ptr[u + 1]].tolist()) for u in range(M)]
    for c in (3, 250, 600):
        for u in rng.choice(M, size=1200, replace=False):
            rows[u].add(c)
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col_idx = np.concatenate([np.array(sorted(r), np.int32) for r in rows])
    gt = np.random.default_rng(2).integers(0, N, M).astype(np.int32)
    sm = SparseMat.from_csr(M, N, row_ptr, col_idx)
    fals = MF_fastALS(sm, gt, factors=K, showLoss=False, device=local)
    assert fals.world == dist.get_world_size() and fals.world > 1
    port = PortModel(M, N, row_ptr, col_idx, factors=K)
    # Start-up race (round-1 advisor finding): the last rank is LATE with its whole-replica overwrite while
    # the others are already past theirs.  Without the barrier that ends setUV the early ranks' first sweep
    # peer-stores finished rows into the late rank's replica and its upload then lands on top of them.
    import time
    U0, V0 = port.U * 1.25, port.V * 0.75
    port.U[:], port.V[:] = U0, V0
    port.init_S()
    if rank == fals.world - 1:
        time.sleep(1.0)
    fals.setUV(U0, V0)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    assert fals.replicas_consistent(), rank
    assert np.abs(fals.U - port.U).max() < 1e-10 and np.abs(fals.V - port.V).max() < 1e-10, rank
    fals.barrier()        # raw replica reads are local: no rank may start the next sweep while another still reads
    for it in range(3):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
        lg, lc = fals.loss(), port.loss()
        assert abs(lg - lc) <= 1e-10 * abs(lc), (rank, it, lg, lc)
    # setTrain on several ranks (chunked upload + all-gather of the index arrays, caches rebuilt and
    # re-shared), then one more epoch
    fals.setTrain(sm)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc), (rank, "after setTrain", lg, lc)
    # a LARGER matrix: the prediction caches have to grow, so the ranks exchange IPC handles again
    from oracle.bindings import csr_to_csc
    for u in rng.choice(M, size=900, replace=False):
        rows[u].update(int(c) for c in rng.choice(N, size=12, replace=False))
    row_ptr2 = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col_idx2 = np.concatenate([np.array(sorted(r), np.int32) for r in rows])
    fals.setTrain(SparseMat.from_csr(M, N, row_ptr2, col_idx2))
    port.row_ptr, port.col_idx = row_ptr2, col_idx2
    port.col_ptr, port.row_idx, port.cval, _ = csr_to_csc(M, N, row_ptr2, col_idx2, None)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc), (rank, "after growing setTrain", lg, lc)
    assert np.abs(fals.U - port.U).max() < 1e-10, rank        # every replica is complete
    assert np.abs(fals.V - port.V).max() < 1e-10, rank
    assert np.abs(fals.SU - port.SU).max() <= 1e-11 * np.abs(port.SU).max()
    assert fals.replicas_consistent(), rank
    # online update on the sharded model (one process per GPU): the owner runs the row kernel, the new row reaches
    # every replica through the kernel's peer stores, every rank patches its S caches
    if fals.peer_store:
        rows2 = [list(col_idx2[row_ptr2[r]:row_ptr2[r + 1]]) for r in range(M)]
        uu, ii = 5, next(c for c in range(N) if c not in rows2[5])
        port.SU = port.p.gram_plain(port.U); port.SV = port.p.gram_weighted(port.V, port.Wi)
        fals.refresh_S()
        fals.updateModel(uu, ii)
        rows2[uu] = sorted(rows2[uu] + [ii])
        rp3 = np.concatenate([[0], np.cumsum([len(r) for r in rows2])]).astype(np.int64)
        ci3 = np.concatenate([np.array(r, np.int32) for r in rows2])
        port.row_ptr, port.col_idx = rp3, ci3
        port.col_ptr, port.row_idx, port.cval, _ = csr_to_csc(M, N, rp3, ci3, None)
        for _ in range(10):
            port.update_user(uu, uu + 1)
            port.update_item(ii, ii + 1)
        assert fals.replicas_consistent(), rank
        assert np.abs(fals.U - port.U).max() < 1e-10 and np.abs(fals.V - port.V).max() < 1e-10, rank
        assert np.abs(fals.SU - port.SU).max() <= 1e-10 * np.abs(port.SU).max(), rank
        assert np.abs(fals.SV - port.SV).max() <= 1e-10 * np.abs(port.SV).max(), rank
    res = fals.evaluate()
    want = port.evaluate(gt, 10, compat=True)[0]
    assert np.allclose(res, want, rtol=0, atol=1e-12), (res, want)
    dist.barrier()
    if rank == 0:
        print(f"dist parity ok: world {fals.world}, peer_store {fals.peer_store}, peer_pred_cache {fals.peer_pred_cache}, loss {lg!r}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
