"""Host-side logic that needs no GPU: partitioning, the dual-orientation container, the synthetic
generators, and the multi-rank exchange (world_size 2 over gloo on CPU tensors)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import random_csr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_by_nnz_balances_and_covers():
    from eals_cpp_b200.model import partition_by_nnz
    rng = np.random.default_rng(0)
    lens = np.minimum(rng.zipf(1.5, 5000), 4000)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for world in (1, 2, 3, 8):
        b = partition_by_nnz(ptr, world)
        assert b[0] == 0 and b[-1] == 5000 and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world))
        per = [ptr[b[i + 1]] - ptr[b[i]] for i in range(world)]
        assert max(per) <= ptr[-1] / world + lens.max()


def test_sparsemat_from_csr_is_consistent():
    from eals_cpp_b200.model import SparseMat
    row_ptr, col_idx = random_csr(80, 50, 6, seed=1)
    val = np.arange(len(col_idx), dtype=np.float64)
    sm = SparseMat.from_csr(80, 50, row_ptr, col_idx, val)
    assert sm.nnz == len(col_idx) and sm.col_ptr[-1] == sm.nnz
    dense = np.zeros((80, 50))
    for u in range(80):
        dense[u, col_idx[row_ptr[u]:row_ptr[u + 1]]] = val[row_ptr[u]:row_ptr[u + 1]] + 1
    for i in range(50):
        rows = sm.row_idx[sm.col_ptr[i]:sm.col_ptr[i + 1]]
        assert np.all(np.diff(rows) > 0)
        assert np.array_equal(dense[rows, i], sm.col_val[sm.col_ptr[i]:sm.col_ptr[i + 1]] + 1)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_generator_contract(name):
    from eals_cpp_b200 import datasets
    d = datasets.make(name)
    spec = datasets.WORKLOADS[name]
    assert d.M == spec["M"] and d.N == spec["N"]
    assert abs(d.nnz - spec["nnz"]) < 0.25 * spec["nnz"]
    assert d.col_idx.dtype == np.int32 and d.row_ptr.dtype == np.int64
    for u in range(0, d.M, 37):
        r = d.col_idx[d.row_ptr[u]:d.row_ptr[u + 1]]
        assert np.all(np.diff(r) > 0)
    assert d.test_items.min() >= 0 and d.test_items.max() < d.N
    d2 = datasets.make(name)
    assert np.array_equal(d.col_idx, d2.col_idx)            # seeded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from eals_cpp_b200.model import allreduce_sum, exchange_rows, partition_by_nnz
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        lens = rng.integers(0, 30, 101)
        ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        bounds = partition_by_nnz(ptr, world)
        truth = torch.arange(101 * 8, dtype=torch.float64).reshape(101, 8)
        full = torch.full((101, 8), -1.0, dtype=torch.float64)
        full[bounds[rank]:bounds[rank + 1]] = truth[bounds[rank]:bounds[rank + 1]]   # "my updated rows"
        exchange_rows(full, bounds, rank)
        ok_rows = bool(torch.equal(full, truth))
        # partial Grams over the owned rows sum to the full Gram
        mine = truth[bounds[rank]:bounds[rank + 1]]
        S = mine.T @ mine
        allreduce_sum(S)
        ok_gram = bool(torch.allclose(S, truth.T @ truth, rtol=1e-14))
        # setTrain on several ranks: each rank copies 1/world of an index array, the all-gather completes it
        from eals_cpp_b200.model import gather_full_array
        ok_gather = True
        for n in (1, 7, 64, 1001):
            host = (np.arange(n, dtype=np.int32) * 3 + 1)
            full, buf = gather_full_array(host, n, rank, world, "cpu")
            ok_gather &= bool(np.array_equal(full.numpy()[:n], host))
            full2, buf2 = gather_full_array(host[::-1].copy(), n, rank, world, "cpu", buf)   # buffer reuse
            ok_gather &= buf2 is buf and bool(np.array_equal(full2.numpy()[:n], host[::-1]))
        q.put((rank, ok_rows, ok_gram and ok_gather, bounds))
    finally:
        dist.destroy_process_group()


def test_exchange_and_gram_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == res[1][3]


def test_insert_interaction_keeps_both_orientations_sorted():
    """Host part of the online updateModel (MF_fastALS.cpp:223-224): the new entry lands at its sorted
    position in the CSR and in the CSC arrays; an existing entry changes nothing."""
    from eals_cpp_b200.model import SparseMat, insert_interaction
    from conftest import random_csr
    M, N = 40, 30
    row_ptr, col_idx = random_csr(M, N, 5, seed=3, empty_frac=0.1)
    sm = SparseMat.from_csr(M, N, row_ptr, col_idx)
    dense = np.zeros((M, N), bool)
    for u in range(M):
        dense[u, col_idx[row_ptr[u]:row_ptr[u + 1]]] = True
    rng = np.random.default_rng(0)
    for _ in range(50):
        u, i = int(rng.integers(M)), int(rng.integers(N))
        grown = insert_interaction(sm, u, i)
        if dense[u, i]:
            assert grown is None
            continue
        dense[u, i] = True
        sm = grown
        want = SparseMat.from_csr(M, N, *_csr_of(dense))
        for a, b in ((sm.row_ptr, want.row_ptr), (sm.col_idx, want.col_idx), (sm.col_ptr, want.col_ptr), (sm.row_idx, want.row_idx)):
            assert np.array_equal(a, b)
    vals = SparseMat.from_csr(M, N, sm.row_ptr, sm.col_idx, np.full(sm.nnz, 2.0))
    u, i = np.argwhere(~dense)[0]
    g = insert_interaction(vals, int(u), int(i))
    assert g.row_val.sum() == 2.0 * sm.nnz + 1.0 and g.col_val.sum() == g.row_val.sum()   # rating 1 = w_new


def _csr_of(dense):
    row_ptr = np.concatenate([[0], np.cumsum(dense.sum(1))]).astype(np.int64)
    col_idx = np.concatenate([np.flatnonzero(r) for r in dense]).astype(np.int32)
    return row_ptr, col_idx


def test_update_model_host_flow_with_a_mocked_library():
    """updateModel's host side (MF_fastALS.cpp:223-242) without a GPU: the matrix handed to setTrain has the
    new entry, a brand-new item gets w0 / itemCount, and the 10 alternating row updates are each followed by
    their S patch."""
    from eals_cpp_b200 import model as mdl
    from conftest import random_csr
    M, N = 30, 20
    row_ptr, col_idx = random_csr(M, N, 4, seed=1)
    keep = [[c for c in col_idx[row_ptr[u]:row_ptr[u + 1]] if c != 3] for u in range(M)]      # item 3 unseen
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in keep])]).astype(np.int64)
    col_idx = np.array([c for r in keep for c in r], np.int32)
    calls = []

    class Fake(mdl.MF_fastALS):
        def __init__(self):                          # no library, no device
            self.world, self.rank, self.userCount, self.itemCount, self.w0 = 1, 0, M, N, 10.0
            self.user_bounds, self.item_bounds = [0, M], [0, N]
            self.trainMatrix = mdl.SparseMat.from_csr(M, N, row_ptr, col_idx)
            self._wi = np.ones(N); self._wi[3] = 0.0
        Wi = property(lambda self: self._wi.copy(), lambda self, w: (calls.append(("Wi", w.copy())), setattr(self, "_wi", w)))
        def setTrain(self, sm): calls.append(("setTrain", sm)); self.trainMatrix = sm
        def _factor_row(self, which, r): return np.full(4, float(len(calls)))
        def update_user_thread(self, u): calls.append(("user", u))
        def update_item_thread(self, i): calls.append(("item", i))
        def update_user_SU(self, old, new): calls.append(("SU", old[0] < new[0]))
        def update_item_SV(self, i, old, new): calls.append(("SV", i))
        def close(self): pass

    f = Fake()
    f.updateModel(5, 3)
    kinds = [c[0] for c in calls]
    assert kinds[0] == "setTrain" and kinds[1] == "Wi"
    sm = calls[0][1]
    assert 3 in sm.col_idx[sm.row_ptr[5]:sm.row_ptr[6]] and sm.nnz == len(col_idx) + 1
    assert 5 in sm.row_idx[sm.col_ptr[3]:sm.col_ptr[4]]
    assert calls[1][1][3] == 10.0 / N
    assert kinds[2:] == ["user", "SU", "item", "SV"] * 10
    calls.clear()
    f.updateModel(5, 3, patch_S=False)                # already present: no setTrain, weight already set, stale caches
    assert [c[0] for c in calls] == ["user", "item"] * 10
    with pytest.raises(IndexError):
        f.updateModel(M, 0)


def test_update_model_on_two_ranks_reads_the_old_row_before_the_owner_stores():
    """Sharded updateModel: the owner's single-row kernel stores the new row into EVERY replica.  A rank that is
    slow to read the "old" row of its S patch must still see the row as it was before that store — two ranks as
    threads over shared replicas, rank 1 late at every read; every patch on both ranks must be (k, k + 1)."""
    import threading
    import time
    from eals_cpp_b200 import model as mdl
    from eals_cpp_b200 import _lib
    M, N = 8, 6
    dense = np.ones((M, N), bool)                              # (u, i) already present: no setTrain
    row_ptr, col_idx = _csr_of(dense)
    rep = [{_lib.BUF_U: np.zeros(2), _lib.BUF_V: np.zeros(2)} for _ in range(2)]     # row u of U / row i of V per replica
    barrier = threading.Barrier(2)

    class FakeRank(mdl.MF_fastALS):
        def __init__(self, rank):
            self.world, self.rank, self.userCount, self.itemCount, self.w0 = 2, rank, M, N, 10.0
            self.user_bounds, self.item_bounds, self.peer_store = [0, M // 2, M], [0, N // 2, N], True
            self.trainMatrix = mdl.SparseMat.from_csr(M, N, row_ptr, col_idx)
            self.patches = []
        Wi = property(lambda self: np.ones(N))
        def _factor_row(self, which, r):
            if self.rank == 1:
                time.sleep(0.01)                               # the late reader
            return rep[self.rank][which].copy()
        def _store_everywhere(self, which):
            new = rep[self.rank][which] + 1.0
            for r in rep:
                r[which] = new.copy()
        def update_user_thread(self, u): self._store_everywhere(_lib.BUF_U)
        def update_item_thread(self, i): self._store_everywhere(_lib.BUF_V)
        def update_user_SU(self, old, new): self.patches.append(("SU", old[0], new[0]))
        def update_item_SV(self, i, old, new): self.patches.append(("SV", old[0], new[0]))
        def _rank_barrier(self): barrier.wait(timeout=20)
        def close(self): pass

    ranks = [FakeRank(0), FakeRank(1)]                         # rank 0 owns user 1 and item 1
    threads = [threading.Thread(target=f.updateModel, args=(1, 1)) for f in ranks]
    for t in threads: t.start()
    for t in threads: t.join(30)
    want = [(kind, float(k), float(k + 1)) for k in range(10) for kind in ("SU", "SV")]
    assert ranks[0].patches == want
    assert ranks[1].patches == want


def test_cost_partition_is_contiguous_complete_and_balances_cost():
    """eals_partition (used by eals_group and by the one-process-per-GPU path): bounds are monotone, cover every
    row, give every rank at least one row, and balance the per-row cost model rather than raw nonzeros."""
    from eals_cpp_b200.model import partition_by_cost, partition_by_nnz
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(1, 30, 5000), rng.integers(600, 5000, 40), rng.integers(1, 30, 5000)])
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for world in (1, 2, 3, 8):
        b = partition_by_cost(ptr, world)
        assert b[0] == 0 and b[-1] == len(lens) and all(b[i] < b[i + 1] for i in range(world))
        cost = np.where(lens <= 32, 13.0, np.where(lens <= 128, 0.47 * lens, np.where(lens <= 512, 0.42 * lens, 0.37 * lens)))
        per = [cost[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(per) <= 1.15 * (sum(per) / world) + cost.max()
    # short rows weigh more per nonzero than heavy ones: the cost split differs from the nnz split
    lens2 = np.concatenate([rng.integers(1, 30, 8000), rng.integers(600, 5000, 60)])
    ptr2 = np.concatenate([[0], np.cumsum(lens2)]).astype(np.int64)
    assert partition_by_cost(ptr2, 2) != partition_by_nnz(ptr2, 2)
    assert partition_by_cost(np.array([0, 3, 6], np.int64), 2) == [0, 1, 2]
