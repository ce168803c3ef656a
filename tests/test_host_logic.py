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
        q.put((rank, ok_rows, ok_gram, bounds))
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
