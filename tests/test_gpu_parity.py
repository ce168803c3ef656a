"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (fp64 everywhere; only the summation ORDER over a row's nonzeros and over rows differs
from the reference's sequential loops):
  * factors after a half-epoch:   |dU| <= 1e-11 absolute (values are O(0.01..1))
  * per-epoch loss:               <= 1e-10 relative (north star allows 1e-6)
  * S caches:                     <= 1e-12 relative to the largest entry
  * evaluation: count_larger and the HR/NDCG/reciprocal-rank tuples identical per user — scores are
    computed with the reference's exact operation order (sequential k, no FMA), so they are
    bit-identical, and the ranking replay uses the same libstdc++ partial_sort_copy.
"""
import numpy as np
import pytest

from conftest import random_csr

pytestmark = pytest.mark.gpu


def _models(M, N, row_ptr, col_idx, K, val=None, **kw):
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    from oracle.bindings import PortModel
    sm = SparseMat.from_csr(M, N, row_ptr, col_idx, val)
    fals = MF_fastALS(sm, None, factors=K, showLoss=False, debug_sync=True, **kw)
    port = PortModel(M, N, row_ptr, col_idx, val, factors=K, **kw)
    return fals, port


def _heavy_matrix(M, N, heavy_cols, heavy_len, seed):
    """A few columns rated by `heavy_len` users each (heavy rows on the item side), sparse rest."""
    rng = np.random.default_rng(seed)
    rows = [set() for _ in range(M)]
    for c in heavy_cols:
        for u in rng.choice(M, size=heavy_len, replace=False):
            rows[u].add(int(c))
    for u in range(M):
        for c in rng.choice(N, size=int(rng.integers(1, 4)), replace=False):
            rows[u].add(int(c))
    row_ptr = np.zeros(M + 1, np.int64)
    cols = []
    for u in range(M):
        c = np.array(sorted(rows[u]), np.int32)
        cols.append(c)
        row_ptr[u + 1] = row_ptr[u] + len(c)
    return row_ptr, np.concatenate(cols)


@pytest.mark.parametrize("K", [8, 16, 64, 128, 20])
def test_init_and_S_match_oracle(K):
    row_ptr, col_idx = random_csr(300, 200, 12, seed=K)
    fals, port = _models(300, 200, row_ptr, col_idx, K)
    assert np.array_equal(fals.U, port.U)            # same libstdc++ stream
    assert np.array_equal(fals.V, port.V)
    assert np.allclose(fals.Wi, port.Wi, rtol=1e-15, atol=0)
    for got, want in ((fals.SU, port.SU), (fals.SV, port.SV)):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("K", [8, 64, 128])
@pytest.mark.parametrize("shape", [(400, 300, 14), (120, 900, 150), (900, 60, 10)])
def test_half_epochs_and_loss_match_oracle(K, shape):
    M, N, dens = shape
    row_ptr, col_idx = random_csr(M, N, dens, seed=M + K, empty_frac=0.05)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for it in range(3):
        fals.update_user(); port.update_user()
        assert np.abs(fals.U - port.U).max() < 1e-11, f"U after user sweep {it}"
        fals.update_item(); port.update_item()
        assert np.abs(fals.V - port.V).max() < 1e-11, f"V after item sweep {it}"
        lg, lc = fals.loss(), port.loss()
        assert abs(lg - lc) <= 1e-10 * abs(lc), (it, lg, lc)
    assert np.abs(fals.SU - port.SU).max() <= 1e-11 * np.abs(port.SU).max()
    assert np.abs(fals.SV - port.SV).max() <= 1e-11 * np.abs(port.SV).max()


@pytest.mark.parametrize("mode", ["cache", "nocache", "refresh2"])
def test_prediction_cache_modes_agree_with_oracle(mode, monkeypatch):
    """The symmetric prediction cache (default on one GPU), the recompute-every-sweep path that
    multi-rank models use, and periodic refresh all track the oracle over several epochs, across
    all three kernel families (warp rows, one-CTA rows, slab pipeline)."""
    if mode == "nocache":
        monkeypatch.setenv("EALS_NO_PRED_CACHE", "1")
    if mode == "refresh2":
        monkeypatch.setenv("EALS_PRED_REFRESH_EVERY", "2")
    M, N, K = 2500, 80, 32
    row_ptr, col_idx = _heavy_matrix(M, N, heavy_cols=[1, 9], heavy_len=1800, seed=3)
    row_ptr2, col_idx2 = _heavy_matrix(M, N, heavy_cols=[4, 5, 6], heavy_len=300, seed=4)
    rows = [sorted(set(col_idx[row_ptr[u]:row_ptr[u + 1]]) | set(col_idx2[row_ptr2[u]:row_ptr2[u + 1]])) for u in range(M)]
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col_idx = np.concatenate([np.array(r, np.int32) for r in rows])
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for it in range(5):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
        assert np.abs(fals.U - port.U).max() < 1e-10, it
        assert np.abs(fals.V - port.V).max() < 1e-10, it
        lg, lc = fals.loss(), port.loss()
        assert abs(lg - lc) <= 1e-10 * abs(lc)
    # factors replaced from outside: the cache must be dropped, not reused
    U, V = port.U * 0.5, port.V * 2.0
    port.U[:], port.V[:] = U, V
    port.init_S()
    fals.setUV(U, V)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    assert np.abs(fals.U - port.U).max() < 1e-10
    assert np.abs(fals.V - port.V).max() < 1e-10
    # single-row updates also invalidate it
    fals.update_user_thread(3); port.update_user(3, 4)
    fals.refresh_S(); port.init_S()
    fals.update_item(); port.update_item()
    assert np.abs(fals.V - port.V).max() < 1e-10


def test_set_train_replaces_the_matrix_and_keeps_weights():
    """setTrain (MF_fastALS.cpp:94-104): new matrix of the same shape, factors and Wi kept; device
    buffers are reused, the prediction cache is rebuilt."""
    from eals_cpp_b200.model import SparseMat
    from oracle.bindings import PortModel
    M, N, K = 700, 500, 32
    rpA, ciA = random_csr(M, N, 15, seed=1)
    rpB, ciB = random_csr(M, N, 22, seed=2, empty_frac=0.1)      # more nonzeros than A: buffers must grow
    fals, port = _models(M, N, rpA, ciA, K)
    fals.update_user(); fals.update_item()
    for rp, ci in ((rpB, ciB), (rpA, ciA)):
        fals.setTrain(SparseMat.from_csr(M, N, rp, ci))
        port = PortModel(M, N, rp, ci, factors=K)
        port.U[:], port.V[:], port.Wi[:] = fals.U, fals.V, fals.Wi
        port.init_S()
        for _ in range(2):
            fals.update_user(); port.update_user()
            fals.update_item(); port.update_item()
        assert np.abs(fals.U - port.U).max() < 1e-10
        assert np.abs(fals.V - port.V).max() < 1e-10
        lg, lc = fals.loss(), port.loss()          # loss() streams the prediction cache here
        assert abs(lg - lc) <= 1e-10 * abs(lc)


def test_weighted_ratings_match_oracle():
    """Non-unit ratings: W is a copy of the rating values (MF_fastALS.cpp:75-82)."""
    M, N, K = 200, 150, 16
    row_ptr, col_idx = random_csr(M, N, 10, seed=5)
    val = np.random.default_rng(1).uniform(0.5, 3.0, size=len(col_idx))
    fals, port = _models(M, N, row_ptr, col_idx, K, val=val)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    assert np.abs(fals.U - port.U).max() < 1e-11
    assert np.abs(fals.V - port.V).max() < 1e-11
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)


def test_long_rows_use_row_block_path():
    """Rows of 129..1500 nonzeros (one-CTA blocked kernel and the slab pipeline) and dense columns."""
    M, N, K = 64, 2000, 64
    rng = np.random.default_rng(3)
    row_ptr, cols = [0], []
    for u in range(M):
        n = 1500 if u % 7 == 0 else int(rng.integers(1, 300))
        cols.append(np.sort(rng.choice(N, size=n, replace=False)).astype(np.int32))
        row_ptr.append(row_ptr[-1] + n)
    row_ptr, col_idx = np.asarray(row_ptr, np.int64), np.concatenate(cols)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    assert np.abs(fals.U - port.U).max() < 1e-10
    assert np.abs(fals.V - port.V).max() < 1e-10
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)


@pytest.mark.parametrize("K", [16, 64])
def test_heavy_rows_slab_pipeline(K, monkeypatch):
    """Columns of 1300..5000 nonzeros (> 1024: slab pipeline), several batches, plus one-CTA rows."""
    monkeypatch.setenv("EALS_HEAVY_BATCH_NNZ", "6000")
    M, N = 6000, 60
    row_ptr, col_idx = _heavy_matrix(M, N, heavy_cols=[3, 7, 11, 20, 33], heavy_len=1300, seed=K)
    row_ptr2, col_idx2 = _heavy_matrix(M, N, heavy_cols=[5], heavy_len=5000, seed=K + 1)
    # merge the two patterns
    rows = [sorted(set(col_idx[row_ptr[u]:row_ptr[u + 1]]) | set(col_idx2[row_ptr2[u]:row_ptr2[u + 1]])) for u in range(M)]
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col_idx = np.concatenate([np.array(r, np.int32) for r in rows])
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        assert np.abs(fals.U - port.U).max() < 1e-10
        fals.update_item(); port.update_item()
        assert np.abs(fals.V - port.V).max() < 1e-10
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)
    # single-row API on a heavy row and on a one-CTA row
    fals.update_item_thread(5); port.update_item(5, 6)
    fals.update_item_thread(3); port.update_item(3, 4)
    assert np.abs(fals.V - port.V).max() < 1e-10


def test_ultra_heavy_row_two_level_reduction():
    """One column rated by 270k users: 528 slabs -> the grouped (two-level) partial reduction."""
    M, N, K = 270_000, 6, 16
    rng = np.random.default_rng(12)
    extra = rng.integers(1, N, size=M).astype(np.int32)
    row_ptr = np.arange(0, 2 * M + 1, 2, dtype=np.int64)
    col_idx = np.empty(2 * M, np.int32)
    col_idx[0::2] = 0
    col_idx[1::2] = extra
    fals, port = _models(M, N, row_ptr, col_idx, K)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    assert np.abs(fals.U - port.U).max() < 1e-10
    assert np.abs(fals.V - port.V).max() < 1e-9
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)


def test_slab_launch_order_does_not_change_a_single_bit(monkeypatch):
    """The heavy slabs are launched in neighbour order for L2 reuse; partial sums are stored and added
    by canonical slot, so the result must be bit-identical to the canonical launch order."""
    M, N, K = 5000, 40, 32
    row_ptr, col_idx = _heavy_matrix(M, N, heavy_cols=[1, 9, 17, 30], heavy_len=2500, seed=5)
    out = []
    for order in ("1", "0"):
        monkeypatch.setenv("EALS_HEAVY_ORDER", order)
        fals, _ = _models(M, N, row_ptr, col_idx, K)
        for _ in range(2):
            fals.update_user(); fals.update_item()
        out.append((fals.U, fals.V))
        fals.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_rows_of_every_bucket_boundary_length():
    """Row lengths on both sides of every kernel-family boundary (32/33, 64/65, 128/129, 256/257,
    512/513) in one matrix, K not a multiple of 16."""
    lens = [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 700]
    M, N, K = len(lens) * 3, 1200, 40
    rng = np.random.default_rng(21)
    row_ptr, cols = [0], []
    for u in range(M):
        n = lens[u % len(lens)]
        cols.append(np.sort(rng.choice(N, size=n, replace=False)).astype(np.int32))
        row_ptr.append(row_ptr[-1] + n)
    row_ptr, col_idx = np.asarray(row_ptr, np.int64), np.concatenate(cols)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        assert np.abs(fals.U - port.U).max() < 1e-10
        fals.update_item(); port.update_item()
        assert np.abs(fals.V - port.V).max() < 1e-10
    lg, lc = fals.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)


def test_single_row_api_and_patches():
    M, N, K = 150, 120, 16
    row_ptr, col_idx = random_csr(M, N, 9, seed=9)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    u = 17
    old = fals.U[u].copy()
    fals.update_user_thread(u)
    port.update_user(u, u + 1)                   # oracle: same row, SU patched for that row
    new = fals.U[u]
    assert np.abs(new - port.U[u]).max() < 1e-12
    fals.update_user_SU(old, new)
    assert np.abs(fals.SU - port.SU).max() <= 1e-12 * np.abs(port.SU).max()
    i = 5
    oldv = fals.V[i].copy()
    fals.update_item_thread(i)
    port.update_item(i, i + 1)
    assert np.abs(fals.V[i] - port.V[i]).max() < 1e-12
    fals.update_item_SV(i, oldv, fals.V[i])
    assert np.abs(fals.SV - port.SV).max() <= 1e-12 * np.abs(port.SV).max()
    assert fals.predict(3, 4) == port_predict(port, 3, 4)   # sequential-k, no FMA: bit-identical


def port_predict(port, u, i):
    acc = 0.0
    for k in range(port.K):
        acc += port.U[u, k] * port.V[i, k]
    return acc


@pytest.mark.parametrize("first_chunk", [None, "64"])
@pytest.mark.parametrize("scale", [1.0, 40.0])
def test_evaluate_matches_oracle_bug_for_bug(scale, first_chunk, monkeypatch):
    """scale=40 inflates the factors so that truncated scores are non-zero (and some negative):
    exercises the heap replay beyond the all-zero-keys case.  first_chunk=64 forces several rounds
    of the chunked early-out scan (64, 256, ... items) on this small catalogue."""
    if first_chunk:
        monkeypatch.setenv("EALS_EVAL_FIRST_CHUNK", first_chunk)
    M, N, K, topK = 500, 333, 16, 10
    row_ptr, col_idx = random_csr(M, N, 12, seed=21)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    if scale != 1.0:
        rng = np.random.default_rng(2)
        U = port.U * scale + rng.normal(0, 0.5, port.U.shape)
        V = port.V * scale + rng.normal(0, 0.5, port.V.shape)
        port.U[:], port.V[:] = U, V
        fals.setUV(U, V)
    gt = np.random.default_rng(4).integers(0, N, size=M).astype(np.int32)
    gt[:40] = np.arange(40) % 12                  # some gt items inside the first topK ids
    for compat in (True, False):
        want_mean, whr, wndcg, wprec, wcnt = port.evaluate(gt, topK, compat=compat)
        got_mean, hr, ndcg, prec, cnt = fals.evaluate(gt, topK, exact=not compat, per_user=True)
        assert np.array_equal(cnt, wcnt)
        assert np.array_equal(hr, whr) and np.array_equal(ndcg, wndcg) and np.array_equal(prec, wprec)
        assert np.allclose(got_mean, want_mean, rtol=0, atol=1e-15)
    _, whr, wndcg, wprec, _ = port.evaluate(gt, topK, compat=True)
    for u in (0, 7, 39, 123):
        assert fals.evaluate_for_user(u, int(gt[u]), topK) == [whr[u], wndcg[u], wprec[u]]


@pytest.mark.parametrize("patch_S", [True, False])
def test_online_update_model(patch_S):
    """updateModel (MF_fastALS.cpp:223-242): new interaction + 10 alternating single-row updates, with
    the S caches patched after every row (intended behaviour) or left stale (the reference's arithmetic).
    Checked against the oracle's single-row sweeps on the matrix with the entry inserted; one case adds
    the first rating of an item that had none (weight w0 / itemCount, SV rebuilt)."""
    M, N, K = 300, 200, 16
    row_ptr, col_idx = random_csr(M, N, 12, seed=9, empty_frac=0.02)
    # make item 7 unseen so that the "new item" branch runs
    rows0 = [[c for c in col_idx[row_ptr[r]:row_ptr[r + 1]] if c != 7] for r in range(M)]
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows0])]).astype(np.int64)
    col_idx = np.array([c for r in rows0 for c in r], np.int32)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    fals.update_user(); port.update_user(); port.SU = port.p.gram_plain(port.U)
    fals.update_item(); port.update_item(); port.SV = port.p.gram_weighted(port.V, port.Wi)
    for (u, i) in [(5, 7), (11, 40)]:
        fals.updateModel(u, i, patch_S=patch_S)
        # oracle: insert the entry, then the same single-row schedule
        rows = [list(port.col_idx[port.row_ptr[r]:port.row_ptr[r + 1]]) for r in range(M)]
        if i not in rows[u]:
            rows[u] = sorted(rows[u] + [i])
        rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        ci = np.concatenate([np.array(r, np.int32) for r in rows])
        from oracle.bindings import csr_to_csc
        port.row_ptr, port.col_idx = rp, ci
        port.col_ptr, port.row_idx, port.cval, _ = csr_to_csc(M, N, rp, ci, None)
        if port.Wi[i] == 0.0:
            port.Wi[i] = 10.0 / N
            port.SV = port.p.gram_weighted(port.V, port.Wi)
        for _ in range(10):   # the oracle's row sweeps patch SU / SV themselves (MF_fastALS.cpp:127-132, 146-152)
            su, sv = port.SU.copy(), port.SV.copy()
            port.update_user(u, u + 1)
            if not patch_S:
                port.SU[...] = su
            port.update_item(i, i + 1)
            if not patch_S:
                port.SV[...] = sv
        assert np.abs(fals.U - port.U).max() < 1e-10
        assert np.abs(fals.V - port.V).max() < 1e-10
        assert np.abs(fals.SU - port.SU).max() <= 1e-10 * np.abs(port.SU).max()
        assert np.abs(fals.SV - port.SV).max() <= 1e-10 * np.abs(port.SV).max()


def test_errors_are_reported_not_swallowed():
    from eals_cpp_b200._lib import EalsError
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    row_ptr = np.array([0, 2, 3], np.int64)
    col_idx = np.array([1, 0, 2], np.int32)         # row 0 not ascending
    sm = SparseMat.from_csr(2, 3, row_ptr, col_idx)
    sm.col_idx = col_idx                             # keep the bad order
    with pytest.raises(EalsError):
        MF_fastALS(sm, None, factors=8)


@pytest.mark.parametrize("K", [129, 200, 256])
def test_factors_above_128_all_kernel_families(K):
    """K in 129..256 takes its own code paths (LD = 256: S cache read from global memory in the warp
    kernels, two block pairs in the Gram): short, mid and heavy rows, S caches, loss and evaluation."""
    lens = [1, 5, 31, 32, 33, 64, 100, 128, 129, 200, 256, 300, 384, 512, 513, 900]
    M, N = len(lens) * 4, 1000
    rng = np.random.default_rng(K)
    row_ptr, cols = [0], []
    for u in range(M):
        n = lens[u % len(lens)]
        c = set(rng.choice(N, size=n, replace=False).tolist())
        c.add(7)                                   # column 7 is rated by everybody -> item-side rows differ in length too
        if u % 2:
            c.add(11)
        cols.append(np.array(sorted(c), np.int32))
        row_ptr.append(row_ptr[-1] + len(c))
    row_ptr, col_idx = np.asarray(row_ptr, np.int64), np.concatenate(cols)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    assert np.array_equal(fals.U, port.U) and np.array_equal(fals.V, port.V)
    for got, want in ((fals.SU, port.SU), (fals.SV, port.SV)):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    for it in range(2):
        fals.update_user(); port.update_user()
        assert np.abs(fals.U - port.U).max() < 1e-10, it
        fals.update_item(); port.update_item()
        assert np.abs(fals.V - port.V).max() < 1e-10, it
        lg, lc = fals.loss(), port.loss()
        assert abs(lg - lc) <= 1e-10 * abs(lc), (it, lg, lc)
    assert np.abs(fals.SU - port.SU).max() <= 1e-11 * np.abs(port.SU).max()
    assert np.abs(fals.SV - port.SV).max() <= 1e-11 * np.abs(port.SV).max()
    gt = rng.integers(0, N, size=M).astype(np.int32)
    for compat in (True, False):
        want_mean, whr, wndcg, wprec, wcnt = port.evaluate(gt, 10, compat=compat)
        got_mean, hr, ndcg, prec, cnt = fals.evaluate(gt, 10, exact=not compat, per_user=True)
        assert np.array_equal(cnt, wcnt)
        assert np.array_equal(hr, whr) and np.array_equal(ndcg, wndcg) and np.array_equal(prec, wprec)


def test_loss_with_a_stale_SU_cache_follows_the_reference():
    """update_user_thread leaves SU alone (MF_fastALS.cpp:243-322); the reference's loss() takes
    sum_u u^T SV u from U itself (:199-200), so it stays right while SU is stale.  Ours must too."""
    M, N, K = 180, 140, 16
    row_ptr, col_idx = random_csr(M, N, 9, seed=31)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    su_before = fals.SU
    for u in (3, 77, 120):
        fals.update_user_thread(u)                 # no update_user_SU
        port.update_user(u, u + 1)
    assert np.array_equal(fals.SU, su_before)      # our SU is stale, like the reference's
    lg, lc = fals.loss(), port.loss()              # the oracle's loss never reads SU
    assert abs(lg - lc) <= 1e-10 * abs(lc), (lg, lc)
    assert np.array_equal(fals.SU, su_before)      # ... and loss() did not refresh it behind the caller's back
    fals.updateModel(9, 33, patch_S=False)         # the reference's stale-cache arithmetic end to end
    import io
    fals.out = io.StringIO()
    got = fals.showLoss(0, 0.0, float("inf"))
    # ground truth: the reference's formula on the grown matrix with the model's own (stale) SV — :193-200
    tm = fals.trainMatrix
    want = port.p.loss(np.ascontiguousarray(tm.row_ptr, np.int64), np.ascontiguousarray(tm.col_idx, np.int32), None,
                       fals.U, fals.V, fals.SV, fals.Wi, 0.01)
    assert abs(got - want) <= 1e-10 * abs(want)


def test_checkpoint_save_load_resume(tmp_path):
    """save / load of U, V, Wi: a model rebuilt from the checkpoint resumes training on
    the interrupted run's trajectory."""
    from eals_cpp_b200._lib import EalsError
    M, N, K = 400, 260, 20
    row_ptr, col_idx = random_csr(M, N, 11, seed=77, empty_frac=0.04)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    w = fals.Wi
    w[5] *= 1.5                                    # a non-default weight must survive the round trip
    fals.Wi = w; port.Wi[:] = w; port.init_S()
    path = str(tmp_path / "factors.eals")
    fals.save(path)
    U2, V2 = fals.U, fals.V
    fals.close()
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    again = MF_fastALS(SparseMat.from_csr(M, N, row_ptr, col_idx), None, factors=K, showLoss=False, init=False)
    again.load(path)
    assert np.array_equal(again.U, U2) and np.array_equal(again.V, V2) and np.array_equal(again.Wi, w)
    for _ in range(2):
        again.update_user(); port.update_user()
        again.update_item(); port.update_item()
    assert np.abs(again.U - port.U).max() < 1e-10 and np.abs(again.V - port.V).max() < 1e-10
    lg, lc = again.loss(), port.loss()
    assert abs(lg - lc) <= 1e-10 * abs(lc)
    other = MF_fastALS(SparseMat.from_csr(M, N, row_ptr, col_idx), None, factors=K + 4, showLoss=False, init=False)
    with pytest.raises(EalsError):
        other.load(path)                           # wrong K
    with pytest.raises(EalsError):
        other.load(str(tmp_path / "missing.eals"))


def test_update_model_overwrites_an_existing_rating_with_one():
    """trainMatrix.setValue(u, i, 1) / W.setValue(u, i, w_new) (MF_fastALS.cpp:224-226) on an entry that
    already exists: rating and weight become 1."""
    from oracle.bindings import csr_to_csc
    M, N, K = 120, 90, 16
    row_ptr, col_idx = random_csr(M, N, 8, seed=13)
    val = np.random.default_rng(3).uniform(0.5, 3.0, size=len(col_idx))
    fals, port = _models(M, N, row_ptr, col_idx, K, val=val)
    fals.update_user(); port.update_user(); port.SU = port.p.gram_plain(port.U)
    fals.update_item(); port.update_item(); port.SV = port.p.gram_weighted(port.V, port.Wi)
    u = 10
    i = int(col_idx[row_ptr[u]])                   # an existing entry with a non-unit rating
    assert val[row_ptr[u]] != 1.0
    fals.updateModel(u, i)
    val2 = val.copy(); val2[row_ptr[u]] = 1.0
    port.val = val2
    port.col_ptr, port.row_idx, port.cval, _ = csr_to_csc(M, N, row_ptr, col_idx, val2)
    for _ in range(10):
        port.update_user(u, u + 1)
        port.update_item(i, i + 1)
    assert np.abs(fals.U - port.U).max() < 1e-10 and np.abs(fals.V - port.V).max() < 1e-10
    assert fals.trainMatrix.row_val[row_ptr[u]] == 1.0


@pytest.mark.parametrize("K", [8, 64, 128, 130, 200])
def test_tensor_core_filter_equals_exact_scan(K, monkeypatch):
    """K3 on tcgen05 (csrc/eval_tc.cuh): fp16 tensor-core scores with a rigorous error bound decide the
    clear cases, every close call is re-scored in fp64 — count_larger / HR / NDCG / reciprocal rank per user
    must be IDENTICAL to the all-fp64 tile scan (EALS_EVAL_SCALAR=1) and to the oracle.  Includes exact ties
    (duplicated item rows score exactly like the held-out item: strict '>' must not count them), a catalogue
    that is not a multiple of the 128-item tile, several item blocks, and inflated factors (non-zero
    int-truncated keys for the ranking replay)."""
    M, N, topK = 1000, 777, 10
    row_ptr, col_idx = random_csr(M, N, 14, seed=K)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    for _ in range(2):
        fals.update_user(); port.update_user()
        fals.update_item(); port.update_item()
    rng = np.random.default_rng(K + 1)
    gt = rng.integers(0, N, size=M).astype(np.int32)
    for variant in ("trained", "ties", "inflated"):
        U, V = port.U.copy(), port.V.copy()
        if variant == "ties":
            for u in range(0, M, 3):                # items scoring EXACTLY like the held-out one
                V[(gt[u] + 1 + np.arange(5) * 7) % N] = V[gt[u]]
        if variant == "inflated":
            U = U * 30 + rng.normal(0, 0.4, U.shape)
            V = V * 30 + rng.normal(0, 0.4, V.shape)
        port.U[:], port.V[:] = U, V
        fals.setUV(U, V)
        for first_chunk in ("128", None):
            if first_chunk:
                monkeypatch.setenv("EALS_EVAL_FIRST_CHUNK", first_chunk)
            else:
                monkeypatch.delenv("EALS_EVAL_FIRST_CHUNK", raising=False)
            for compat in (True, False):
                want = port.evaluate(gt, topK, compat=compat)
                monkeypatch.delenv("EALS_EVAL_SCALAR", raising=False)
                got = fals.evaluate(gt, topK, exact=not compat, per_user=True)
                assert fals.eval_stats()["engine"] == "tcgen05"
                monkeypatch.setenv("EALS_EVAL_SCALAR", "1")
                ref = fals.evaluate(gt, topK, exact=not compat, per_user=True)
                assert fals.eval_stats()["engine"] == "fp64"
                monkeypatch.delenv("EALS_EVAL_SCALAR", raising=False)
                for k in range(1, 5):
                    assert np.array_equal(got[k], ref[k]), (variant, compat, k)
                assert np.array_equal(got[4], want[4]) and np.array_equal(got[1], want[1])
                assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])


def test_tensor_core_filter_pair_overflow_falls_back_to_exact(monkeypatch):
    """All-equal scores (zero factors): every item is a candidate for every user; with a tiny pair buffer the
    filter reports overflow and the exact engine takes over — same answer."""
    M, N, K = 300, 200, 16
    row_ptr, col_idx = random_csr(M, N, 8, seed=3)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    Z = np.zeros_like(port.U)
    port.U[:] = Z
    fals.setUV(Z, port.V)
    gt = np.arange(M, dtype=np.int32) % N
    monkeypatch.setenv("EALS_EVAL_PAIR_CAP", "1000")
    monkeypatch.setenv("EALS_EVAL_PAIR_HARD_CAP", "5000")
    got = fals.evaluate(gt, 10, per_user=True)
    want = port.evaluate(gt, 10, compat=True)
    assert fals.eval_stats()["engine"] == "fp64"
    assert np.array_equal(got[4], want[4]) and np.array_equal(got[1], want[1])


def test_epochs_replayed_as_a_cuda_graph_are_bit_identical(monkeypatch):
    """eals_run_epochs(use_graph=1): one captured epoch replayed — same kernels in the same order, so U and V must
    equal the ordinary update_user / update_item calls to the last bit; the graph is rebuilt after setTrain."""
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    M, N, K = 2500, 80, 32
    row_ptr, col_idx = _heavy_matrix(M, N, heavy_cols=[1, 9], heavy_len=1800, seed=3)     # warp rows + slab pipeline
    row_ptr2, col_idx2 = _heavy_matrix(M, N, heavy_cols=[4, 5, 6], heavy_len=300, seed=4)   # + one-CTA rows
    rows = [sorted(set(col_idx[row_ptr[u]:row_ptr[u + 1]]) | set(col_idx2[row_ptr2[u]:row_ptr2[u + 1]])) for u in range(M)]
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    col_idx = np.concatenate([np.array(r, np.int32) for r in rows])
    sm = SparseMat.from_csr(M, N, row_ptr, col_idx)
    a = MF_fastALS(sm, None, factors=K, showLoss=False)
    b = MF_fastALS(sm, None, factors=K, showLoss=False)
    for _ in range(5):
        a.update_user(); a.update_item()
    launches0 = b.kernel_launches()
    b.run_epochs(5, graph=True)
    assert np.array_equal(a.U, b.U) and np.array_equal(a.V, b.V)
    assert b.kernel_launches() > launches0
    la, lb = a.loss(), b.loss()
    assert la == lb
    # a new matrix: the captured graph must not be replayed on stale pointers / sizes
    sm2 = SparseMat.from_csr(M, N, row_ptr2, col_idx2)
    a.setTrain(sm2); b.setTrain(sm2)
    for _ in range(3):
        a.update_user(); a.update_item()
    b.run_epochs(3, graph=True)
    assert np.array_equal(a.U, b.U) and np.array_equal(a.V, b.V)
    # factors replaced from outside between graph runs: the cache state is re-settled by an ordinary epoch first
    U, V = a.U * 0.5, a.V * 2.0
    a.setUV(U, V); b.setUV(U, V)
    for _ in range(2):
        a.update_user(); a.update_item()
    b.run_epochs(2, graph=True)
    assert np.array_equal(a.U, b.U) and np.array_equal(a.V, b.V)


@pytest.mark.parametrize("shape", [(128, 50, 1, 1), (129, 128, 5, 10), (257, 129, 64, 3), (300, 1000, 16, 60), (640, 127, 33, 200)])
def test_tensor_core_filter_edge_shapes(shape, monkeypatch):
    """Edges of the tcgen05 evaluation: exactly / just over one 128-user tile, catalogues smaller than, equal to and
    just over one 128-item tile, K = 1, K not a multiple of 16, topK larger than the catalogue — against the oracle
    and the exact fp64 scan, per user."""
    M, N, K, topK = shape
    rng = np.random.default_rng(M + N + K)
    row_ptr, col_idx = random_csr(M, N, max(2, min(10, N // 4)), seed=M + K)
    fals, port = _models(M, N, row_ptr, col_idx, K)
    fals.update_user(); port.update_user()
    fals.update_item(); port.update_item()
    gt = rng.integers(0, N, size=M).astype(np.int32)
    for scale in (1.0, 25.0):
        if scale != 1.0:
            U = port.U * scale + rng.normal(0, 0.3, port.U.shape)
            V = port.V * scale + rng.normal(0, 0.3, port.V.shape)
            port.U[:], port.V[:] = U, V
            fals.setUV(U, V)
        for compat in (True, False):
            want = port.evaluate(gt, topK, compat=compat)
            monkeypatch.delenv("EALS_EVAL_SCALAR", raising=False)
            got = fals.evaluate(gt, topK, exact=not compat, per_user=True)
            assert fals.eval_stats()["engine"] == "tcgen05"
            for k in range(1, 5):
                assert np.array_equal(got[k], want[k]), (shape, scale, compat, k)
            assert np.allclose(got[0], want[0], rtol=0, atol=1e-12)
