"""Synthetic implicit-feedback matrices of the shapes BASELINE.json names.

The reference's only data source is the text file ``yelp.rating`` (main.cpp:73-130), which is not
shipped; everything here is synthetic and seeded.  A matrix is returned in the layout contract of
the product: CSR ``row_ptr[M+1]`` (int64) / ``col_idx[nnz]`` (int32, ascending inside each row, no
duplicates) — exactly the order main.cpp:198-205 fills ``SparseMat.rows`` in — plus one held-out
test item per user (main.cpp:175-180).  All ratings are 1, as the reference's loader stores them
(main.cpp:184-185,202).

Two generators:
  * ``powerlaw_csr``        numpy, host — tests and the small configs (C1/C2-shaped).
  * ``powerlaw_csr_device`` torch on the current CUDA device — C3/C4-shaped matrices, where a host
                            generator would take minutes (torch here is plumbing: RNG, sort, unique).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# name -> (M, N, target nnz, K, topK, user-degree sigma, degree clip, zipf exponent, seed)
WORKLOADS = {
    # tiny cases for tests
    "tiny": dict(M=400, N=300, nnz=5400, K=8, topK=10, sigma=0.8, clip=(1, 120), zipf=0.9, seed=7),
    "small": dict(M=3000, N=2000, nnz=90_000, K=64, topK=10, sigma=1.0, clip=(1, 800), zipf=0.9, seed=11),
    # C1: yelp.rating shape (Outputs.txt:9-11), run.sh defaults K=64
    "c1": dict(M=25_677, N=25_815, nnz=672_795, K=64, topK=10, sigma=1.0, clip=(1, 4000), zipf=0.9, seed=20261018),
    # C2: same matrix, K=128
    "c2": dict(M=25_677, N=25_815, nnz=672_795, K=128, topK=10, sigma=1.0, clip=(1, 4000), zipf=0.9, seed=20261018),
    # C3: MovieLens-20M shape
    "c3": dict(M=138_493, N=26_744, nnz=20_000_263, K=64, topK=10, sigma=1.0, clip=(20, 10_000), zipf=1.0, seed=20261019),
    # C4: Amazon scale. Column length is capped (SURVEY §8d allows "cap column length at e.g. 2M").
    "c4": dict(M=10_000_000, N=2_000_000, nnz=500_000_000, K=128, topK=100, sigma=1.0, clip=(1, 5000), zipf=1.0, seed=20261020, col_cap=2_000_000),
    # mid-size stand-in used when a GPU has too little free memory for c4
    "c4s": dict(M=2_000_000, N=400_000, nnz=100_000_000, K=128, topK=100, sigma=1.0, clip=(1, 5000), zipf=1.0, seed=20261021, col_cap=400_000),
}


@dataclass
class Interactions:
    M: int
    N: int
    row_ptr: np.ndarray      # int64 [M+1]
    col_idx: np.ndarray      # int32 [nnz]
    test_items: np.ndarray   # int32 [M]

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])


def _degrees(rng, M, nnz, sigma, clip):
    """Log-normal user degrees rescaled so they sum to ~nnz, clipped to [lo, hi]."""
    lo, hi = clip
    d = rng.lognormal(mean=0.0, sigma=sigma, size=M)
    d *= nnz / d.sum()
    d = np.clip(np.rint(d), lo, hi)
    # one cheap correction pass for the clipping
    d = np.clip(np.rint(d * (nnz / d.sum())), lo, hi)
    return d.astype(np.int64)


def _item_cdf(rng, N, zipf, col_cap_frac=None):
    """Zipf-like popularity p(r) ~ r^-zipf over a seeded permutation of item ids."""
    p = np.arange(1, N + 1, dtype=np.float64) ** (-zipf)
    p /= p.sum()
    if col_cap_frac is not None:
        for _ in range(8):                       # water-fill the clipped mass onto the tail
            over = p > col_cap_frac
            if not over.any():
                break
            excess = (p[over] - col_cap_frac).sum()
            p[over] = col_cap_frac
            p[~over] += excess * p[~over] / p[~over].sum()
    perm = rng.permutation(N)
    return np.cumsum(p), perm


def powerlaw_csr(M, N, nnz, sigma=1.0, clip=(1, 4000), zipf=0.9, seed=0, calibrate=2, **_) -> Interactions:
    """Host generator. Draw degree+1 items per user (with replacement), dedup, hold the last
    draw out as the test item.  Popular items are drawn repeatedly, so de-duplication loses
    nonzeros; `calibrate` extra rounds rescale the number of draws until the train matrix holds
    ~nnz entries."""
    draws_total = nnz
    for _ in range(calibrate + 1):
        out = _powerlaw_csr_once(M, N, draws_total, nnz, sigma, clip, zipf, seed)
        if abs(out.nnz - nnz) <= 0.01 * nnz:
            break
        draws_total = int(draws_total * nnz / max(out.nnz, 1))
    return out


def _powerlaw_csr_once(M, N, draws_total, nnz, sigma, clip, zipf, seed) -> Interactions:
    rng = np.random.default_rng(seed)
    deg = _degrees(rng, M, draws_total, sigma, (clip[0], min(clip[1] * max(1, draws_total // nnz), N - 1)))
    cdf, perm = _item_cdf(rng, N, zipf)
    draws = deg + 1
    owner = np.repeat(np.arange(M, dtype=np.int64), draws)
    items = perm[np.minimum(np.searchsorted(cdf, rng.random(owner.size)), N - 1)].astype(np.int64)
    # the LAST draw of each user is the test candidate
    last = np.cumsum(draws) - 1
    test_items = items[last].astype(np.int32)
    keep = np.ones(owner.size, bool)
    keep[last] = False
    key = np.unique(owner[keep] * N + items[keep])      # sorted => rows asc, items asc; dedup
    rows = key // N
    col_idx = (key % N).astype(np.int32)
    row_ptr = np.zeros(M + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=row_ptr[1:])
    return Interactions(M, N, row_ptr, col_idx, test_items)


def powerlaw_csr_device(M, N, nnz, sigma=1.0, clip=(1, 5000), zipf=1.0, seed=0, col_cap=None,
                        device="cuda", chunk=64_000_000, calibrate=2, **_):
    """Device generator (torch tensors on ``device``): returns (row_ptr int64, col_idx int32,
    test_items int32).  Same scheme as ``powerlaw_csr`` (incl. the calibration rounds that make up
    for de-duplication losses); generated in chunks of users so the sort keys stay well inside HBM."""
    draws_total = nnz
    for _ in range(calibrate + 1):
        row_ptr, col_idx, test_items = _powerlaw_csr_device_once(M, N, draws_total, nnz, sigma, clip, zipf, seed,
                                                                 col_cap, device, chunk)
        got = int(row_ptr[-1])
        if abs(got - nnz) <= 0.01 * nnz:
            break
        del col_idx
        draws_total = int(draws_total * nnz / max(got, 1))
    return row_ptr, col_idx, test_items


def _powerlaw_csr_device_once(M, N, draws_total, nnz, sigma, clip, zipf, seed, col_cap, device, chunk):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    deg_np = _degrees(rng, M, draws_total, sigma, (clip[0], min(clip[1] * max(1, draws_total // nnz), N - 1)))
    cdf_np, perm_np = _item_cdf(rng, N, zipf, None if col_cap is None else col_cap / max(nnz, 1))
    cdf = torch.from_numpy(cdf_np).to(device)
    perm = torch.from_numpy(perm_np).to(device)
    deg = torch.from_numpy(deg_np).to(device)
    draws = deg + 1
    test_items = torch.empty(M, dtype=torch.int32, device=device)
    counts = torch.zeros(M, dtype=torch.int64, device=device)
    cols = []
    cum = np.concatenate([[0], np.cumsum(deg_np + 1)])
    u0 = 0
    while u0 < M:
        u1 = int(np.searchsorted(cum, cum[u0] + chunk, side="right"))
        u1 = min(max(u1 - 1, u0 + 1), M)
        d = draws[u0:u1]
        owner = torch.repeat_interleave(torch.arange(u0, u1, device=device, dtype=torch.int64), d)
        r = torch.rand(owner.numel(), generator=g, device=device, dtype=torch.float64)
        it = perm[torch.clamp(torch.searchsorted(cdf, r), max=N - 1)]
        last = torch.cumsum(d, 0) - 1
        test_items[u0:u1] = it[last].to(torch.int32)
        keep = torch.ones(owner.numel(), dtype=torch.bool, device=device)
        keep[last] = False
        key = torch.unique(owner[keep] * N + it[keep])            # sorted + dedup
        del owner, r, it, keep
        rows = torch.div(key, N, rounding_mode="floor")
        counts[u0:u1] = torch.bincount(rows - u0, minlength=u1 - u0)
        cols.append((key - rows * N).to(torch.int32))
        del key, rows
        u0 = u1
    col_idx = torch.cat(cols)
    row_ptr = torch.zeros(M + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    return row_ptr, col_idx, test_items


def csr_to_csc_device(M, N, row_ptr, col_idx):
    """Device CSR->CSC transpose (rows ascending inside each column): one stable sort by column."""
    import torch

    rows = torch.repeat_interleave(torch.arange(M, device=row_ptr.device, dtype=torch.int32),
                                   row_ptr[1:] - row_ptr[:-1])
    order = torch.argsort(col_idx.to(torch.int64) * M + rows.to(torch.int64))
    row_idx = rows[order].contiguous()
    col_ptr = torch.zeros(N + 1, dtype=torch.int64, device=row_ptr.device)
    torch.cumsum(torch.bincount(col_idx.to(torch.int64), minlength=N), 0, out=col_ptr[1:])
    return col_ptr, row_idx, order


def make(name: str) -> Interactions:
    spec = WORKLOADS[name]
    return powerlaw_csr(**spec)
