"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU checkers.

* ``Port``      -> oracle/_build/libeals_oracle.so, our plain-C restatement (eals_oracle.c).
* ``Reference`` -> oracle/_ref/libeals_ref.so, the reference's own unmodified translation units
                   behind the C shim ref_harness.cpp (``fast=True`` picks the -O3 build that the
                   CPU speed baseline uses).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package (eals_cpp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libeals_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libeals_ref.so")
REF_FAST_SO = os.path.join(HERE, "_ref", "libeals_ref_fast.so")

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(target: str = "all") -> None:
    """Run oracle/Makefile (``ref`` is a no-op on a box without /root/reference)."""
    # make's chatter goes to stderr: bench.py prints exactly one JSON line on stdout
    subprocess.run(["make", "-s", "-C", HERE, target], check=True, stdout=sys.stderr)


def _opt(arr, dtype):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a.ctypes.data_as(C.c_void_p)


def csr_to_csc(M, N, row_ptr, col_idx, val=None):
    """Column-major view with rows ascending inside each column (what SparseMat.cols holds after
    main.cpp:198-205 appends in (u asc, i asc) order).  Stable sort by column keeps u ascending."""
    row_ptr = np.asarray(row_ptr, np.int64)
    col_idx = np.asarray(col_idx, np.int32)
    rows = np.repeat(np.arange(M, dtype=np.int32), np.diff(row_ptr))
    order = np.argsort(col_idx, kind="stable")
    col_ptr = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(col_idx, minlength=N), out=col_ptr[1:])
    cval = None if val is None else np.ascontiguousarray(np.asarray(val, np.float64)[order])
    return col_ptr, np.ascontiguousarray(rows[order]), cval, order.astype(np.int64)


class Port:
    """The C restatement.  Stateless functions over numpy arrays."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build("port")
        L = self.lib = C.CDLL(PORT_SO)
        L.eo_normal_fill.argtypes = [_f64p, C.c_size_t, C.c_double, C.c_double]
        L.eo_item_weights.argtypes = [C.c_int, _i64p, C.c_double, C.c_double, _f64p]
        L.eo_gram_plain.argtypes = [_f64p, C.c_int, C.c_int, _f64p]
        L.eo_gram_weighted.argtypes = [_f64p, _f64p, C.c_int, C.c_int, _f64p]
        sweep = [C.c_int, C.c_int, C.c_int, _i64p, _i32p, C.c_void_p, _f64p, _f64p, _f64p, _f64p,
                 _f64p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.eo_update_user_sweep.argtypes = sweep
        L.eo_update_item_sweep.argtypes = sweep
        L.eo_loss.argtypes = [C.c_int, C.c_int, C.c_int, _i64p, _i32p, C.c_void_p, _f64p, _f64p,
                              _f64p, _f64p, C.c_double]
        L.eo_loss.restype = C.c_double
        L.eo_evaluate.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _i32p, C.c_int, C.c_int,
                                  _f64p, _f64p, _f64p, _i32p, _f64p]

    def normal_fill(self, n, mean=0.0, sigma=0.01):
        out = np.empty(n, np.float64)
        self.lib.eo_normal_fill(out, n, mean, sigma)
        return out

    def item_weights(self, col_ptr, w0, alpha):
        col_ptr = np.ascontiguousarray(col_ptr, np.int64)
        Wi = np.empty(len(col_ptr) - 1, np.float64)
        self.lib.eo_item_weights(len(Wi), col_ptr, w0, alpha, Wi)
        return Wi

    def gram_plain(self, X):
        X = np.ascontiguousarray(X, np.float64)
        S = np.empty((X.shape[1], X.shape[1]))
        self.lib.eo_gram_plain(X, X.shape[0], X.shape[1], S)
        return S

    def gram_weighted(self, X, w):
        X = np.ascontiguousarray(X, np.float64)
        S = np.empty((X.shape[1], X.shape[1]))
        self.lib.eo_gram_weighted(X, np.ascontiguousarray(w, np.float64), X.shape[0], X.shape[1], S)
        return S

    def update_user_sweep(self, row_ptr, col_idx, row_val, U, V, SU, SV, Wi, reg, begin=0, end=None,
                          patch_S=True):
        M, K = U.shape
        self.lib.eo_update_user_sweep(M, V.shape[0], K, row_ptr, col_idx, _opt(row_val, np.float64),
                                      U, V, SU, SV, Wi, reg, begin, M if end is None else end,
                                      int(patch_S))

    def update_item_sweep(self, col_ptr, row_idx, col_val, U, V, SU, SV, Wi, reg, begin=0, end=None,
                          patch_S=True):
        N, K = V.shape
        self.lib.eo_update_item_sweep(U.shape[0], N, K, col_ptr, row_idx, _opt(col_val, np.float64),
                                      U, V, SU, SV, Wi, reg, begin, N if end is None else end,
                                      int(patch_S))

    def loss(self, row_ptr, col_idx, row_val, U, V, SV, Wi, reg):
        return self.lib.eo_loss(U.shape[0], V.shape[0], U.shape[1], row_ptr, col_idx,
                                _opt(row_val, np.float64), U, V, SV, Wi, reg)

    def evaluate(self, U, V, gt_items, topK, compat=True):
        M = U.shape[0]
        hr, ndcg, prec = np.empty(M), np.empty(M), np.empty(M)
        cnt = np.empty(M, np.int32)
        mean = np.empty(3)
        self.lib.eo_evaluate(M, V.shape[0], U.shape[1], U, V, np.ascontiguousarray(gt_items, np.int32),
                             topK, int(compat), hr, ndcg, prec, cnt, mean)
        return mean, hr, ndcg, prec, cnt


class PortModel:
    """A whole eALS model on top of ``Port`` — constructor/buildModel flow of MF_fastALS
    (MF_fastALS.cpp:29-92, 112-161) on flat arrays.  Used as the reference-shaped checker where the
    compiled reference is not available and as the ``port`` CPU baseline."""

    def __init__(self, M, N, row_ptr, col_idx, val=None, factors=64, w0=10.0, alpha=0.75, reg=0.01,
                 init_mean=0.0, init_stdev=0.01, port: Port | None = None):
        self.p = port or Port()
        self.M, self.N, self.K, self.reg = M, N, factors, reg
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        self.col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.val = None if val is None else np.ascontiguousarray(val, np.float64)
        self.col_ptr, self.row_idx, self.cval, _ = csr_to_csc(M, N, self.row_ptr, self.col_idx, self.val)
        self.Wi = self.p.item_weights(self.col_ptr, w0, alpha)
        stream = self.p.normal_fill(max(M, N) * factors, init_mean, init_stdev)
        self.U = stream[: M * factors].reshape(M, factors).copy()   # U and V share one stream
        self.V = stream[: N * factors].reshape(N, factors).copy()   # (DenseMat.cpp:54-62)
        self.init_S()

    def init_S(self):
        self.SU = self.p.gram_plain(self.U)
        self.SV = self.p.gram_weighted(self.V, self.Wi)

    def update_user(self, begin=0, end=None):
        self.p.update_user_sweep(self.row_ptr, self.col_idx, self.val, self.U, self.V, self.SU,
                                 self.SV, self.Wi, self.reg, begin, end)

    def update_item(self, begin=0, end=None):
        self.p.update_item_sweep(self.col_ptr, self.row_idx, self.cval, self.U, self.V, self.SU,
                                 self.SV, self.Wi, self.reg, begin, end)

    def loss(self):
        return self.p.loss(self.row_ptr, self.col_idx, self.val, self.U, self.V, self.SV, self.Wi,
                           self.reg)

    def evaluate(self, gt_items, topK, compat=True):
        return self.p.evaluate(self.U, self.V, gt_items, topK, compat)


class Reference:
    """The real MF_fastALS object from /root/reference behind ref_harness.cpp."""

    def __init__(self, M, N, row_ptr, col_idx, val=None, test_items=None, topK=10, factors=64,
                 maxIter=20, w0=10.0, alpha=0.75, reg=0.01, init_mean=0.0, init_stdev=0.01,
                 fast=False):
        so = REF_FAST_SO if fast else REF_SO
        if not os.path.exists(so):
            build("ref")
        if not os.path.exists(so):
            raise FileNotFoundError(f"{so} missing and /root/reference not available to build it")
        L = self.lib = C.CDLL(so)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int, _i64p, _i32p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int] + [C.c_double] * 5
        for name in ("ref_get_U", "ref_get_V", "ref_get_SU", "ref_get_SV", "ref_get_Wi", "ref_set_Wi"):
            getattr(L, name).argtypes = [C.c_void_p, _f64p]
        L.ref_set_UV.argtypes = [C.c_void_p, _f64p, _f64p]
        for name in ("ref_update_user_sweep", "ref_update_item_sweep", "ref_loss"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_double
        for name in ("ref_update_user_range", "ref_update_item_range"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
            getattr(L, name).restype = C.c_double
        L.ref_update_user_row.argtypes = [C.c_void_p, C.c_int]
        L.ref_update_item_row.argtypes = [C.c_void_p, C.c_int]
        L.ref_predict.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_predict.restype = C.c_double
        L.ref_build_model.argtypes = [C.c_void_p, C.c_int]
        L.ref_evaluate.argtypes = [C.c_void_p, _i32p, C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.ref_dense_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _f64p]
        self.M, self.N, self.K = M, N, factors
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.h = L.ref_create(M, N, row_ptr, col_idx, _opt(val, np.float64),
                              _opt(test_items, np.int32), topK, factors, maxIter, w0, alpha, reg,
                              init_mean, init_stdev)

    def _get(self, name, shape):
        out = np.empty(shape, np.float64)
        getattr(self.lib, name)(self.h, out)
        return out

    U = property(lambda s: s._get("ref_get_U", (s.M, s.K)))
    V = property(lambda s: s._get("ref_get_V", (s.N, s.K)))
    SU = property(lambda s: s._get("ref_get_SU", (s.K, s.K)))
    SV = property(lambda s: s._get("ref_get_SV", (s.K, s.K)))
    Wi = property(lambda s: s._get("ref_get_Wi", (s.N,)))

    def set_UV(self, U, V):
        self.lib.ref_set_UV(self.h, np.ascontiguousarray(U, np.float64), np.ascontiguousarray(V, np.float64))

    def set_Wi(self, Wi):
        self.lib.ref_set_Wi(self.h, np.ascontiguousarray(Wi, np.float64))

    def update_user(self, begin=None, end=None):
        if begin is None:
            return self.lib.ref_update_user_sweep(self.h)
        return self.lib.ref_update_user_range(self.h, begin, end)

    def update_item(self, begin=None, end=None):
        if begin is None:
            return self.lib.ref_update_item_sweep(self.h)
        return self.lib.ref_update_item_range(self.h, begin, end)

    def update_user_row(self, u):
        self.lib.ref_update_user_row(self.h, u)

    def update_item_row(self, i):
        self.lib.ref_update_item_row(self.h, i)

    def loss(self):
        return self.lib.ref_loss(self.h)

    def predict(self, u, i):
        return self.lib.ref_predict(self.h, u, i)

    def build_model(self, iters):
        self.lib.ref_build_model(self.h, iters)

    def evaluate(self, gt_items, topK):
        M = self.M
        hr, ndcg, prec, mean = np.empty(M), np.empty(M), np.empty(M), np.empty(3)
        self.lib.ref_evaluate(self.h, np.ascontiguousarray(gt_items, np.int32), topK, hr, ndcg, prec, mean)
        return mean, hr, ndcg, prec

    def dense_init(self, rows, cols, mean=0.0, sigma=0.01):
        out = np.empty((rows, cols), np.float64)
        self.lib.ref_dense_init(rows, cols, mean, sigma, out)
        return out


def reference_available(fast=False) -> bool:
    so = REF_FAST_SO if fast else REF_SO
    return os.path.exists(so) or os.path.exists("/root/reference/MF_fastALS.cpp")
