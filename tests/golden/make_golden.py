"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference): drives the reference's own
MF_fastALS object, compiled by oracle/Makefile into oracle/_ref/libeals_ref.so, and stores its
outputs.  The fixtures travel to the GPU box (where /root/reference does not exist) and pin both the
C restatement (oracle/eals_oracle.c, CPU tests) and the CUDA path (GPU tests).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from eals_cpp_b200 import datasets          # noqa: E402
from oracle.bindings import Reference       # noqa: E402


def run_case(name, data, K, iters, topK, scale_eval=None):
    ref = Reference(data.M, data.N, data.row_ptr, data.col_idx, test_items=data.test_items, topK=topK,
                    factors=K, maxIter=iters)
    out = {"M": data.M, "N": data.N, "K": K, "topK": topK, "iters": iters}
    arrays = {"row_ptr": data.row_ptr, "col_idx": data.col_idx, "test_items": data.test_items,
              "Wi": ref.Wi, "U0_head": ref.U[:4].copy(), "SU0": ref.SU, "SV0": ref.SV}
    losses = [ref.loss()]
    for it in range(iters):
        ref.update_user()
        if it == 0:
            arrays["U_after_first_user_sweep"] = ref.U
        ref.update_item()
        if it == 0:
            arrays["V_after_first_item_sweep"] = ref.V
        losses.append(ref.loss())
    out["losses"] = [float(x) for x in losses]
    arrays["losses"] = np.asarray(losses)
    arrays["U_final"], arrays["V_final"] = ref.U, ref.V
    arrays["SU_final"], arrays["SV_final"] = ref.SU, ref.SV
    mean, hr, ndcg, prec = ref.evaluate(data.test_items, topK)
    arrays["eval_mean"], arrays["eval_hr"], arrays["eval_ndcg"], arrays["eval_prec"] = mean, hr, ndcg, prec
    out["eval_mean"] = mean.tolist()
    if scale_eval:
        # inflate the factors so that int-truncated scores are non-zero: exercises the ranking bug
        rng = np.random.default_rng(99)
        U = ref.U * scale_eval + rng.normal(0, 0.5, (data.M, K))
        V = ref.V * scale_eval + rng.normal(0, 0.5, (data.N, K))
        ref.set_UV(U, V)
        mean2, hr2, ndcg2, prec2 = ref.evaluate(data.test_items, topK)
        arrays.update(U_scaled=U, V_scaled=V, eval2_mean=mean2, eval2_hr=hr2, eval2_ndcg=ndcg2, eval2_prec=prec2)
        out["eval2_mean"] = mean2.tolist()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **arrays)
    return out


def main():
    summary = {}
    tiny = datasets.powerlaw_csr(**datasets.WORKLOADS["tiny"])
    summary["tiny_k8"] = run_case("tiny_k8", tiny, K=8, iters=5, topK=10, scale_eval=30.0)
    summary["tiny_k64"] = run_case("tiny_k64", tiny, K=64, iters=3, topK=10)
    # DenseMat::init known answers (default-seeded minstd_rand0 + polar normal)
    ref = Reference(4, 4, np.array([0, 1, 2, 3, 4], np.int64), np.array([0, 1, 2, 3], np.int32), factors=2)
    init = ref.dense_init(64, 8, 0.0, 0.01)
    np.savez_compressed(os.path.join(HERE, "dense_init.npz"), init=init)
    summary["dense_init_first8"] = [float(x) for x in init.ravel()[:8]]
    with open(os.path.join(HERE, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1)[:1500])


if __name__ == "__main__":
    main()
