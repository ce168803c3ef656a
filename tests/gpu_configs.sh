#!/bin/bash
# The other BASELINE.json configurations (parity-test cases, not the bench line): one bench line each for
# c1/c2/c3 (+ the reference arm on c1: the real compiled reference, 1 thread) and evaluate() at c3 and c5 (= c4 shapes, top-100).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for W in c1 c2 c3; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/cfg_$W.json 2> gpurun_out/cfg_$W.err
  echo "bench $W exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cfg_$W.json").read().strip().split("\n")[-1])
    print("  $W ms/step %.3f"%d["ms_per_step"], "value %.3e"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"], "cpu", d["cpu_baseline"] and ("%.3e (%s, %d cores)"%(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])))
except Exception as e: print("parse failed", e)
PY
done
timeout 900 python bench.py --workload c1 --impl reference --steps 2 --warmup 1 > gpurun_out/cfg_c1_ref.json 2> gpurun_out/cfg_c1_ref.err; echo "ref c1 exit $?"; cut -c1-600 gpurun_out/cfg_c1_ref.json
timeout 600 python tests/eval_probe.py c3 3 2>&1 | grep evaluate
timeout 900 python tests/eval_probe.py c4 2 2>&1 | grep evaluate
