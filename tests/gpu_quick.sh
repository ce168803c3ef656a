#!/bin/bash
# quick guarded check of a new kernel build: smoke() under a 90 s timeout, then a few parity tests under 240 s
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke exit $?"
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
