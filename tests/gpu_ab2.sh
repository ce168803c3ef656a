#!/bin/bash
# usage: gpu_ab2.sh <workload> "<tag>:<ENV=V ...>;..."   then: launch list (ncu gpu__time_duration) + e2e stage timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=$1; CFGS=$2
bash tests/gpu_ab.sh $W "$CFGS" || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$W.csv \
  python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_$W.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
# second epoch only: skip launches up to the middle
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg = collections.OrderedDict()
for r in data[len(data) // 2:]:
    k = r[ix["Kernel Name"]][:70]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print(f"{a[1]:9.2f} ms {a[1]/tot*100:5.1f}% x{a[0]:5d}  {k}")
PY
EALS_VERBOSE=1 timeout 600 python bench.py --workload $W --steps 1 --warmup 1 --no-cpu --e2e-steps 1 2>&1 | grep "\[eals\]" | tail -40
