#!/bin/bash
# Full default bench (what the driver runs), the reference arm, then DRAM traffic of every sweep kernel of one
# epoch (ncu, two metrics only) -> gpurun_out/traffic_launches.csv
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=${1:-c4}
( time timeout 1200 python bench.py --workload $W > gpurun_out/full_$W.json 2> gpurun_out/full_$W.err ) 2>&1 | grep real
echo "bench exit $?"; tail -3 gpurun_out/full_$W.err; cat gpurun_out/full_$W.json | cut -c1-3000
( time timeout 1200 python bench.py --workload $W --impl reference --steps 2 --warmup 1 > gpurun_out/ref_$W.json 2> gpurun_out/ref_$W.err ) 2>&1 | grep real
echo "reference exit $?"; cat gpurun_out/ref_$W.json | cut -c1-1500
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'cd_|heavy_' -c 6000 --csv --log-file gpurun_out/traffic_launches.csv \
  python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1
echo "traffic capture exit $?"
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/traffic_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ids = sorted({int(r[ix["ID"]]) for r in data})
half = ids[len(ids) // 2]
agg = collections.OrderedDict()
def scale(v, u):
    u = u.lower()
    if u.startswith("g"): return v * 1e9
    if u.startswith("m") and "byte" in u: return v * 1e6
    if u.startswith("k"): return v * 1e3
    return v
for r in data:
    if int(r[ix["ID"]]) < half: continue
    k = r[ix["Kernel Name"]].split("(")[0][-48:]
    m = r[ix["Metric Name"]]; v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    a = agg.setdefault(k, {"n": 0, "rd": 0.0, "wr": 0.0, "ms": 0.0})
    if m.startswith("dram__bytes_read"): a["rd"] += scale(v, u); a["n"] += 1
    elif m.startswith("dram__bytes_write"): a["wr"] += scale(v, u)
    else: a["ms"] += v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u.startswith("u") else v)
tot = sum(a["rd"] + a["wr"] for a in agg.values())
print("second epoch: DRAM bytes of the sweep kernels = %.1f GB" % (tot / 1e9))
for k, a in agg.items():
    print(f"{a['ms']:9.2f} ms  rd {a['rd']/1e9:8.2f} GB  wr {a['wr']/1e9:7.2f} GB  x{a['n']:5d}  {k}")
PY
