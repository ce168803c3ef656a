#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the default bench line, the reference arm; then the
# launch list of the same bench command (ncu gpu__time_duration, all kernels) for profiles/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real
echo "bench exit $?"; cut -c1-400 gpurun_out/final_bench.json
( time timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/final_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/final_ncu.log 2>&1
echo "launch list exit $?"
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/final_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
# the two timed epochs = the last 2/5 of the library's launches (3 warm-up + 2 timed)
lib = [r for r in data if "eals::" in r[ix["Kernel Name"]]]
tail = lib[len(lib) * 3 // 5:]
agg = collections.OrderedDict()
for r in tail:
    k = r[ix["Kernel Name"]].split("(")[0][-46:]
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u.startswith("u") else v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("two timed epochs, library kernels: %.1f ms under ncu" % tot)
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1]/2:9.2f} ms/epoch {a[1]/tot*100:5.1f}%  x{a[0]//2:4d}  {k}")
PY
