#!/bin/bash
# Second gpurun call of the next round (after tools/round2_first_call.sh is green): launch list + one full ncu
# capture each of the bucketed feature kernel and the march kernel, summarised into gpurun_out/ for profiles/.
#   gpurun --timeout 1200 -- 'bash tools/round2_profile.sh'
mkdir -p gpurun_out
CUT=${CUT:-5.96e-8}
echo "== launch list (bucketed feature kernel)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bucket.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --feature-cutoff $CUT > gpurun_out/r2_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_bucket.csv")) if len(r) > 10]
hdr = rows[0]; k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        t[r[k][:60]][0] += 1; t[r[k][:60]][1] += float(r[v].replace(",", ""))
    except ValueError:
        pass
for name, (n, ns) in sorted(t.items(), key=lambda x: -x[1][1])[:15]:
    print(f"{ns/1e3:10.1f} us  x{n:4d}  {name}")
PY
for K in sam_bucket_kernel march_kernel; do
  echo "== ncu --set full: $K"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -f -o gpurun_out/r2_$K \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --feature-cutoff $CUT > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r2_$K.ncu-rep > gpurun_out/r2_${K}_ncu.txt 2>&1; head -30 gpurun_out/r2_${K}_ncu.txt
done
