#!/bin/bash
# Round-2 final profile call (1 GPU): suite, bench lines of the three configs + reference arm, launch list, one
# `ncu --set full` capture per hot kernel, and profiles/r02_traffic.json tied to the digest of the sources that ran.
mkdir -p gpurun_out
T=f1
echo "== GPU tests"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
echo "== bench (default)"; timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_share_ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'path', round(d['roofline']['path']['frac'],3), 'cpu', d['cpu_baseline']['value'])"
echo "== bench rgb"; timeout 300 python bench.py --config rgb --no-cpu-baseline > gpurun_out/${T}_bench_rgb.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_rgb.json')); print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
echo "== bench clipseg_patch"; timeout 300 python bench.py --config clipseg_patch --no-cpu-baseline > gpurun_out/${T}_bench_clipseg.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_clipseg.json')); print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_ref.json')); print(d['value'], d['cpu_baseline']['cores'])"
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${T}_launches.csv")) if len(r) > 10]
hdr = rows[0]; k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        t[r[k][:70]][0] += 1; t[r[k][:70]][1] += float(r[v].replace(",", ""))
    except ValueError:
        pass
tot = sum(x[1] for x in t.values())
with open("gpurun_out/${T}_launch_summary.txt", "w") as f:
    for name, (n, ns) in sorted(t.items(), key=lambda x: -x[1][1])[:15]:
        line = f"{ns/1e3:10.1f} us  {100*ns/tot:5.1f} %  x{n:4d}  {name}"
        print(line); f.write(line + "\n")
PY
for K in march_kernel sam_bucket_kernel tapgemm_tma_kernel; do
  echo "== ncu --set full: $K"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o gpurun_out/${T}_$K \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/${T}_$K.ncu-rep > gpurun_out/${T}_${K}_ncu.txt 2>&1; head -8 gpurun_out/${T}_${K}_ncu.txt
  python tools/ncu_opmix.py gpurun_out/${T}_$K.ncu-rep 30 >> gpurun_out/${T}_${K}_ncu.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${T}_$K.ncu-rep 30 >> gpurun_out/${T}_${K}_ncu.txt 2>&1
done
python tools/ncu_traffic.py gpurun_out/r02_traffic.json --rays=131072 march=gpurun_out/${T}_march_kernel.ncu-rep \
  feature=gpurun_out/${T}_sam_bucket_kernel.ncu-rep tapgemm=gpurun_out/${T}_tapgemm_tma_kernel.ncu-rep | head -3
