"""profiles/rNN_traffic.json from `ncu --set full` reports of bench.py's hot kernels (one report per kernel).

    python tools/ncu_traffic.py profiles/r02_traffic.json march=gpurun_out/x_march.ncu-rep feature=... tapgemm=...

Records dram__bytes_read.sum / dram__bytes_write.sum of the captured launch, the rays that launch processed (from the
kernel's grid: bench.py prints nothing under ncu, so the caller passes --rays) and the digest of the CUDA sources the
reports were captured from (bench.py compares it with the sources it is running)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return {h: v for h, v in zip(rows[0], rows[2])}, {h: u for h, u in zip(rows[0], rows[1])}


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    from bench import source_digest

    out_path, rays = sys.argv[1], None
    res = {"source_digest": source_digest(), "how": "ncu --set full --clock-control none, one launch of bench.py's workload (cold caches)"}
    for arg in sys.argv[2:]:
        if arg.startswith("--rays="):
            rays = int(arg.split("=")[1])
            continue
        k, rep = arg.split("=")
        d, u = raw(rep)
        res[k] = {"dram_bytes_read": to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
                  "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
                  "rays_per_launch": rays, "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")),
                  "kernel": d.get("Kernel Name", "?"), "report": os.path.basename(rep)}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
