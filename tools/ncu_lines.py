"""Per-source-line executed warp-instructions and stall samples from an ncu report captured with --import-source on."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, items, seen_fn = None, None, [], set()
kern = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kern = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0] not in ("", "Line No") and len(r) >= 8:
        try:
            n = int(r[7])
            st = int(r[4])
        except ValueError:
            continue
        key = (kern, fname, r[0])
        if key in seen_fn:
            continue
        seen_fn.add(key)
        items.append((n, st, fname, int(r[0]), r[1].strip()[:110], kern))
first = items[0][5] if items else None
items = [i for i in items if i[5] == first]
tot = sum(i[0] for i in items)
tst = sum(i[1] for i in items)
print(f"kernel: {first}\ntotal warp-instructions {tot}, stall samples {tst}")
for n, st, f, ln, src, _ in sorted(items, reverse=True)[:top]:
    print(f"{100*n/tot:6.2f}% inst {100*st/max(tst,1):6.2f}% stall  {f}:{ln:<4d} {src}")
