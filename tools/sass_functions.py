#!/usr/bin/env python
"""Per-kernel SASS fingerprints of a built library, to prove that an edit left the GPU-verified kernels untouched.

    python tools/sass_functions.py segment-anything-in-nerf_b200/libsnrf.so > /tmp/a.json     # before
    python tools/sass_functions.py segment-anything-in-nerf_b200/libsnrf.so --diff /tmp/a.json  # after

Prints {demangled kernel name: [instruction count, md5 of the instruction stream]}; --diff lists kernels whose stream
changed, appeared or disappeared.  Addresses and encodings are stripped, so only real code changes show up."""
import hashlib
import json
import re
import subprocess
import sys


def fingerprints(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    fps, cur, i = {}, None, -1
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            i += 1
            cur = names[i] if i < len(names) and names[i] else m.group(1)
            fps[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            fps[cur].append(m.group(1).strip())
    return {k: [len(v), hashlib.md5("\n".join(v).encode()).hexdigest()] for k, v in fps.items()}


def main():
    lib = sys.argv[1]
    fp = fingerprints(lib)
    if "--diff" in sys.argv:
        base = json.load(open(sys.argv[sys.argv.index("--diff") + 1]))
        changed = [k for k in base if k in fp and fp[k] != base[k]]
        gone = [k for k in base if k not in fp]
        new = [k for k in fp if k not in base]
        for tag, ks in (("CHANGED", changed), ("REMOVED", gone), ("NEW", new)):
            for k in ks:
                print(tag, k[:150], base.get(k), "->", fp.get(k))
        print(f"{len(base) - len(changed) - len(gone)} of {len(base)} baseline kernels identical; {len(new)} new")
        sys.exit(1 if changed or gone else 0)
    json.dump(fp, sys.stdout, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
