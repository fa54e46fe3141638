"""Top source lines by warp-stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
    python profiles/top_lines.py report.ncu-rep [N]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = None, None, {}
stall_cols = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0].strip().isdigit():  # a CUDA-C line row (its SASS rows follow with empty Line No)
        key = (cur_file, int(r[0]), r[1].strip())
        try:
            n = int(r[si])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, {}])
        a[0] += n
        for i in stall_cols:
            try:
                v = int(r[i])
            except (ValueError, IndexError):
                v = 0
            if v:
                a[1][hdr[i]] = a[1].get(hdr[i], 0) + v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for key, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{n:8d} {100 * n / max(tot, 1):5.1f}%  {key[0]}:{key[1]:<4d} {key[2][:90]:90s} {top}")
