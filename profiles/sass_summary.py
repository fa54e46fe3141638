"""Per-kernel counts of the Blackwell-only SASS mnemonics in libgenpose_b200.so (cuobjdump -sass):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk = TMA bulk
copy), SYNCS (mbarrier), UCGABAR (cluster barrier).  Writes profiles/sass_summary.txt.

    python profiles/sass_summary.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "genpose2_b200", "libgenpose_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "UCGABAR", "REDUX", "FFMA", "HFMA2", "DFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    lines = ["SASS summary of genpose2_b200/libgenpose_b200.so (cuobjdump -sass; regenerate with python profiles/sass_summary.py)",
             "arch: " + ", ".join(sorted(set(arch))), "",
             "%-110s %8s " % ("kernel", "instrs") + " ".join("%8s" % k for k in MNEMONICS)]
    tot = collections.Counter()
    for fn in sorted(order, key=lambda f: -counts[f]["_total"]):
        c = counts[fn]
        name = demangle(fn)
        name = re.sub(r"\(.*\)$", "", name)[:110]
        lines.append("%-110s %8d " % (name, c["_total"]) + " ".join("%8d" % c[k] for k in MNEMONICS))
        tot.update(c)
    lines.append("%-110s %8d " % ("TOTAL (%d kernels)" % len(order), tot["_total"]) + " ".join("%8d" % tot[k] for k in MNEMONICS))
    out = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
    print("...")
    print(lines[-1])


if __name__ == "__main__":
    main()
