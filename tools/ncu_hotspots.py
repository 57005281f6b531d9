"""Warp-stall sample summary of an ncu report (source page, SASS view): totals per stall reason, totals per opcode class, and the 25
instructions with the most samples.  usage: python tools/ncu_hotspots.py rep.ncu-rep [out.txt]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, universal_newlines=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(lines[start:]))
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    tot = defaultdict(float)
    per_op = defaultdict(float)
    total = 0.0
    for r in rows:
        try:
            n = float(r["# Samples"])
        except (ValueError, KeyError):
            continue
        total += n
        op = r["Source"].split()[0] if r["Source"].split() else "?"
        if op.startswith("@"):
            op = r["Source"].split()[1]
        per_op[op.split(".")[0]] += n
        for c in stall_cols:
            try:
                tot[c] += float(r[c])
            except ValueError:
                pass
    text = ["%s" % lines[0][:160], "total warp-stall samples: %d" % total, "", "by stall reason:"]
    for c, v in sorted(tot.items(), key=lambda x: -x[1])[:12]:
        text.append("  %-26s %9d  %5.1f %%" % (c, v, 100.0 * v / max(total, 1)))
    text += ["", "by opcode:"]
    for c, v in sorted(per_op.items(), key=lambda x: -x[1])[:14]:
        text.append("  %-14s %9d  %5.1f %%" % (c, v, 100.0 * v / max(total, 1)))
    text += ["", "top instructions:"]
    for r in sorted(rows, key=lambda r: -float(r["# Samples"] or 0))[:25]:
        n = float(r["# Samples"] or 0)
        top = max(stall_cols, key=lambda c: float(r[c] or 0))
        text.append("  %7d  %5.2f %%  %-60s  %s" % (n, 100.0 * n / max(total, 1), r["Source"].strip()[:60], top))
    s = "\n".join(text) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s)
    else:
        sys.stdout.write(s)


if __name__ == "__main__":
    main()
