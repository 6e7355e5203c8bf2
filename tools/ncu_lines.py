#!/usr/bin/env python
"""
Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep captured with `--set full --import-source on`
(the library is built with -lineinfo), read with `ncu -i` (no GPU needed); optionally the line-by-line DIFFERENCE of two reports
(e.g. a 'constant' and a 'reflect' instance of the same kernel), which is how the cost of a feature is located.

    python tools/ncu_lines.py A.ncu-rep [B.ncu-rep] [--top 40] [--out profiles/x.txt]
"""
import argparse
import csv
import subprocess
import sys
from collections import defaultdict


def per_line(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, check=True).stdout
    cur, d = None, {}
    for r in csv.reader(raw.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0].isdigit():
            try:
                inst, samp = int(r[7]), int(r[6])
            except ValueError:
                continue
            a = d.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:110]])
            a[0] += inst
            a[1] += samp
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("rep2", nargs="?")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--out")
    args = ap.parse_args()
    out = open(args.out, "w") if args.out else sys.stdout
    A = per_line(args.rep)
    ta, sa = sum(v[0] for v in A.values()), sum(v[1] for v in A.values())
    if not args.rep2:
        print(f"# {args.rep}: {ta} warp instructions, {sa} stall samples; columns: instructions, share, share of samples, file:line, source", file=out)
        for (f, ln), v in sorted(A.items(), key=lambda kv: -kv[1][0])[: args.top]:
            print(f"{v[0]:>12d} {100 * v[0] / ta:5.1f}% s={100 * v[1] / max(sa, 1):4.1f}%  {f}:{ln}  {v[2]}", file=out)
        return
    B = per_line(args.rep2)
    tb = sum(v[0] for v in B.values())
    print(f"# A = {args.rep}: {ta} warp instructions;  B = {args.rep2}: {tb};  B - A per line", file=out)
    rows = []
    for k in set(A) | set(B):
        a, b = A.get(k, [0, 0, ""])[0], B.get(k, [0, 0, ""])[0]
        rows.append((b - a, k, a, b, (B.get(k) or A.get(k))[2]))
    rows.sort(reverse=True)
    for dlt, (f, ln), a, b, src in rows[: args.top]:
        print(f"{dlt:>+12d}  {a:>10d} -> {b:>10d}  {f}:{ln}  {src}", file=out)
    byfile = defaultdict(int)
    for dlt, (f, _), *_ in rows:
        byfile[f] += dlt
    print("# per file:", dict(byfile), file=out)


if __name__ == "__main__":
    main()
