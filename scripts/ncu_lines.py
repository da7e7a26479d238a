#!/usr/bin/env python
"""Attribute the per-instruction counters of an `ncu --page source --csv` dump (SASS view) to CUDA source lines, using the line
table of the SAME binary (nvdisasm -g of the cubin extracted with cuobjdump -xelf).  Instructions are matched by order.
Usage: ncu_lines.py <src.csv.gz> <disassembly.dis> <kernel mangled-name substring> [launch index] [top N]"""
import csv, gzip, io, re, sys, collections


def dis_lines(dis, kern):
    out = []; cur = None; on = False; stack = []
    for ln in open(dis, errors="replace"):
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    src, dis, kern = sys.argv[1:4]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    rows = list(csv.reader(io.StringIO(gzip.open(src, "rt").read())))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    s = starts[launch]; e = starts[launch + 1] - 1 if launch + 1 < len(starts) else len(rows)
    hdr = rows[s]; body = [r for r in rows[s + 1:e] if r and r[0].startswith("0x")]
    lines = dis_lines(dis, kern)
    print(f"launch {launch}: {len(body)} instructions in the profile, {len(lines)} in the disassembly")
    n = min(len(body), len(lines))
    ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); ct = hdr.index("Thread Instructions Executed")
    acc = collections.defaultdict(lambda: [0, 0, 0])
    for k in range(n):
        key = lines[k][:2] if lines[k] else ("?", 0)
        a = acc[key]; a[0] += int(body[k][cs] or 0); a[1] += int(body[k][ci] or 0); a[2] += int(body[k][ct] or 0)
    tot = [sum(a[j] for a in acc.values()) for j in range(3)]
    print(f"totals: samples {tot[0]}, warp instructions {tot[1]}, thread instructions {tot[2]} (avg lanes {tot[2] / max(tot[1], 1):.1f})")
    print("  samples%  inst%  lanes  file:line")
    for key, a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100 * a[0] / max(tot[0], 1):6.1f}  {100 * a[1] / max(tot[1], 1):6.1f}  {a[2] / max(a[1], 1):5.1f}  {key[0]}:{key[1]}")


if __name__ == "__main__":
    main()
