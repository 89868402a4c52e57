#!/usr/bin/env python
"""Source-line view of an `ncu --set full --import-source on` report: stall samples per CUDA source line (inlined helpers are
listed under their own file), the block barriers ranked by the samples spent waiting behind them, and the overall stall mix.
Usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, top_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    f, hdr, cur, lines, after_bar, stalls = None, None, None, [], [], {}
    prev_bar = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] != "":
            cur = (f, int(r[0]), r[1].strip())
            try:
                lines.append((int(r[6]), int(r[7]), f, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
            continue
        if r[2] in ("...", "-"):
            continue
        try:
            samp = int(r[6])
        except ValueError:
            continue
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit():
                stalls[h] = stalls.get(h, 0) + int(r[i])
        if prev_bar is not None:          # a warp waiting at a barrier is sampled at the instruction AFTER the BAR
            after_bar.append((samp, prev_bar))
            prev_bar = None
        if "BAR.SYNC" in r[3]:
            prev_bar = cur
    tot = sum(l[0] for l in lines) or 1
    print(f"total samples {tot}")
    print("== stall mix (all samples)")
    ts = sum(stalls.values()) or 1
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {h[6:]:18s} {100 * v / ts:5.1f}%")
    print("== block barriers by samples spent waiting behind them")
    for s, (ff, ln, src) in sorted(after_bar, reverse=True)[:12]:
        print(f"  {100 * s / tot:5.1f}%  {ff}:{ln}: {src[:90]}")
    print(f"== top {top_n} source lines by samples")
    for s, i, ff, ln, src in sorted(lines, reverse=True)[:top_n]:
        print(f"  {100 * s / tot:5.1f}%  inst {i:>10}  {ff}:{ln}: {src}")


if __name__ == "__main__":
    main()
