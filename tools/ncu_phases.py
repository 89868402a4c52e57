#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report of k_simulate: key raw metrics + instruction / stall-sample
share of every barrier-delimited code region (SASS order).  Usage: tools/ncu_phases.py report.ncu-rep [kernel-index]"""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, rows = raw(rep)
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fma.sum"]
    for r in rows:
        print("== launch")
        for h, v in zip(hdr, r):
            if h in want:
                print(f"  {h} = {v}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # split per kernel
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            kern.append(cur)
        elif cur is not None:
            cur.append(r)
    k = kern[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
    hdr = k[0]
    ci = {h: i for i, h in enumerate(hdr)}
    ie, ns, src, te = ci["Instructions Executed"], ci["# Samples"], ci["Source"], ci["Thread Instructions Executed"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    segs, cur, tot = [], dict(n=0, inst=0, samp=0, thr=0, st={}), 0
    for r in k[1:]:
        try:
            n, s, t = int(r[ie]), int(r[ns]), int(r[te])
        except Exception:
            continue
        tot += n
        cur["n"] += 1; cur["inst"] += n; cur["samp"] += s; cur["thr"] += t
        for h in stall_cols:
            try:
                cur["st"][h] = cur["st"].get(h, 0) + int(r[ci[h]])
            except Exception:
                pass
        if "BAR.SYNC" in r[src] or "EXIT" in r[src]:
            segs.append(cur)
            cur = dict(n=0, inst=0, samp=0, thr=0, st={})
    segs.append(cur)
    tsamp = sum(s["samp"] for s in segs) or 1
    print(f"== regions between block barriers (SASS order); total warp-instructions {tot}")
    for i, s in enumerate(segs):
        if s["inst"] * 200 < tot:
            continue
        top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:3]
        print(f"  region {i:3d}: {s['n']:5d} SASS  inst {100 * s['inst'] / tot:5.1f}%  samples {100 * s['samp'] / tsamp:5.1f}%  "
              f"threads/inst {s['thr'] / max(s['inst'], 1):4.1f}  top stalls {[(k[6:], round(100 * v / max(s['samp'], 1))) for k, v in top]}")


if __name__ == "__main__":
    main()
