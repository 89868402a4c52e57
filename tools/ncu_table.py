#!/usr/bin/env python
"""One line per captured launch of an `ncu --set full` report: the counters DESIGN.md / bench.py quote (run where ncu is installed,
no GPU needed):  tools/ncu_table.py gpurun_out/x.ncu-rep [kernel-name regex]"""
import csv
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__registers_per_thread", "regs", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1),
        ("dram__bytes_read.sum", "dramR", 1), ("dram__bytes_write.sum", "dramW", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%", 1),
        ("smsp__inst_executed.sum", "winst", 1),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 1)]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: one line per captured launch; units as ncu reports them ({', '.join(f'{n}={units[ix[m]]}' for m, n, _ in COLS if m in ix and units[ix[m]])})")
    print("kernel".ljust(44) + " ".join(n.rjust(11) for m, n, _ in COLS if m in ix))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
        if pat and not pat.search(name):
            continue
        vals = []
        for m, n, sc in COLS:
            if m not in ix:
                continue
            try:
                v = float(r[ix[m]].replace(",", ""))
                vals.append(f"{v:11.3f}" if abs(v) < 1e6 else f"{v:11.4g}")
            except ValueError:
                vals.append(r[ix[m]].rjust(11))
        print(name[:43].ljust(44) + " ".join(vals))


if __name__ == "__main__":
    main()
