#!/usr/bin/env python
"""Where ptxas put k_simulate's register spills: STL / LDL instructions per CUDA source line (needs -lineinfo).
Usage: tools/spill_lines.py path/to/libseqdex_b200.so [kernel-substring]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so = os.path.abspath(sys.argv[1])
kern = sys.argv[2] if len(sys.argv) > 2 else "k_simulate"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, check=True, capture_output=True)
    for cub in sorted(os.listdir(d)):
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
        starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l]
        if not starts:
            continue
        start = starts[0]
        end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.")), len(dis))
        cur, cnt = None, collections.Counter()
        for l in dis[start:end]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif re.search(r"\b(STL|LDL)\b", l):
                cnt[cur] += 1
        for k, v in sorted(cnt.items()):
            print(f"{k[0]}:{k[1]}  {v}")
        break
