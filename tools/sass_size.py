#!/usr/bin/env python
"""Code size of a kernel by CUDA source line (needs -lineinfo): SASS instructions attributed to each line, summed per source
region.  The contact step is instruction-cache sensitive (sm__icc_request_hit_rate 83 %, four CTAs per SM in different phases
of an 8 k-instruction kernel), so code size is a performance number here.
Usage: tools/sass_size.py path/to/lib.so [kernel-substring] [bucket]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so = os.path.abspath(sys.argv[1])
kern = sys.argv[2] if len(sys.argv) > 2 else "k_simulate"
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
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
        total = 0
        for l in dis[start:end]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                cnt[cur] += 1
                total += 1
        print(f"{kern}: {total} SASS instructions = {total * 16 / 1024:.0f} KB")
        agg = collections.Counter()
        for (f, ln), v in cnt.items():
            agg[(f, ln // bucket * bucket)] += v
        for (f, b), v in sorted(agg.items()):
            print(f"  {f}:{b:4d}-{b + bucket - 1:4d}  {v:6d}")
        break
