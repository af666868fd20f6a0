#!/usr/bin/env python
"""Developer tool: join an ncu source-page CSV (per SASS instruction) with nvdisasm line info and aggregate executed
instructions / stall samples per CUDA source line and per function-sized region.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_by_line.py /tmp/src.csv <lib.so> <kernel-substring> [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    src_csv, lib, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    # instruction index -> source line, for the requested kernel
    lines = []
    cur = None; infn = False
    for l in dis:
        if l.startswith(".text."):
            infn = kname in l
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
            lines.append(cur)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]; data = rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}
    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    assert abs(len(lines) - len(data)) < 8, (len(lines), len(data))
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0])
    for ln, r in zip(lines, data):
        a = agg[ln]
        a[0] += f(r, "Instructions Executed"); a[1] += f(r, "# Samples"); a[2] += f(r, "stall_no_inst"); a[3] += f(r, "stall_barrier"); a[4] += 1
    ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print("total warp instructions %.3e, samples %d, static instructions %d" % (ti, ts, len(data)))
    print("%-22s %8s %8s %8s %8s %7s" % ("line", "exec%", "samp%", "noinst%", "barrier%", "static"))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-22s %8.2f %8.2f %8.2f %8.2f %7d" % ("%s:%d" % ln if ln else "?", 100 * a[0] / ti, 100 * a[1] / ts, 100 * a[2] / ts, 100 * a[3] / ts, a[4]))


if __name__ == "__main__":
    main()
