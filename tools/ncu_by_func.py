#!/usr/bin/env python
"""Developer tool: like ncu_by_line.py but aggregated per function of obca_cta.cuh (by source-line ranges).

    python tools/ncu_by_func.py /tmp/src.csv <lib.so> <kernel-substring>
"""
import bisect, collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.environ.get("OBCA_HDR") or os.path.join(ROOT, "vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200", "csrc", "obca_cta.cuh")


def main():
    src_csv, lib, kname = sys.argv[1:4]
    funcs = []
    for n, l in enumerate(open(HDR), 1):
        if re.match(r"\s*(OB_HD|template|struct)\b", l) and ("(" in l or l.startswith("struct")):
            funcs.append((n, l.strip()[:58]))
    starts = [f[0] for f in funcs]
    # one cubin per kernel variant (one translation unit each, all called obca_variant - extracted from the library they
    # would overwrite each other, so they are taken from the objects of the build): take the one that holds the kernel
    objdir = os.path.join(os.path.dirname(os.path.abspath(lib)), "_obj_prof" if "prof" in os.path.basename(lib) else "_obj")
    dis = []
    for ob in sorted(f for f in os.listdir(objdir) if f.startswith("obca_kv_") and f.endswith(".o")):
        tmp = tempfile.mkdtemp()
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(objdir, ob)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubs = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        if not cubs:
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubs[0])], capture_output=True, text=True).stdout
        if any(l.startswith(".text.") and kname in l for l in out.splitlines()):
            dis = out.splitlines()
            break
    lines = []; cur = None; infn = False
    for l in dis:
        if l.startswith(".text."):
            infn = kname in l; continue
        if not infn: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): lines.append(cur)
    rows = list(csv.reader(open(src_csv))); hdr = rows[1]; data = rows[2:]; ix = {k: i for i, k in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]])
        except Exception: return 0.0
    agg = collections.defaultdict(lambda: [0.0] * 7)
    for ln, r in zip(lines, data):
        if ln and ln[0] == "obca_cta.cuh":
            i = bisect.bisect_right(starts, ln[1]) - 1
            key = funcs[i][1] if i >= 0 else "?"
        else:
            key = "%s:%s" % (ln[0], ln[1]) if ln and ln[0] == "obca_b200.cu" else (str(ln[0]) if ln else "?")
        a = agg[key]
        a[0] += f(r, "Instructions Executed"); a[1] += f(r, "# Samples"); a[2] += f(r, "stall_no_inst"); a[3] += f(r, "stall_barrier")
        a[4] += 1; a[5] += f(r, "stall_short_sb") + f(r, "stall_long_sb"); a[6] += f(r, "stall_wait")
    ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print("%-60s %6s %6s %7s %6s %6s %6s %6s" % ("function", "exec%", "samp%", "noinst%", "barr%", "sb%", "wait%", "static"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%-60s %6.2f %6.2f %7.2f %6.2f %6.2f %6.2f %6d" % (k, 100 * a[0] / ti, 100 * a[1] / ts, 100 * a[2] / ts, 100 * a[3] / ts, 100 * a[5] / ts, 100 * a[6] / ts, a[4]))


if __name__ == "__main__":
    main()
