#!/usr/bin/env python
"""Developer tool: SASS of one kernel with the ncu source-page counters beside it, for the instructions whose source
line falls in a range of obca_cta.cuh (e.g. the Riccati sweep).

    python tools/ncu_sass_region.py /tmp/src.csv <objdir> <kernel-substring> <first line> <last line> [file]
"""
import csv, os, re, subprocess, sys, tempfile


def disasm(objdir, kname):
    for ob in sorted(f for f in os.listdir(objdir) if f.startswith("obca_kv_") and f.endswith(".o")):
        tmp = tempfile.mkdtemp()
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(objdir, ob)], cwd=tmp, stdout=subprocess.DEVNULL)
        cubs = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        if not cubs:
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubs[0])], capture_output=True, text=True).stdout
        if any(l.startswith(".text.") and kname in l for l in out.splitlines()):
            return out.splitlines()
    return []


def main():
    src_csv, objdir, kname, lo, hi = sys.argv[1:6]
    objdir = os.path.abspath(objdir)
    fname = sys.argv[6] if len(sys.argv) > 6 else "obca_cta.cuh"
    lo, hi = int(lo), int(hi)
    lines = []; cur = None; infn = False
    for l in disasm(objdir, kname):
        if l.startswith(".text."):
            infn = kname in l; continue
        if not infn: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): lines.append(cur)
    rows = list(csv.reader(open(src_csv))); hdr = rows[1]; data = rows[2:]; ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = 0
    for n, (ln, r) in enumerate(zip(lines, data)):
        if not ln or ln[0] != fname or not (lo <= ln[1] <= hi):
            continue
        samp = float(r[ix["# Samples"]] or 0); ex = float(r[ix["Instructions Executed"]] or 0)
        tot += samp
        top = sorted(((float(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print("%6d %5d %-70s ex %10.0f samp %7.0f  %s" % (n, ln[1], r[ix["Source"]].strip()[:70], ex, samp,
              " ".join("%s=%.0f" % (k, v) for v, k in top if v > 0)))
    print("total samples in range:", tot)


if __name__ == "__main__":
    main()
