"""Hot source lines of one kernel: joins the per-SASS-instruction counters of an .ncu-rep with the line table of
the cubin (nvdisasm -g; needs -lineinfo).  python tools/ncu_source.py report.ncu-rep <kernel substring> [launch index] [N]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "drmlt-mitsuba_b200", "csrc", "libdrmlt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = {}      # mangled function -> {offset: (file, line)}
for cub in glob.glob(os.path.join(tmp, "*sm_100a.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    fn, cur = None, None
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1); lines[fn] = {}; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m and fn:
            lines[fn][int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
launches, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1].replace("void ", "").replace("(int)", "").replace("(bool)", ""), "rows": []}; launches.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
sel = [l for l in launches if kern in l["name"] and l["rows"]]
L = sel[which]
fn = [f for f in lines if kern.split("<")[0] in f]
exact = [f for f in fn if ("%d%s7Machine" % (len(kern), kern)) in f]      # k_trace(Machine), not k_trace_rays(...)
fn = exact[0] if exact else fn[0] if len(fn) == 1 else [f for f in fn if (("ILi%s" % kern.split("<")[1].rstrip(">")) in f or ("ILb%s" % kern.split("<")[1].rstrip(">")) in f)][0] if "<" in kern else fn[0]
table = lines[fn]
hdr = L["hdr"]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(L["rows"][0][ia], 16)
agg = {}
for r in L["rows"]:
    off = int(r[ia], 16) - base
    key = table.get(off) or ("?", 0)
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += float(r[ii] or 0); a[1] += float(r[it] or 0); a[2] += float(r[isamp] or 0)
tot, tots = sum(a[0] for a in agg.values()) or 1, sum(a[2] for a in agg.values()) or 1
print("%s (%s): %.3g warp-instructions, %d samples, thr/inst %.1f" % (L["name"], fn, tot, tots, sum(a[1] for a in agg.values()) / tot))
src = {}
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    f, ln = key
    if f not in src:
        p = os.path.join(ROOT, "drmlt-mitsuba_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print("%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d  %s" % (100 * a[2] / tots, 100 * a[0] / tot, a[1] / a[0] if a[0] else 0, f, ln, text))
