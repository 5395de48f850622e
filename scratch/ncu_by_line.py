"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line, using nvdisasm --print-line-info on the cubin.
usage: ncu_by_line.py <source.csv> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# find the .text section of the kernel
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and kname in l)
off2line, cur = {}, None
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
opc = collections.Counter()
for r in rows[2:]:
    if len(r) <= isamp:
        continue
    off = int(r[ia], 16) - base
    ln, ins = off2line.get(off, (None, "?"))
    s, e = int(r[isamp] or 0), int(r[iexec] or 0)
    by[ln][0] += s; by[ln][1] += e
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            by[ln][2][hdr[i]] += v
    tot_s += s; tot_e += e
    opc[r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]] += e
print(f"total samples {tot_s}, warp instructions executed {tot_e}")
lines = {}
for (k, v) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{a[6:]}:{b}" for a, b in v[2].most_common(3))
    print(f"{100*v[0]/tot_s:5.1f}% samp {100*v[1]/tot_e:5.1f}% inst  {k}  [{st}]")
print("opcode mix (executed warp instructions):")
for k, v in opc.most_common(25):
    print(f"  {k:14s} {100*v/tot_e:5.1f}%")
# ---- optional: aggregate by line ranges given as name:lo-hi,... in env NCU_REGIONS (file psvi_mf_fn1.cu / others by name)
import os
reg = os.environ.get("NCU_REGIONS")
if reg:
    regs = []
    for it in reg.split(","):
        n, r = it.split(":"); lo, hi = r.split("-"); regs.append((n, int(lo), int(hi)))
    agg = collections.defaultdict(lambda: [0, 0])
    main = os.environ.get("NCU_MAIN", "psvi_mf_fn1.cu")
    for k, v in by.items():
        name = "other"
        if k is None:
            name = "noline"
        elif k[0] == main:
            for n, lo, hi in regs:
                if lo <= k[1] <= hi:
                    name = n; break
        else:
            name = k[0]
        agg[name][0] += v[0]; agg[name][1] += v[1]
    print("regions:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:28s} {100*v[0]/tot_s:5.1f}% samples  {100*v[1]/tot_e:5.1f}% instructions")
