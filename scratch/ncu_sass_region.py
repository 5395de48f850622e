"""Print SASS instructions (address order) whose source line falls in [lo, hi] of a file, with executed counts / samples.
usage: ncu_sass_region.py <source.csv> <cubin> <kernel-substring> <file> <lo> <hi>"""
import csv, re, subprocess, sys
src_csv, cubin, kname, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and kname in l)
off2line, cur = {}, None
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
n = 0
for r in rows[2:]:
    off = int(r[ia], 16) - base
    ln = off2line.get(off)
    if ln and ln[0] == fname and lo <= ln[1] <= hi:
        st = ",".join(f"{hdr[i][6:]}:{r[i]}" for i in stall_cols if int(r[i] or 0))
        print(f"{off:6x} L{ln[1]:4d} ex={r[iexec]:>6s} s={r[isamp]:>3s} {r[1].strip()[:70]:70s} {st}")
        n += 1
print(n, "instructions")
