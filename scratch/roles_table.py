"""Side-by-side table of two ncu --csv metric captures of the tn_gemm launches of one cfg5 step (roles_cg0.csv / roles_cg1.csv)."""
import csv, re, collections, sys
def load(fn):
    rd = csv.reader(l for l in open(fn) if l.startswith('"'))
    hdr = next(rd); rows = collections.OrderedDict()
    for r in rd:
        d = dict(zip(hdr, r)); k = int(d['ID'])
        rows.setdefault(k, {'name': re.sub(r'.*tn_gemm_kernel', 'tn', d['Kernel Name'])[:8]})
        rows[k][d['Metric Name']] = (float(d['Metric Value'].replace(',', '')), d['Metric Unit'])
    return rows
def fmt(r):
    t = r['gpu__time_duration.sum']; tt = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[t[1]] * t[0]
    ta = r['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'][0]
    l2 = r['lts__t_sectors_srcunit_tex_op_read.sum'][0] * 32 / 1e6
    return f"{r['name']:8s} {tt:7.1f}us ta {ta:5.1f}% L2rd {l2:6.0f}MB", tt
a, b = load(sys.argv[1]), load(sys.argv[2])
sa = sb = 0
for k in list(a.keys())[:int(sys.argv[3]) if len(sys.argv) > 3 else 21]:
    fa, ta = fmt(a[k]); fb, tb = fmt(b[k]); sa += ta; sb += tb
    print(k, fa, ' | ', fb)
print('sum', round(sa, 1), round(sb, 1))
