"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,dram bytes] --csv): per kernel and grid size, for the LAST solve of the run.
python tools/launch_agg.py gpurun_out/x.csv [vcycles]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mn = h.index('Metric Name'); mv = h.index('Metric Value'); idc = h.index('ID'); gs = h.index('Grid Size'); bs = h.index('Block Size')
data = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    d = data.setdefault(r[idc], {'name': r[kn].split('(')[0][:70], 'grid': r[gs], 'block': r[bs]})
    d[r[mn]] = float(r[mv].replace(',', ''))
L = list(data.values())
idx = [i for i, d in enumerate(L) if 'k_make_rhs' in d['name']]
seg = L[idx[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in seg:
    a = agg[(d['name'], d['grid'], d['block'])]; a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0); a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{k[0][5:47]:42s} {k[1]:18s} {k[2]:12s} n={a[0]:4d} per-vcycle {a[1]/nv/1e3:8.1f} us  avg {a[1]/a[0]/1e3:8.1f} us  {a[2]/a[0]/1e6:9.1f} MB/launch {a[2]/max(a[1],1):7.2f} TB/s" )
