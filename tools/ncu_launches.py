"""print a per-launch table from an `ncu --csv --metrics ...` log"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr = r; start = i + 1; break
ix = {h: i for i, h in enumerate(hdr)}
d = OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr): continue
    key = (r[ix['ID']], r[ix['Kernel Name']].replace('srsb200::', '')[:28])
    d.setdefault(key, {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
tot = 0
for k, v in d.items():
    t = float(v.get('gpu__time_duration.sum', '0').replace(',', ''))
    tot += t
    print("%3s %-28s %9.1f us  rd %7.1f MB wr %7.1f MB  inst %6.1f M  issue %5.1f%%  warps %5.1f%%" % (
        k[0], k[1], t / 1e3, float(v.get('dram__bytes_read.sum', '0').replace(',', '')) / 1e6, float(v.get('dram__bytes_write.sum', '0').replace(',', '')) / 1e6,
        float(v.get('smsp__inst_executed.sum', '0').replace(',', '')) / 1e6, float(v.get('smsp__issue_active.avg.pct_of_peak_sustained_active', '0')),
        float(v.get('sm__warps_active.avg.pct_of_peak_sustained_active', '0'))))
print("total %.1f us" % (tot / 1e3))
