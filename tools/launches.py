"""Per-launch table from an `ncu --csv --metrics ...` launch list (development aid)."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki]), {})[r[mi]] = r[vi]
f = lambda m, k: float(m.get(k, '0').replace(',', ''))
for (i, k), m in d.items():
    if 'at::' in k and len(sys.argv) < 3:
        continue
    k = re.sub(r'\(.*', '', k)[:64]
    print(f"{i:>4} {k:64s} {f(m, 'gpu__time_duration.sum') / 1000:9.1f}us rd {f(m, 'dram__bytes_read.sum') / 1e6:8.1f} "
          f"wr {f(m, 'dram__bytes_write.sum') / 1e6:8.1f} inst {f(m, 'smsp__inst_executed.sum') / 1e6:7.1f}M "
          f"issue {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', '')}")
