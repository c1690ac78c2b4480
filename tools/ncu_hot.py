"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output
(development aid).  usage: ncu_hot.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h = heads[sec]
end = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
hdr = rows[h]
col = {name: i for i, name in enumerate(hdr)}
data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
print("sections:", len(heads), "showing", sec)
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
print("instructions executed", sum(int(r[col["Instructions Executed"]] or 0) for r in data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[col['# Samples']]):7d} {100.0 * int(r[col['# Samples']]) / tot:5.1f}%  exec {r[col['Instructions Executed']]:>9s}  {r[col['Source']][:70]:70s} {top}")
