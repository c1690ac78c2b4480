"""Development aid: phase time line of the stream compress kernel (library built
with EXTRA=-DB200_CS_TRACE)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr
dr.jit_init()
n = 1 << 28
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
m = (torch.rand(n, device="cuda") < d).to(torch.uint8)
out = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    dr.jit_compress(1, m, n, out)
torch.cuda.synchronize()
buf = np.zeros((320, 16, 12), dtype=np.uint64)
lib = dr.lib()
lib.b200_debug_cs_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
rc = lib.b200_debug_cs_trace(buf.ctypes.data, buf.nbytes)
assert rc == 0, rc
t0 = int(buf[:296, 0, 0].min())
names = ["iter start", "load landed", "packed+arrived", "-", "resolve start", "totals in", "wave read", "resolved", "pfx wait", "pfx ok", "expanded"]
for cta in (0, 1, 147, 148, 295):
    print("CTA", cta)
    for k in range(0, 14):
        row = buf[cta, k]
        print(f"  k={k:2d} " + " ".join(f"{names[e][:10]}={(int(row[e]) - int(t0)) / 1000:7.2f}" if row[e] else f"{names[e][:10]}=   -   " for e in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10)))
for cta in (0, 147, 295):
    print("CTA", cta, "resolve: [totals in -> first loads back -> wave read], repolls")
    for k in range(0, 13):
        row = buf[cta, k].astype(np.int64)
        print(f"  k={k:2d} totals in {(row[5] - t0) / 1000:7.2f}  first loads +{(row[3] - row[5]) / 1000:5.2f}  wave read +{(row[6] - row[3]) / 1000:5.2f}  repolls {row[11]}")
# per-iteration wave statistics
for k in range(0, 14):
    a = buf[:296, k, 2].astype(np.int64) - int(t0)
    r = buf[:296, k, 7].astype(np.int64) - int(t0)
    print(f"wave {k}: packed min/max {a.min() / 1000:.2f}/{a.max() / 1000:.2f} us  resolved min/max {r.min() / 1000:.2f}/{r.max() / 1000:.2f} us")
