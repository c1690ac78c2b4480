"""Development aid: times jit_block_prefix_reduce(f32, Add) for the geometry
selected by B200_SCAN_GEOM (tuning build only: make EXTRA=-DB200_SCAN_TUNING)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch
    import drjit_core_b200 as dr
    dr.jit_init()
    n = 1 << 28
    x = torch.rand(n, device="cuda", dtype=torch.float32)
    out = torch.empty(n, device="cuda", dtype=torch.float32)
    res = []
    for bs in (2, 128, 4096, 1 << 16, n):
        fn = lambda: dr.jit_block_prefix_reduce(1, 14, 1, n, bs, 1, 0, x, out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        import ctypes
        dbg = (ctypes.c_ulonglong * 16)()
        have_dbg = hasattr(dr.lib(), "b200_scan_debug")
        if have_dbg:
            dr.lib().b200_scan_debug(dbg, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        res.append(f"bs={bs if bs < n else 'N'}: {ms:.3f} ms {8 * n / ms / 1e6:.0f} GB/s")
        if have_dbg:
            dr.lib().b200_scan_debug(dbg, 1)
            d = list(dbg)
            if d[5]:
                t = d[5]
                res.append(f"[tiles {t // 10} steps/tile {d[0] / t:.2f} polls/step {d[1] / max(d[0], 1):.2f} "
                           f"lb_cyc {d[2] / t:.0f} lb_wait_agg {d[7] / t:.0f} agg_wait_land {d[6] / t:.0f} "
                           f"agg_cyc {d[8] / t:.0f} cmp_wait_pref {d[3] / t:.0f} cmp_wait_land {d[4] / t:.0f} "
                           f"issue_cyc {d[9] / t:.0f} t0_wait_agg {d[11] / t:.0f} loop_cyc/tile {d[10] / t:.0f}]")
    print(f"geom {os.environ.get('B200_SCAN_GEOM', '0')}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
    else:
        for g in (sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6", "7"]):
            geom, _, dbg = g.partition(":")
            env = dict(os.environ, B200_SCAN_GEOM=geom, B200_DEBUG="1", B200_SCAN_DBG=dbg or "0")
            print(f"--- geom {geom} dbg {dbg or 0}", flush=True)
            subprocess.run([sys.executable, __file__, "worker"], env=env)
