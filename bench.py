#!/usr/bin/env python
"""bench.py -- throughput of the data-parallel primitive path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], extended by the whole-array case of
configs[0]): one STEP is one pass of

    jit_block_reduce(f32, Add, n, bs)  and
    jit_block_prefix_reduce(f32, Add, n, bs, exclusive, forward)

for bs in {1, 2, 4, ..., 4096, n} over n = 2^28 fp32 per GPU (28 primitive
calls, 28 * n elements per step and GPU).  Inputs are synthetic
(x[i] = (fmix32(i) >> 8) * 2^-24), 1 GiB per array, i.e.
larger than the 126 MB L2, so no explicit L2 flush is needed between
iterations.

With --gpus N > 1 (launched under torchrun, one rank per GPU) every rank owns
a contiguous 2^28 shard of a global array of N * 2^28 elements (weak scaling).
Block sizes <= 4096 divide the shard, so those calls need no exchange; the
whole-array reduce / scan (bs = global size) go through the sharded front end
(drjit-core_b200/sharded.py: local pass + NCCL all_gather of one scalar per
rank + local single-pass scan seeded with the rank's carry).

Printed keys (one JSON line, rank 0): the contract of the build driver plus
  roofline      dominant kernel (scan_stream_kernel<float, Add>): algorithmic bytes
                per launch / mean launch duration from CUDA events recorded
                inside the timed region, against MEASURED_PEAKS.json
  e2e           same step through the C-ABI with HOST (pinned) input and output
                buffers: H2D of the input and D2H of every result inside the
                timed region (PCIe bound)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, nanothread
                pool on all host cores) on the same step, N = 1 only
  primitives    per-primitive figures for the other BASELINE configs (reduce /
                prefix sum u32, compress, mkperm, scatter-add, dot)
  sharded       configs[4]: 2^32-element fp32 reduce + exclusive scan sharded
                over the N GPUs (strong scaling, reported separately)

--impl reference times the reference's CPU implementation of the same step
(oracle/_ref when present, else the single-threaded oracle port) and prints the
same line with "impl": "reference".  This is the only place besides
cpu_baseline where bench.py executes anything under oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LOG2N_DEFAULT = 28
BLOCK_SIZES = [1 << k for k in range(13)]  # 1 .. 4096
F32, U32, ADD, CUDA = 14, 8, 1, 1
METRIC = "elements/s (block_reduce + block_prefix_reduce, fp32 Add, 2^28 per GPU, bs 1..4096 and N)"


def step_calls(n_global):
    """The 28 primitive calls of one step as (kind, block_size)."""
    calls = []
    for bs in BLOCK_SIZES + [n_global]:
        calls.append(("reduce", bs))
        calls.append(("scan", bs))
    return calls


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/roofline_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# --------------------------------------------------------------------------- CPU arm

def np_fmix32(i):
    h = (np.asarray(i, dtype=np.uint32) + np.uint32(1)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85ebca6b)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xc2b2ae35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def np_input_f32(n, start=0):
    out = np.empty(n, dtype=np.float32)
    chunk = 1 << 24
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = np.arange(start + s, start + e, dtype=np.uint64).astype(np.uint32)
        out[s:e] = (np_fmix32(i) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return out


class CpuArm:
    """The reference's CPU implementation of the step (checker code: oracle/)."""

    def __init__(self):
        import oracle
        if oracle.ref_available():
            self.impl = oracle.Reference()
            self.kind = "reference"
            self.cores = self.impl.threads
        else:
            self.impl = oracle.Oracle()
            self.kind = "port"
            self.cores = 1

    def run_step(self, x, out, calls):
        n = x.shape[0]
        for kind, bs in calls:
            bs = min(bs, n)
            if self.kind == "reference":
                if kind == "reduce":
                    self.impl.block_reduce(F32, ADD, x, bs, out=out)
                else:
                    self.impl.block_prefix_reduce(F32, ADD, x, bs, 1, 0, out=out)
            else:
                if kind == "reduce":
                    self.impl.block_reduce(F32, ADD, x, bs)
                else:
                    self.impl.block_prefix_reduce(F32, ADD, x, bs, 1, 0)

    def measure(self, log2n, steps, warmup, budget_s):
        """Returns (elements/s, ms/step, description of the sample)."""
        log2 = log2n if self.kind == "reference" else min(log2n, 22)
        while True:
            n = 1 << log2
            x = np_input_f32(n)
            out = np.zeros(n, dtype=np.float32)
            calls = step_calls(n)
            t0 = time.perf_counter()
            self.run_step(x, out, calls)  # first-touch + warm-up
            t_first = time.perf_counter() - t0
            if t_first * (steps + warmup) <= budget_s or log2 <= 20:
                break
            log2 -= 2
        for _ in range(max(0, warmup - 1)):
            self.run_step(x, out, calls)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.run_step(x, out, calls)
        dt = (time.perf_counter() - t0) / steps
        sample = (f"{steps} timed step(s) after {max(1, warmup)} warm-up of the same 28-call step at "
                  f"n=2^{log2} fp32 per call, wall clock incl. jit_sync_thread")
        return len(calls) * n / dt, dt * 1e3, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    value, ms, sample = arm.measure(args.log2n, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.log2n, args.gpus),
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": arm.cores,
                         "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(log2n, world):
    return {
        "workload": (f"BASELINE configs[1]+[0]: jit_block_reduce + jit_block_prefix_reduce(exclusive) "
                     f"fp32 Add over 2^{log2n} elements per GPU, block sizes 1,2,4,..,4096 and the whole "
                     f"array (28 calls per step)"),
        "elements_per_gpu": 1 << log2n, "calls_per_step": 2 * (len(BLOCK_SIZES) + 1),
        "global_elements": world << log2n,
        "parallelism": f"shard{world}" if world > 1 else "single",
        "l2": "inputs (1 GiB per array) larger than the 126 MB L2; no flush needed",
    }


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and clock-event reasons of one GPU with NVML while the
    timed region runs."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
    BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}

    def __init__(self, torch_device_index):
        self.samples, self.reasons, self.ok = [], set(), False
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = torch_device_index
                if vis:
                    try:
                        idx = int(vis.split(",")[torch_device_index])
                    except Exception:
                        pass
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.h = h
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        self.samples, self.reasons = [], set()
        self._stop.clear()
        if self.ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [],
                    "note": "NVML unavailable" if not self.ok else "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}

    def rejected(self):
        return bool(self.reasons & self.BAD)


# --------------------------------------------------------------------------- GPU arm

def torch_fmix32(torch, i):
    M = 0xFFFFFFFF
    h = (i + 1) & M
    h = h ^ (h >> 16)
    h = (h * 0x85ebca6b) & M
    h = h ^ (h >> 13)
    h = (h * 0xc2b2ae35) & M
    h = h ^ (h >> 16)
    return h


def fill_input(torch, out_f32=None, out_u32=None, start=0, xor=0, mod=None):
    """Counter-based synthetic data generated on the device in chunks."""
    t = out_f32 if out_f32 is not None else out_u32
    n = t.numel()
    chunk = 1 << 26
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = torch.arange(start + s, start + e, dtype=torch.int64, device=t.device)
        if xor:
            i = i ^ xor
        h = torch_fmix32(torch, i & 0xFFFFFFFF)
        if out_f32 is not None:
            out_f32[s:e] = (h >> 8).to(torch.float32) * (2.0 ** -24)
        else:
            if mod:
                h = h % mod
            out_u32[s:e] = h.to(torch.int32) if mod else (h - ((h >> 31) << 32)).to(torch.int32)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import drjit_core_b200 as dr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dr.jit_init()

    from importlib import import_module
    sharded_mod = import_module("drjit_core_b200.sharded")
    sh = sharded_mod.Sharded(device=dev, exchange=args.exchange)

    n = 1 << args.log2n
    n_global = n * world
    calls = step_calls(n_global)
    x = torch.empty(n, dtype=torch.float32, device=dev)
    fill_input(torch, out_f32=x, start=rank * n)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    scalar = torch.zeros(4, dtype=torch.float32, device=dev)

    def call(kind, bs, src, dst):
        if bs >= n_global:
            if world > 1:
                if kind == "reduce":
                    sh.reduce(F32, ADD, src, n, dst)
                else:
                    sh.prefix_reduce(F32, ADD, src, n, True, False, dst)
                return
            bs = n
        if kind == "reduce":
            dr.jit_block_reduce(CUDA, F32, ADD, n, bs, src, dst)
        else:
            dr.jit_block_prefix_reduce(CUDA, F32, ADD, n, bs, 1, 0, src, dst)

    def out_elems(kind, bs):
        if kind == "scan":
            return n
        return 1 if bs >= n_global else (n + bs - 1) // bs

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    K, W = args.steps, args.warmup

    def timed_region():
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in calls] for _ in range(K)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        launches0 = dr.launch_count()
        sampler.start()
        t0.record()
        for k in range(K):
            for c, (kind, bs) in enumerate(calls):
                ev[k][c][0].record()
                call(kind, bs, x, out)
                ev[k][c][1].record()
        t1.record()
        torch.cuda.synchronize()
        sampler.stop()
        barrier()
        launches = dr.launch_count() - launches0
        total_ms = t0.elapsed_time(t1)
        per_call = np.array([[a.elapsed_time(b) for a, b in row] for row in ev])  # K x calls
        return total_ms, per_call, launches

    for _ in range(W):
        for kind, bs in calls:
            call(kind, bs, x, out)
    total_ms, per_call, launches = timed_region()
    remeasured = False
    if sampler.rejected():
        remeasured = True
        total_ms, per_call, launches = timed_region()
    clocks = sampler.summary()
    if remeasured:
        clocks["remeasured"] = True

    # sanity: the timed kernels did the work (sum of n uniform [0,1) values ~ n/2)
    call("reduce", n_global, x, scalar)
    torch.cuda.synchronize()
    total = float(scalar[0].item())
    if not (0.49 * n_global < total < 0.51 * n_global):
        raise SystemExit(f"bench.py: whole-array reduce returned {total}, expected ~{n_global / 2}")

    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_max.item())
    ms_per_step = total_ms_max / K
    value = world * len(calls) * n / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: scan_stream_kernel<float, Add> (1 < bs <= 4096)
    peak, peak_kind = load_peaks()
    mean_call = per_call.mean(axis=0)
    scan_idx = [c for c, (kind, bs) in enumerate(calls) if kind == "scan" and bs > 1 and bs < n_global]
    scan_ms = float(mean_call[scan_idx].mean())
    scan_share = float(mean_call[scan_idx].sum() / mean_call.sum())
    achieved = 8.0 * n / (scan_ms * 1e-3) / 1e9
    traffic = load_traffic()
    roofline = {
        "bound": "hbm", "kernel": "scan_stream_kernel<float, Add, 4, 3, 512, CHAIN=false> (jit_block_prefix_reduce, bs 2..4096)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
        "frac_of_nominal_8000": achieved / 8000.0,
        "algorithmic_bytes_per_launch": 8 * n, "avg_launch_ms": scan_ms,
        "share_of_step": scan_share,
        "traffic": traffic.get("scan_kernel_bytes_per_launch") if traffic else None,
    }
    per_call_report = {
        f"{kind}_bs{'N' if bs >= n_global else bs}": {
            "ms": float(mean_call[c]),
            "GBs": (4.0 * n * (1 + 1 / bs) if kind == "reduce" else (8.0 * n if bs > 1 else 4.0 * n))
            / (float(mean_call[c]) * 1e-3) / 1e9}
        for c, (kind, bs) in enumerate(calls)}
    for v in per_call_report.values():
        v["frac"] = v["GBs"] / peak

    # ---- end to end: host buffers, copies inside the timed region
    e2e = measure_e2e(torch, dr, dist, world, dev, n, n_global, calls, call, out_elems, K,
                      max(1, min(W, 2)), barrier)

    line = {
        "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.log2n, world),
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        "hbm_GBs_step": sum((4.0 * n * (1 + 1 / min(bs, n)) if kind == "reduce"
                             else (8.0 * n if bs > 1 else 4.0 * n)) for kind, bs in calls)
        / (ms_per_step * 1e-3) / 1e9,
        "calls": per_call_report,
    }
    line["hbm_frac_step"] = line["hbm_GBs_step"] / peak

    if not args.no_primitives:
        del out
        torch.cuda.empty_cache()
        if world == 1:
            prim = measure_primitives(torch, dr, dev, peak)
            line["primitives"] = prim
            step2 = prim.pop("_step_compress_mkperm")
            line["value_all_primitives"] = {
                "value": (len(calls) * n + 3 * (1 << 28) + 3 * (1 << 26)) / ((ms_per_step + step2["ms"]) * 1e-3),
                "unit": "elements/s",
                "note": "the headline step (28 reduce / scan calls) plus the compress / mkperm step, "
                        "elements of both divided by the sum of both times",
                "step_compress_mkperm": step2}
            # worst primitive at the BASELINE sizes (the north_star bar is 0.8 for every one)
            fr = {k: v["frac"] for k, v in prim.items() if "frac" in v}
            fr.update({f"step:{k}": v["frac"] for k, v in per_call_report.items()})
            worst = min(fr, key=fr.get)
            line["roofline_min"] = {"name": worst, "frac": fr[worst],
                                    "by_family": {fam: min(v for k, v in fr.items() if fam in k)
                                                  for fam in ("reduce", "scan", "prefix_sum", "compress",
                                                              "mkperm", "scatter_add", "scatter_inc")
                                                  if any(fam in k for k in fr)}}
        if not args.no_sharded:
            line["sharded"] = measure_sharded(torch, dr, dist, sh, world, rank, dev, peak, barrier)

    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        del x
        torch.cuda.empty_cache()
        try:
            arm = CpuArm()
            v, ms, sample = arm.measure(args.log2n, steps=2, warmup=1, budget_s=25.0)
            line["cpu_baseline"] = {"value": v, "unit": "elements/s", "cores": arm.cores,
                                    "kind": arm.kind, "sample": sample, "ms_per_step": ms}
        except Exception as e:  # the checker is optional for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0,
                                    "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(torch, dr, dist, world, dev, n, n_global, calls, call, out_elems, K, W, barrier):
    """The same step with HOST input/output buffers: per step one H2D of the
    input and one D2H of every call's result (double-buffered on a copy stream so
    that PCIe transfers overlap the kernels).  All copies go through the C-ABI
    (jit_memcpy_async)."""
    h_x = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_x.copy_(torch.rand(n))  # contents do not matter for the timing
    h_out = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    d_x = torch.empty(n, dtype=torch.float32, device=dev)
    d_out = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
    main = torch.cuda.current_stream()
    copy = torch.cuda.Stream(device=dev)
    h2d = 4 * n
    d2h = sum(4 * out_elems(kind, bs) for kind, bs in calls)

    free = [None, None]  # result buffer b may be overwritten once its D2H copy has finished

    def step():
        # (the H2D of this step's input is ordered behind the previous step's kernels by the main stream; it
        # overlaps the tail of the previous step's D2H traffic -- the two directions of the link are independent)
        dr.jit_memcpy_async(CUDA, d_x, h_x, 4 * n, stream=main)
        for c, (kind, bs) in enumerate(calls):
            b = c & 1
            if free[b] is not None:
                main.wait_event(free[b])
            call(kind, bs, d_x, d_out[b])
            done = torch.cuda.Event()
            done.record(main)
            copy.wait_event(done)
            dr.jit_memcpy_async(CUDA, h_out[b], d_out[b], 4 * out_elems(kind, bs), stream=copy)
            free[b] = torch.cuda.Event()
            free[b].record(copy)

    for _ in range(W):
        step()
    main.wait_stream(copy)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(K):
        step()
    main.wait_stream(copy)  # every result of the K steps has reached the host
    t1.record()
    torch.cuda.synchronize()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / K
    return {"value": world * len(calls) * n / (ms_step * 1e-3), "unit": "elements/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_step,
            "pcie_GBs": (h2d + d2h) / (ms_step * 1e-3) / 1e9,
            "host": host_topology(),
            "note": "pinned host input -> H2D -> 28 C-ABI calls -> D2H of every result "
                    "(copy stream overlapped with the kernels, the next step's H2D with the tail of the D2H traffic); PCIe bound.  pcie_GBs is PER RANK; at N > 1 "
                    "the ranks share the host's memory system (the 8-GPU boxes of this pool expose ONE NUMA "
                    "node with 32 cores to all GPUs -- profiles/r2_topo_n8.txt -- so there is nothing to bind "
                    "a rank to: the aggregate of ~96 GB/s is the host's limit)"}


def host_topology():
    """CPUs and NUMA nodes the process can see (evidence for the e2e figure at N > 1)."""
    nodes = 0
    try:
        nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        pass
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count()
    return {"cpus": cpus, "numa_nodes": nodes}


def time_call(torch, fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def measure_primitives(torch, dr, dev, peak):
    """The other BASELINE configs on one GPU (median of 5 after 2 warm-ups)."""
    res = {}

    def rec(name, ms, elems, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "elements_per_s": elems / (ms * 1e-3), "GBs": gbs,
                     "frac": gbs / peak, "frac_of_nominal_8000": gbs / 8000.0}

    n = 1 << 28
    xi = torch.empty(n, dtype=torch.int32, device=dev)
    fill_input(torch, out_u32=xi)
    oi = torch.empty(n, dtype=torch.int32, device=dev)
    rec("reduce_u32_2^28", time_call(torch, lambda: dr.jit_reduce(CUDA, U32, ADD, xi, n, oi)), n, 4 * n)
    rec("prefix_sum_u32_excl_2^28",
        time_call(torch, lambda: dr.jit_block_prefix_reduce(CUDA, U32, ADD, n, n, 1, 0, xi, oi)), n, 8 * n)
    xf, of = xi.view(torch.float32), oi.view(torch.float32)
    rec("reduce_dot_f32_2^28", time_call(torch, lambda: dr.jit_reduce_dot(CUDA, F32, xf, of, n, oi)), n, 8 * n)

    # compress (C3): mask[i] = fmix32(i ^ 0x9E3779B9) < d * 2^32
    h = torch.empty(n, dtype=torch.int32, device=dev)
    fill_input(torch, out_u32=h, xor=0x9E3779B9)
    masks = {}
    for d in (0.01, 0.5, 0.99):
        thr = int(d * 2 ** 32)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        chunk = 1 << 26
        for s in range(0, n, chunk):
            hu = h[s:s + chunk].to(torch.int64) & 0xFFFFFFFF
            mask[s:s + chunk] = (hu < thr).to(torch.uint8)
        cnt = dr.jit_compress(CUDA, mask, n, oi)
        ms = time_call(torch, lambda: dr.jit_compress(CUDA, mask, n, oi))
        rec(f"compress_2^28_d{d}", ms, n, n + 4 * cnt)
        res[f"compress_2^28_d{d}"]["count"] = int(cnt)
        masks[d] = (mask, int(cnt))
    del h

    # mkperm (C4): key[i] = fmix32(i) % B, n = 2^26, one group, offsets requested
    n2 = 1 << 26
    keys = torch.empty(n2, dtype=torch.int32, device=dev)
    perm = oi[:n2]
    mk = {}
    for B in (16, 1024, 65536):
        kB = torch.empty(n2, dtype=torch.int32, device=dev)
        fill_input(torch, out_u32=kB, mod=B)
        offs = torch.zeros(4 * B + 1, dtype=torch.int32).pin_memory()
        ms = time_call(torch, lambda: dr.jit_block_mkperm(CUDA, kB, n2, n2, B, perm, offs))
        rec(f"mkperm_2^26_B{B}", ms, n2, 8 * n2)
        mk[B] = (kB, offs)

    # second timed step: the OTHER primitives of BASELINE.json's metric back to back
    # (configs[2] + configs[3]: compress x3 densities at 2^28, mkperm x3 bucket counts at 2^26)
    def step2():
        for d, (mask, _) in masks.items():
            dr.jit_compress(CUDA, mask, n, oi)
        for B, (kB, offs) in mk.items():
            dr.jit_block_mkperm(CUDA, kB, n2, n2, B, perm, offs)

    ms2 = time_call(torch, step2, iters=5, warmup=3)
    elems2 = 3 * n + 3 * n2
    bytes2 = sum(n + 4 * c for _, c in masks.values()) + 3 * 8 * n2
    res["_step_compress_mkperm"] = {"ms": ms2, "elements_per_s": elems2 / (ms2 * 1e-3),
                                    "GBs": bytes2 / (ms2 * 1e-3) / 1e9,
                                    "frac": bytes2 / (ms2 * 1e-3) / 1e9 / peak,
                                    "calls": "jit_compress 2^28 x {0.01, 0.5, 0.99} + jit_block_mkperm 2^26 x "
                                             "{16, 1024, 65536} buckets (with offsets), synchronous calls"}
    del masks, mk

    # scatter-add (C5): 2^26 -> 2^20, random and coherent indices
    m = 1 << 20
    val = xf[:n2]
    tgt = torch.zeros(m, dtype=torch.float32, device=dev)
    fill_input(torch, out_u32=keys, mod=m)
    for mode, mname in ((0, "auto"), (1, "direct")):
        ms = time_call(torch, lambda: dr.scatter_reduce(F32, ADD, tgt, val, keys, None, n2, mode=mode))
        rec(f"scatter_add_f32_2^26_to_2^20_random_{mname}", ms, n2, 8 * n2)
    keys.copy_((torch.arange(n2, dtype=torch.int64, device=dev) >> 6).to(torch.int32))
    ms = time_call(torch, lambda: dr.scatter_reduce(F32, ADD, tgt, val, keys, None, n2, mode=0))
    rec("scatter_add_f32_2^26_to_2^20_coherent_auto", ms, n2, 8 * n2)

    # scatter_inc (jit_var_scatter_inc): one shared counter (queue compaction)
    cnt = torch.zeros(m, dtype=torch.int32, device=dev)
    keys.zero_()
    ms = time_call(torch, lambda: dr.scatter_inc(cnt, keys, None, perm, n2))
    rec("scatter_inc_2^26_one_counter", ms, n2, 8 * n2)
    # reducing packet scatter: 2^24 four-component splats into 2^20 x 4 floats
    n3 = 1 << 24
    fill_input(torch, out_u32=keys, mod=m)
    comps = [xf[k * n3:(k + 1) * n3] for k in range(4)]
    tgt4 = torch.zeros(4 * m, dtype=torch.float32, device=dev)
    ms = time_call(torch, lambda: dr.scatter_reduce_packet(F32, ADD, tgt4, comps, keys, None, n3, mode=0))
    rec("scatter_add_packet_f32x4_2^24_to_2^20_random_auto", ms, n3, 20 * n3)
    return res


def load_sharded_basis():
    """N = 1 times of the sharded workload from the committed 1-GPU run
    (profiles/sharded_n1_basis.json), so that an N > 1 line carries its own
    strong-scaling figures."""
    try:
        with open(os.path.join(ROOT, "profiles", "sharded_n1_basis.json")) as f:
            return json.load(f)
    except Exception:
        return None


def timed_collective(torch, dist, dev, world, barrier, fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts))


def measure_sharded(torch, dr, dist, sh, world, rank, dev, peak, barrier):
    """BASELINE configs[4]: 2^32 fp32 elements in total, contiguous shards of
    2^32 / N per GPU; whole-array reduce and exclusive scan, plus the mkperm histogram
    of 2^26 keys per rank.  At N = 1 the array exceeds the uint32 sizes of the reference
    API: it is processed as chunks of 2^31 chained through the carry entry point of the
    C-ABI.  `parity` = the sharded results were checked on the device against an
    independent torch computation (u32 bit-exact, fp32 within 1e-5)."""
    total = 1 << 32
    n_local = total // world
    x = torch.empty(n_local, dtype=torch.float32, device=dev)
    fill_input(torch, out_f32=x, start=rank * n_local)
    out = torch.empty(n_local, dtype=torch.float32, device=dev)
    res_buf = torch.zeros(4, dtype=torch.float32, device=dev)
    carry = torch.zeros(4, dtype=torch.float32, device=dev)
    chunk = 1 << 31

    def do_reduce():
        if world == 1:
            dr.jit_reduce(CUDA, F32, ADD, x, n_local, res_buf)
        else:
            sh.reduce(F32, ADD, x, n_local, res_buf)

    def do_scan():
        if world == 1:
            for ci, s in enumerate(range(0, n_local, chunk)):
                e = min(n_local, s + chunk)
                dr.prefix_reduce_carry(F32, ADD, e - s, True, False, x[s:e], out[s:e],
                                       carry if ci else None, carry)
        else:
            sh.prefix_reduce(F32, ADD, x, n_local, True, False, out)

    res = {"total_elements": total, "elements_per_gpu": n_local, "scaling": "strong",
           "exchange": "peer mailboxes over NVLink (csrc/sharded.cu)" if sh.peer is not None
           else ("nccl all_gather" if world > 1 else "single GPU")}
    for name, fn, bpe in (("reduce", do_reduce, 4), ("exclusive_scan", do_scan, 8)):
        ms = timed_collective(torch, dist, dev, world, barrier, fn)
        gbs = bpe * total / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "elements_per_s": total / (ms * 1e-3), "GBs_aggregate": gbs,
                     "frac_of_n_gpu_peak": gbs / (peak * world)}
    torch.cuda.synchronize()
    res["reduce_value"] = float(res_buf[0].item())

    # ---- opt-in block-cyclic layout (global block b of 2^21 elements on rank b % N): ONE
    # pass over the data (8 B / element / GPU instead of 12), block totals through
    # peer-mapped tables.  Same x, read as this rank's blocks; checked on the device below.
    cyc_block = 1 << 21
    if world > 1 and sh.peer is not None:
        out_c = out  # (the contiguous result is checked further down: use a second buffer)
        out_c = torch.empty(n_local, dtype=torch.float32, device=dev)

        def do_scan_cyclic():
            sh.prefix_reduce_cyclic(F32, ADD, x, n_local, cyc_block, True, out_c)

        # (block size: large enough to amortise a block's exchange, small enough that the tiles in
        # flight per GPU -- 888 x 32 KiB -- span several blocks; all three are timed, the best is kept)
        by_block = {}
        for cb in (1 << 20, 1 << 21, 1 << 22):
            cyc_block = cb
            by_block[cb] = timed_collective(torch, dist, dev, world, barrier, do_scan_cyclic)
        cyc_block = min(by_block, key=by_block.get)
        ms = by_block[cyc_block]
        do_scan_cyclic()  # (the result that is checked below)
        torch.cuda.synchronize()
        gbs = 8 * total / (ms * 1e-3) / 1e9
        res["exclusive_scan_block_cyclic"] = {
            "ms": ms, "elements_per_s": total / (ms * 1e-3), "GBs_aggregate": gbs,
            "frac_of_n_gpu_peak": gbs / (peak * world), "block_elements": cyc_block,
            "ms_by_block_elements": {str(k): v for k, v in by_block.items()},
            "layout": "global block b lives on rank b % N as local block b // N (opt-in)"}
        # parity: totals of every local block -> global block offsets -> every prefix vs fp64
        nb = n_local // cyc_block
        bt = x.view(nb, cyc_block).sum(dim=1, dtype=torch.float64)
        allbt = torch.zeros(world, nb, dtype=torch.float64, device=dev)
        allbt[rank] = bt
        dist.all_reduce(allbt)
        glob = allbt.t().contiguous().view(-1)              # global block order: (round, rank)
        offs = (torch.cumsum(glob, 0) - glob).view(nb, world)[:, rank]
        worst_c = 0.0
        for j in range(nb):
            seg = x[j * cyc_block:(j + 1) * cyc_block].double()
            ref = torch.cumsum(seg, 0) - seg + offs[j]
            err = ((out_c[j * cyc_block:(j + 1) * cyc_block].double() - ref).abs() / ref.abs().clamp_min(1.0)).max()
            worst_c = max(worst_c, float(err.item()))
        res["exclusive_scan_block_cyclic"]["max_rel_err_vs_fp64"] = worst_c
        res["exclusive_scan_block_cyclic"]["parity"] = worst_c <= 1e-5
        del out_c, allbt, bt

    # ---- parity of the fp32 results (every prefix of this shard, fp64 reference on the device)
    parity = True
    tot64 = torch.zeros(world, dtype=torch.float64, device=dev)
    tot64[rank] = x.sum(dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot64)
    ref_total = float(tot64.sum().item())
    parity &= abs(res["reduce_value"] - ref_total) <= 1e-5 * ref_total
    carry64 = float(tot64[:rank].sum().item())
    worst = 0.0
    step = 1 << 27
    run = carry64
    for s0 in range(0, n_local, step):
        seg = x[s0:s0 + step].double()
        inc = torch.cumsum(seg, 0)
        ref = inc - seg + run
        err = ((out[s0:s0 + step].double() - ref).abs() / ref.abs().clamp_min(1.0)).max()
        worst = max(worst, float(err.item()))
        run += float(inc[-1].item())
        del seg, inc, ref
    res["scan_max_rel_err_vs_fp64"] = worst
    parity &= worst <= 1e-5
    if "exclusive_scan_block_cyclic" in res:
        parity &= bool(res["exclusive_scan_block_cyclic"]["parity"])
    del x, out
    torch.cuda.empty_cache()

    # ---- u32, bit-exact: reduce + exclusive scan over 2^26 elements per rank
    M = 0xFFFFFFFF
    n_u = 1 << 26
    xu = torch.empty(n_u, dtype=torch.int32, device=dev)
    fill_input(torch, out_u32=xu, start=rank * n_u)
    ou = torch.empty(n_u, dtype=torch.int32, device=dev)
    ru = torch.zeros(4, dtype=torch.int32, device=dev)
    if world == 1:
        dr.jit_reduce(CUDA, U32, ADD, xu, n_u, ru)
        dr.jit_block_prefix_reduce(CUDA, U32, ADD, n_u, n_u, 1, 0, xu, ou)
    else:
        sh.reduce(U32, ADD, xu, n_u, ru)
        sh.prefix_reduce(U32, ADD, xu, n_u, True, False, ou)
    torch.cuda.synchronize()
    xl = xu.to(torch.int64) & M
    tots = torch.zeros(world, dtype=torch.int64, device=dev)
    tots[rank] = xl.sum() & M
    if world > 1:
        dist.all_reduce(tots)
    parity &= (int(ru[0].item()) & M) == (int(tots.sum().item()) & M)
    ref = (torch.cumsum(xl, 0) - xl + (int(tots[:rank].sum().item()) & M)) & M
    parity &= bool(torch.equal(ou.to(torch.int64) & M, ref))
    del xl, ref, ou, xu

    # ---- mkperm histogram: 2^26 keys per rank, global counts on every rank
    keys = torch.empty(n_u, dtype=torch.int32, device=dev)
    res["mkperm_histogram"] = {"keys_per_gpu": n_u}
    for B in (16, 1024, 65536):
        fill_input(torch, out_u32=keys, start=rank * n_u, mod=B)
        hist_local = torch.zeros(B, dtype=torch.int32, device=dev)
        holder = {}

        def do_hist():
            if world == 1:
                dr.mkperm_histogram(keys, n_u, B, hist_local)
                holder["h"] = hist_local
            else:
                holder["h"] = sh.mkperm_histogram(keys, n_u, B)

        ms = timed_collective(torch, dist, dev, world, barrier, do_hist)
        refh = torch.bincount(keys.to(torch.int64), minlength=B)
        if world > 1:
            dist.all_reduce(refh)
        parity &= bool(torch.equal(holder["h"].to(torch.int64), refh))
        gbs = 4.0 * n_u * world / (ms * 1e-3) / 1e9
        res["mkperm_histogram"][f"B{B}"] = {"ms": ms, "keys_per_s": n_u * world / (ms * 1e-3),
                                            "GBs_aggregate": gbs, "frac_of_n_gpu_peak": gbs / (peak * world)}
    flag = torch.tensor([1 if parity else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["parity"] = bool(int(flag.item()))
    res["parity_checks"] = ("fp32 2^32: reduce and EVERY exclusive prefix within 1e-5 of fp64; u32 2^26/rank: "
                            "reduce and exclusive scan bit-exact; mkperm histogram 16/1024/65536 buckets "
                            "bit-exact (independent torch reference on the device)")

    basis = load_sharded_basis()
    if basis:
        res["speedup_basis_ms"] = basis
        res["speedup_vs_1gpu"] = {k: basis[k] / res[k]["ms"] for k in ("reduce", "exclusive_scan")
                                  if k in basis}
        if "exclusive_scan_block_cyclic" in res and "exclusive_scan" in basis:
            res["speedup_vs_1gpu"]["exclusive_scan_block_cyclic"] = (
                basis["exclusive_scan"] / res["exclusive_scan_block_cyclic"]["ms"])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2N_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-primitives", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="how the sharded primitives exchange their totals (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun "
                             f"(--nproc-per-node {args.gpus}); WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
