"""GPU parity: atomic scatter-reduce through the C-ABI vs the CPU oracle.
Integer results (and float min/max) are bit-exact; float adds are checked
against the fp64-accumulated oracle with the tolerance stated below."""
import numpy as np
import pytest

import oracle
from cases import f32_input, fmix32, u32_input
from util import empty_dev, rel_err, to_dev, to_host

pytestmark = pytest.mark.gpu
VT, OP = oracle.VT, oracle.OP
MODES = {"auto": 0, "direct": 1, "local": 2}


def index_input(n, m, kind):
    i = np.arange(n, dtype=np.uint32)
    if kind == "random":
        return (fmix32(i) % np.uint32(m)).astype(np.uint32)
    if kind == "coherent":  # runs of 64 equal indices (SURVEY.md 8d secondary set)
        return ((i >> np.uint32(6)) % np.uint32(m)).astype(np.uint32)
    if kind == "same":
        return np.full(n, m // 2, dtype=np.uint32)
    if kind == "runs":  # ragged runs of 1..7 equal indices that straddle warp boundaries
        return ((np.cumsum(fmix32(i) % np.uint32(7) == 0) // 1) % m).astype(np.uint32)
    return ((i % np.uint32(3)) + (fmix32(i) % np.uint32(5) == 0) * (m - 3)).astype(np.uint32)  # few hot slots


def run_scatter(dr, vt, op, target, value, index, mask, mode):
    d_t = to_dev(target)
    dr.scatter_reduce(vt, op, d_t, to_dev(value), to_dev(index),
                      None if mask is None else to_dev(mask), index.size, mode=mode)
    return to_host(d_t, target.dtype)


@pytest.mark.parametrize("tname", ["u32", "i32", "u64", "i64"])
def test_scatter_int(dr, O, tname):
    dt = oracle.NP_OF_VT[VT[tname]]
    bad = []
    for n, m in ((1, 1), (1000, 7), (100003, 997), (1 << 20, 1 << 12)):
        h = u32_input(n)
        if tname[1:] == "32":
            val = h.view(dt)
        else:
            val = ((h.astype(np.uint64) << np.uint64(20)) ^ h.astype(np.uint64)).view(dt)
        mask = (fmix32(h) & np.uint32(3) != 0).astype(np.uint8)
        for kind in ("random", "coherent", "same", "hot", "runs"):
            idx = index_input(n, m, kind)
            for opn in ("add", "min", "max", "and_", "or_"):
                ident = O.reduce_identity(VT[tname], OP[opn])
                if tname[1:] == "32":
                    tgt = np.full(m, ident & 0xFFFFFFFF, dtype=np.uint32).view(dt)
                else:
                    tgt = np.full(m, ident, dtype=np.uint64).view(dt)
                for mname, mode in MODES.items():
                    for mk in (None, mask):
                        got = run_scatter(dr, VT[tname], OP[opn], tgt, val, idx, mk, mode)
                        ref = O.scatter_reduce(VT[tname], OP[opn], tgt, val, idx, mk)
                        if not np.array_equal(got, ref):
                            bad.append((n, m, kind, opn, mname, mk is not None))
    assert not bad, bad[:20]


# float add: the order of atomic additions is not defined, so compare against
# the fp64-accumulated oracle.  Per slot with c addends (all positive):
# relative error <= tol * max(1, sqrt(c / 4096)), tol = 1e-5 (f32) / 1e-12 (f64)
# -- a slot that serially accumulates hundreds of thousands of addends (the
# "hot" pattern) legitimately drifts by a few 1e-5.
@pytest.mark.parametrize("tname,tol", [("f32", 1e-5), ("f64", 1e-12)])
def test_scatter_float(dr, O, tname, tol):
    dt = oracle.NP_OF_VT[VT[tname]]
    bad = []
    for n, m in ((1000, 7), (1 << 20, 1 << 12), (100003, 997)):
        val = f32_input(n).astype(dt)
        sval = (val * 4 - 2).astype(dt)
        for kind in ("random", "coherent", "hot", "runs"):
            idx = index_input(n, m, kind)
            for mname, mode in MODES.items():
                got = run_scatter(dr, VT[tname], OP["add"], np.zeros(m, dtype=dt), val, idx, None, mode)
                ref = O.scatter_reduce(VT[tname], OP["add"], np.zeros(m, dtype=dt), val, idx, wide=True)
                sel = ref != 0
                cnt = np.bincount(idx, minlength=m)[sel]
                slot_tol = tol * np.maximum(1.0, np.sqrt(cnt / 4096.0))
                err = np.abs(got[sel].astype(np.float64) - ref[sel]) / np.abs(ref[sel].astype(np.float64))
                if np.any(err > slot_tol) or np.any(got[~sel] != 0):
                    bad.append((n, m, kind, mname, float(err.max())))
                for opn in ("min", "max"):  # emulated with integer atomics: exact
                    ident = np.array([np.inf if opn == "min" else -np.inf], dtype=dt)[0]
                    tgt = np.full(m, ident, dtype=dt)
                    got = run_scatter(dr, VT[tname], OP[opn], tgt, sval, idx, None, mode)
                    ref = O.scatter_reduce(VT[tname], OP[opn], tgt, sval, idx)
                    if not np.array_equal(got, ref):
                        bad.append((n, m, kind, mname, opn))
    assert not bad, bad[:20]


def test_scatter_f16(dr, O):
    bad = []
    n, m = 20000, 501
    val = (f32_input(n) * 0.01).astype(np.float16)
    for kind in ("random", "hot"):
        idx = index_input(n, m, kind)
        for mname, mode in MODES.items():
            got = run_scatter(dr, VT["f16"], OP["add"], np.zeros(m, dtype=np.float16), val, idx, None, mode)
            ref = O.scatter_reduce(VT["f16"], OP["add"], np.zeros(m, dtype=np.float16), val, idx, wide=True)
            # every addition rounds to f16: allow count * ulp/2 accumulated error
            cnt = np.bincount(idx, minlength=m)
            tol = np.maximum(cnt, 1) * (2.0 ** -11) * np.maximum(ref.astype(np.float64), 1e-3)
            if np.any(np.abs(got.astype(np.float64) - ref.astype(np.float64)) > tol):
                bad.append((kind, mname, "add"))
            for opn in ("min", "max"):
                tgt = np.full(m, np.inf if opn == "min" else -np.inf, dtype=np.float16)
                sval = (val * 100 - 1).astype(np.float16)
                got = run_scatter(dr, VT["f16"], OP[opn], tgt, sval, idx, None, mode)
                ref = O.scatter_reduce(VT["f16"], OP[opn], tgt, sval, idx)
                if not np.array_equal(got, ref):
                    bad.append((kind, mname, opn))
    assert not bad, bad


def test_scatter_reference_test_vector(dr):
    # tests/mem.cpp:137-174 (10_scatter_atomic_rmw): 16 adds of 1.0 with duplicates
    index = np.array([0, 0, 1, 2, 2, 2, 3, 4, 4, 4, 4, 0, 1, 2, 3, 4], dtype=np.uint32)
    one = np.ones(16, dtype=np.float32)
    got = run_scatter(dr, VT["f32"], OP["add"], np.zeros(5, dtype=np.float32), one, index, None, 0)
    assert np.array_equal(got, np.bincount(index, minlength=5).astype(np.float32))
    mask = (np.arange(16) % 2 == 0).astype(np.uint8)
    got = run_scatter(dr, VT["f32"], OP["add"], np.zeros(5, dtype=np.float32), one, index, mask, 0)
    assert np.array_equal(got, np.bincount(index[mask != 0], minlength=5).astype(np.float32))


def test_scatter_no_conflicts_and_errors(dr, O):
    n = 100000
    idx = np.random.default_rng(3).permutation(n).astype(np.uint32)
    val = u32_input(n)
    tgt = fmix32(val)
    got = run_scatter(dr, VT["u32"], OP["add"], tgt, val, idx, None, 3)
    assert np.array_equal(got, O.scatter_reduce(VT["u32"], OP["add"], tgt, val, idx))
    d = to_dev(val)
    # src/op.cpp:2735-2820: no Mul, no And/Or on floats, no 8-bit types
    assert not dr.jit_can_scatter_reduce(1, VT["u32"], OP["mul"])
    assert not dr.jit_can_scatter_reduce(1, VT["f32"], OP["and_"])
    assert not dr.jit_can_scatter_reduce(1, VT["u8"], OP["add"])
    assert dr.jit_can_scatter_reduce(1, VT["f16"], OP["max"])
    with pytest.raises(RuntimeError, match="does not support"):
        dr.scatter_reduce(VT["u32"], OP["mul"], d, d, d, None, 10)
    with pytest.raises(RuntimeError, match="does not support"):
        dr.scatter_reduce(VT["f32"], OP["or_"], d, d, d, None, 10)


def test_scatter_full_size(dr, O):
    # BASELINE.json configs[4] (per device part): 2^26 -> 2^20 ScatterAdd
    n, m = 1 << 26, 1 << 20
    idx = (fmix32(np.arange(n, dtype=np.uint32)) & np.uint32(m - 1)).astype(np.uint32)
    ival = u32_input(n)
    got = run_scatter(dr, VT["u32"], OP["add"], np.zeros(m, dtype=np.uint32), ival, idx, None, 0)
    assert np.array_equal(got, O.scatter_reduce(VT["u32"], OP["add"], np.zeros(m, dtype=np.uint32), ival, idx))
    fval = f32_input(n)
    got = run_scatter(dr, VT["f32"], OP["add"], np.zeros(m, dtype=np.float32), fval, idx, None, 1)
    ref = O.scatter_reduce(VT["f32"], OP["add"], np.zeros(m, dtype=np.float32), fval, idx, wide=True)
    assert rel_err(got, ref) <= 1e-5


def test_scatter_inc(dr, O):
    # jit_var_scatter_inc (tests/mem.cpp:223-310 uses it for queue compaction):
    # one shared counter, a few counters, runs, random; with and without a mask
    bad = []
    for n, m in ((1, 1), (1000, 1), (100003, 1), (100003, 7), (1 << 20, 1 << 10)):
        h = u32_input(n)
        mask = (fmix32(h) & np.uint32(3) != 0).astype(np.uint8)
        for kind in ("same", "random", "coherent", "runs"):
            idx = index_input(n, m, kind)
            for mk in (None, mask):
                before = (fmix32(np.arange(m, dtype=np.uint32)) & np.uint32(0xffff)).astype(np.uint32)
                d_t = to_dev(before)
                d_o = empty_dev(n, np.uint32)
                d_o.fill_(-1)
                dr.scatter_inc(d_t, to_dev(idx), None if mk is None else to_dev(mk), d_o, n)
                viol = O.scatter_inc_check(before, to_host(d_t, np.uint32), idx, mk, to_host(d_o, np.uint32))
                if viol:
                    bad.append((n, m, kind, mk is not None, viol))
    assert not bad, bad[:10]


@pytest.mark.parametrize("width", [1, 2, 4, 8])
def test_scatter_reduce_packet(dr, O, width):
    # target[index * W + k] op= values[k]: integers bit-exact, f32 / f64 Add within the
    # tolerance of the scalar scatter tests (the order of the atomics is unspecified)
    bad = []
    n, m = 100003, 997
    for kind in ("random", "coherent", "runs", "same"):
        idx = index_input(n, m, kind)
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        for mode in (1, 2, 0):
            for mk in (None, mask):
                ivals = [(u32_input(n) >> np.uint32(8 + k)).astype(np.uint32) for k in range(width)]
                for opn in ("add", "min", "max", "and_", "or_"):
                    ident = O.reduce_identity(VT["u32"], OP[opn]) & 0xFFFFFFFF
                    tgt = np.full(m * width, ident, dtype=np.uint32)
                    d_t = to_dev(tgt)
                    dr.scatter_reduce_packet(VT["u32"], OP[opn], d_t, [to_dev(v) for v in ivals], to_dev(idx),
                                             None if mk is None else to_dev(mk), n, mode=mode)
                    ref = O.scatter_reduce_packet(VT["u32"], OP[opn], tgt, ivals, idx, mk)
                    if not np.array_equal(to_host(d_t, np.uint32), ref):
                        bad.append((kind, mode, mk is not None, opn))
                fvals = [((u32_input(n) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) + np.float32(k))
                         for k in range(width)]
                d_t = to_dev(np.zeros(m * width, dtype=np.float32))
                dr.scatter_reduce_packet(VT["f32"], OP["add"], d_t, [to_dev(v) for v in fvals], to_dev(idx),
                                         None if mk is None else to_dev(mk), n, mode=mode)
                ref = O.scatter_reduce_packet(VT["f32"], OP["add"], np.zeros(m * width, dtype=np.float32),
                                              fvals, idx, mk, wide=True)
                got = to_host(d_t, np.float32)
                if not np.allclose(got, ref, rtol=2e-5, atol=1e-30):  # fp32 Add: 2e-5 relative
                    bad.append((kind, mode, mk is not None, "f32 add"))
                for opn in ("min", "max"):
                    tgt = np.full(m * width, np.inf if opn == "min" else -np.inf, dtype=np.float32)
                    svals = [v - np.float32(0.5) for v in fvals]
                    d_t = to_dev(tgt)
                    dr.scatter_reduce_packet(VT["f32"], OP[opn], d_t, [to_dev(v) for v in svals], to_dev(idx),
                                             None if mk is None else to_dev(mk), n, mode=mode)
                    ref = O.scatter_reduce_packet(VT["f32"], OP[opn], tgt, svals, idx, mk)
                    if not np.array_equal(to_host(d_t, np.float32), ref):
                        bad.append((kind, mode, mk is not None, "f32 " + opn))
    assert not bad, bad[:10]
    with pytest.raises(RuntimeError):  # vector size must be a power of two
        dr.scatter_reduce_packet(VT["f32"], OP["add"], 0, [0, 0, 0], 0, None, 1)


@pytest.mark.parametrize("width", [1, 2, 4, 8])
def test_scatter_packet_f16_add(dr, O, width):
    # red.global.v2 / .v4 / .v8.f16.add.noftz (src/cuda_packet.cpp:229-266): every addition
    # rounds to f16 -> count * ulp / 2 of accumulated error allowed, as in test_scatter_f16
    bad = []
    n, m = 20011, 509
    for kind in ("random", "runs", "same"):
        idx = index_input(n, m, kind)
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        vals = [((f32_input(n) * 0.01) + 0.001 * k).astype(np.float16) for k in range(width)]
        for mname, mode in MODES.items():
            for mk in (None, mask):
                d_t = to_dev(np.zeros(m * width, dtype=np.float16))
                dr.scatter_reduce_packet(VT["f16"], OP["add"], d_t, [to_dev(v) for v in vals], to_dev(idx),
                                         None if mk is None else to_dev(mk), n, mode=mode)
                got = to_host(d_t, np.float16).astype(np.float64)
                ref = O.scatter_reduce_packet(VT["f16"], OP["add"], np.zeros(m * width, dtype=np.float16),
                                              vals, idx, mk, wide=True).astype(np.float64)
                on = np.ones(n, bool) if mk is None else mk.astype(bool)
                cnt = np.repeat(np.bincount(idx[on], minlength=m), width)
                tol = np.maximum(cnt, 1) * (2.0 ** -11) * np.maximum(ref, 1e-3)
                if np.any(np.abs(got - ref) > tol):
                    bad.append((kind, mname, mk is not None))
    assert not bad, bad[:10]
    if width > 1:
        with pytest.raises(RuntimeError):  # misaligned target of a vector reduction
            t = to_dev(np.zeros(m * width + 1, dtype=np.float16))
            dr.scatter_reduce_packet(VT["f16"], OP["add"], t.data_ptr() + 2, [0] * width, 0, None, 1)


@pytest.mark.parametrize("tname,dt", [("u8", np.uint8), ("f16", np.float16), ("u32", np.uint32), ("f32", np.float32),
                                      ("f64", np.float64), ("u64", np.uint64)])
def test_packet_scatter_and_gather(dr, O, tname, dt):
    # non-reducing packet scatter (src/cuda_packet.cpp:329-443) with a permutation as index
    # (no duplicate targets: the result is defined), packet gather (:18-166) with arbitrary
    # indices; both bit-exact, every width, masked and unmasked, odd base alignment
    bad = []
    n = 50021
    vt = VT[tname]
    rng = np.random.default_rng(7)
    perm = rng.permutation(n).astype(np.uint32)
    gidx = (fmix32(u32_input(n)) % np.uint32(n)).astype(np.uint32)
    mask = (fmix32(u32_input(n) ^ np.uint32(0xABCD)) & np.uint32(3) != 0).astype(np.uint8)
    for width in (1, 2, 4, 8):
        vals = [(u32_input(n) >> np.uint32(k)).astype(dt) if np.issubdtype(dt, np.integer)
                else ((u32_input(n) >> np.uint32(8 + k)).astype(np.float32) * np.float32(2.0 ** -20)).astype(dt)
                for k in range(width)]
        for off in (0, 1):  # element offset of the AoS base: exercises the narrower chunk sizes
            for mk in (None, mask):
                base = np.full(n * width + off, 7, dtype=dt)
                d_b = to_dev(base)
                view = d_b.data_ptr() + off * base.itemsize
                dr.scatter_packet(vt, view, [to_dev(v) for v in vals], to_dev(perm),
                                  None if mk is None else to_dev(mk), n)
                ref = O.scatter_packet(base[off:], vals, perm, mk)
                got = to_host(d_b, dt)
                if not np.array_equal(got[off:].view(np.uint8), ref.view(np.uint8)) or (off and got[0] != 7):
                    bad.append(("scatter", width, off, mk is not None))
                src = np.concatenate([np.zeros(off, dtype=dt), np.stack(vals, axis=1).reshape(-1)])
                d_s = to_dev(src)
                outs = [empty_dev(n, dt) for _ in range(width)]
                dr.gather_packet(vt, d_s.data_ptr() + off * src.itemsize, outs, to_dev(gidx),
                                 None if mk is None else to_dev(mk), n)
                refs = O.gather_packet(src[off:], width, gidx, mk)
                for k in range(width):
                    if not np.array_equal(to_host(outs[k], dt).view(np.uint8), refs[k].view(np.uint8)):
                        bad.append(("gather", width, off, mk is not None, k))
    assert not bad, bad[:10]
    with pytest.raises(RuntimeError):  # vector size must be a power of two
        dr.scatter_packet(vt, 0, [0, 0, 0], 0, None, 1)


@pytest.mark.parametrize("iname,idt", [("i32", np.int32), ("u64", np.uint64), ("i64", np.int64)])
def test_scatter_index_types_and_identity(dr, O, iname, idt):
    # jitc_var_scatter accepts int32 / uint64 / int64 indices (src/op.cpp:2899-3086);
    # ReduceOp::Identity is the plain scatter
    bad = []
    n, m = 70001, 4099
    for kind in ("random", "runs", "same"):
        idx = index_input(n, m, kind)
        mask = (fmix32(u32_input(n)) & np.uint32(3) != 0).astype(np.uint8)
        ival = (u32_input(n) >> np.uint32(9)).astype(np.uint32)
        fval = f32_input(n)
        for mname, mode in MODES.items():
            for mk in (None, mask):
                d_m = None if mk is None else to_dev(mk)
                for opn in ("add", "min", "max", "and_", "or_"):
                    ident = O.reduce_identity(VT["u32"], OP[opn]) & 0xFFFFFFFF
                    tgt = np.full(m, ident, dtype=np.uint32)
                    d_t = to_dev(tgt)
                    dr.scatter_reduce_idx(VT["u32"], OP[opn], d_t, to_dev(ival), to_dev(idx.astype(idt)), VT[iname],
                                          d_m, n, mode=mode)
                    if not np.array_equal(to_host(d_t, np.uint32), O.scatter_reduce(VT["u32"], OP[opn], tgt, ival, idx, mk)):
                        bad.append((kind, mname, mk is not None, opn))
                d_t = to_dev(np.zeros(m, dtype=np.float32))
                dr.scatter_reduce_idx(VT["f32"], OP["add"], d_t, to_dev(fval), to_dev(idx.astype(idt)), VT[iname],
                                      d_m, n, mode=mode)
                ref = O.scatter_reduce(VT["f32"], OP["add"], np.zeros(m, dtype=np.float32), fval, idx, mk, wide=True)
                if not np.allclose(to_host(d_t, np.float32), ref, rtol=2e-5, atol=1e-30):  # fp32 Add: 2e-5 relative
                    bad.append((kind, mname, mk is not None, "f32 add"))
    # plain scatter through a permutation (no duplicates): defined result, bit-exact
    perm = np.random.default_rng(3).permutation(n).astype(idt)
    for tname, dt in (("u8", np.uint8), ("f16", np.float16), ("f32", np.float32), ("f64", np.float64)):
        val = (u32_input(n) >> np.uint32(5)).astype(dt)
        for mk in (None, mask):
            tgt = np.full(n, 3, dtype=dt)
            d_t = to_dev(tgt)
            dr.scatter_reduce_idx(VT[tname], OP["identity"] if "identity" in OP else 0, d_t, to_dev(val), to_dev(perm),
                                  VT[iname], None if mk is None else to_dev(mk), n)
            ref = O.scatter_packet(tgt, [val], perm.astype(np.int64), mk)
            if not np.array_equal(to_host(d_t, dt).view(np.uint8), ref.view(np.uint8)):
                bad.append(("identity", tname, mk is not None))
    assert not bad, bad[:10]
    assert dr.jit_can_scatter_reduce(1, VT["f32"], 0)  # Identity: the plain scatter exists
