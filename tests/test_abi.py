"""CPU-only: the C-ABI library builds, loads and exports every symbol that
include/*.h declares, the jit.h-signature layer exports the reference's
MANGLED names, and -- without a GPU -- every entry point fails loudly instead of
falling back to the CPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drjit-core_b200", "libdrjit_core_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import drjit_core_b200 as dr
        dr.build()
    return ctypes.CDLL(LIB)


def exported():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True,
                         check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_c_abi_symbols(lib):
    header = open(os.path.join(ROOT, "include", "drjit_b200.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", header))
    assert len(declared) >= 30
    syms = exported()
    missing = declared - syms
    assert not missing, f"declared in drjit_b200.h but not exported: {sorted(missing)}"
    for name in declared:
        getattr(lib, name)


# Itanium-mangled names of the reference's entry points for this path, as
# exported by an unmodified build of the reference (`nm -D libdrjit-core.so`).
REFERENCE_MANGLED = [
    "_Z8jit_initj", "_Z12jit_shutdowni", "_Z15jit_has_backend10JitBackend",
    "_Z15jit_sync_threadv", "_Z15jit_cuda_streamv", "_Z19jit_cuda_set_devicei",
    "_Z21jit_cuda_device_countv", "_Z10jit_malloc10JitBackendmi", "_Z8jit_freePv",
    "_Z10jit_memcpy10JitBackendPvPKvm", "_Z16jit_memcpy_async10JitBackendPvPKvm",
    "_Z16jit_memset_async10JitBackendPvjjPKv",
    "_Z19jit_reduce_identity7VarType8ReduceOp",
    "_Z22jit_can_scatter_reduce10JitBackend7VarType8ReduceOp",
    "_Z16jit_block_reduce10JitBackend7VarType8ReduceOpjjPKvPv",
    "_Z23jit_block_prefix_reduce10JitBackend7VarType8ReduceOpjjiiPKvPv",
    "_Z12jit_compress10JitBackendPKhjPj",
    "_Z16jit_block_mkperm10JitBackendPKjjjjPjS2_",
    "_Z12jit_set_flag7JitFlagi", "_Z13jit_set_flagsj", "_Z9jit_flagsv", "_Z8jit_flag7JitFlag",
    "_Z18jit_kernel_historyv", "_Z24jit_kernel_history_clearv",
    "_Z18jit_malloc_migratePv10JitBackendi", "_Z20jit_cuda_sync_streamm",
    # jit_reduce as DECLARED in jit.h:2219 (the reference never defines it)
    "_Z10jit_reduce10JitBackend7VarType8ReduceOpPKvjPv",
]


def test_jit_h_mangled_symbols(lib):
    syms = exported()
    missing = [s for s in REFERENCE_MANGLED if s not in syms]
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cpu.so")),
                    reason="oracle/_ref not built")
def test_mangled_names_match_reference_build():
    ref = subprocess.run(["nm", "-D", "--defined-only",
                          os.path.join(ROOT, "oracle", "_ref", "libref_cpu.so")],
                         capture_output=True, text=True, check=True).stdout
    ref_syms = {line.split()[-1] for line in ref.splitlines() if line.strip()}
    # everything except the never-defined jit_reduce overload exists in the reference
    for s in REFERENCE_MANGLED[:-1]:
        assert s in ref_syms, s


def test_no_cpu_fallback(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib.b200_last_error.restype = ctypes.c_char_p
    assert lib.b200_init() != 0
    assert b"no CPU fallback" in lib.b200_last_error()
    buf = (ctypes.c_uint32 * 16)()
    rc = lib.b200_block_reduce(None, 8, 1, ctypes.c_uint64(16), ctypes.c_uint64(16), buf, buf)
    assert rc != 0
    import drjit_core_b200 as dr
    with pytest.raises(RuntimeError):
        dr.jit_init()
    with pytest.raises(RuntimeError):
        dr.jit_block_reduce(dr.JitBackend.LLVM, 8, 1, 16, 16, 0, 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "drjit-core_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("# oracle", ""), os.path.join(dirpath, f)


def test_reduce_identity_table(lib):
    lib.b200_reduce_identity.restype = ctypes.c_uint64
    import oracle
    O = oracle.Oracle()
    for vt in (4, 7, 8, 9, 10, 13, 14, 15):
        for op in range(1, 7):
            assert lib.b200_reduce_identity(vt, op) == O.reduce_identity(vt, op)


def test_reference_header_client_links_against_this_library_only():
    """tests/cpp/jit_h_client.cpp compiled against the REFERENCE's own jit.h
    (oracle/Makefile, target tier3; built where /root/reference exists): every jit_*
    symbol it needs must be one this library exports, under the same mangled name."""
    client = os.path.join(ROOT, "oracle", "_ref", "jit_h_client_refhdr")
    if not os.path.exists(client):
        pytest.skip("oracle/_ref/jit_h_client_refhdr not built")
    out = subprocess.run(["nm", "-u", client], capture_output=True, text=True, check=True).stdout
    needed = {line.split()[-1].split("@")[0] for line in out.splitlines() if "jit_" in line}
    assert len(needed) >= 12, needed
    missing = needed - exported()
    assert not missing, f"needed by a caller of the reference's jit.h but not exported: {sorted(missing)}"
    ldd = subprocess.run(["ldd", client], capture_output=True, text=True).stdout
    assert "libdrjit_core_b200.so" in ldd and "libref" not in ldd and "libdrjit-core" not in ldd


def test_tier3_library_needs_only_exported_b200_symbols():
    """oracle/_ref/libref_cuda_b200.so = the unmodified reference objects + the
    CUDAThreadState adapter (oracle/tier3_adapter.cpp): every b200_* symbol it imports
    must be exported by this library, and the reference's test programs must resolve
    both libraries (so that the GPU run cannot silently skip them)."""
    lib3 = os.path.join(ROOT, "oracle", "_ref", "libref_cuda_b200.so")
    if not os.path.exists(lib3):
        pytest.skip("oracle/_ref/libref_cuda_b200.so not built (make -C oracle tier3)")
    out = subprocess.run(["nm", "-D", "-u", lib3], capture_output=True, text=True, check=True).stdout
    needed = {line.split()[-1].split("@")[0] for line in out.splitlines() if " b200_" in line}
    assert {"b200_block_reduce", "b200_block_prefix_reduce", "b200_reduce_dot", "b200_compress_async",
            "b200_block_mkperm_async", "b200_memset_async"} <= needed
    assert not (needed - exported())
    prog = os.path.join(ROOT, "oracle", "_ref", "test_reductions_b200")
    ldd = subprocess.run(["ldd", prog], capture_output=True, text=True).stdout
    assert "libref_cuda_b200.so" in ldd and "libdrjit_core_b200.so" in ldd and "not found" not in ldd
