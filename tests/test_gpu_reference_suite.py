"""GPU parity, tier 3: the reference's OWN test programs
(tests/reductions.cpp, mem.cpp, vcall.cpp, record.cpp, loop.cpp, basics.cpp,
array.cpp -- compiled where they lie by oracle/Makefile) run against the reference
runtime whose CUDAThreadState primitives forward to libdrjit_core_b200.so
(oracle/tier3_adapter.cpp).  Every assertion in those programs is the reference's.

`-c` = CUDA tests only, `-t` = do not diff the trace logs against recorded ones
(tests/test.cpp:285-296, :402): pass / fail is decided by the tests' own
jit_assert()s and exceptions.
"""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PROGRAMS = ["reductions", "mem", "vcall", "record", "loop", "basics", "array"]
# programs whose tests reach the primitives through the variable layer
MUST_FORWARD = {"reductions", "mem", "vcall", "record"}


def run(binary, tmp_path):
    env = dict(os.environ)
    env["HOME"] = str(tmp_path)  # the reference keeps a kernel cache in ~/.drjit
    p = subprocess.run([binary, "-c", "-t"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900)
    m = re.search(r"Passed (\d+)/(\d+) tests", p.stdout)
    return p, m


@pytest.mark.parametrize("name", PROGRAMS)
def test_reference_program_on_b200_kernels(name, tmp_path):
    binary = os.path.join(REF, f"test_{name}_b200")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref not built (make -C oracle tier3)")
    p, m = run(binary, tmp_path)
    tail = p.stdout[-3000:] + p.stderr[-3000:]
    assert m, tail
    passed, total = int(m.group(1)), int(m.group(2))
    assert p.returncode == 0 and passed == total and total > 0, tail
    fw = re.search(r"tier3_adapter: (\d+) primitive calls forwarded .*\((\d+) kernel launches\)", p.stderr)
    if name in MUST_FORWARD:
        assert fw and int(fw.group(1)) > 0 and int(fw.group(2)) > 0, tail
    print(f"{name}: {passed}/{total} reference tests passed; {fw.group(0) if fw else 'no primitive calls'}")


def test_reference_program_on_reference_kernels(tmp_path):
    """Control: the same program on the unmodified reference (driver-JITed compute_75 PTX)."""
    binary = os.path.join(REF, "test_reductions_ref")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref not built (make -C oracle tier3)")
    p, m = run(binary, tmp_path)
    assert m and p.returncode == 0 and int(m.group(1)) == int(m.group(2)) > 0, p.stdout[-3000:] + p.stderr[-3000:]
