"""GPU parity, tier 2: a C++ caller written against the jit.h entry
points (tests/cpp/jit_h_client.cpp, checks modelled on tests/reductions.cpp) linked
with libdrjit_core_b200.so and nothing else.

  jit_h_client_refhdr  compiled against the REFERENCE's own <drjit-core/jit.h>
                       (oracle/Makefile, where /root/reference exists): an
                       unmodified caller of the reference links and runs here;
  jit_h_client         compiled against the mirror header include/drjit_b200_jit.h
                       (drjit-core_b200/csrc/Makefile).
"""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARIES = {
    "reference jit.h": os.path.join(ROOT, "oracle", "_ref", "jit_h_client_refhdr"),
    "mirror header": os.path.join(ROOT, "tests", "cpp", "jit_h_client"),
}


@pytest.mark.parametrize("which", list(BINARIES))
def test_jit_h_client(which):
    binary = BINARIES[which]
    if not os.path.exists(binary):
        pytest.skip(f"{binary} not built")
    p = subprocess.run([binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    m = re.search(r"jit_h_client: (\d+) checks, (\d+) failed", p.stdout)
    assert p.returncode == 0 and m and int(m.group(1)) >= 250 and int(m.group(2)) == 0, \
        p.stdout[-2000:] + p.stderr[-2000:]
