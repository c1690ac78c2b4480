"""SASS opcode histogram per kernel of the shipped library (development aid / evidence
that the kernels are sm_100a code using the Blackwell copy engine and warp reductions):

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt

Reads drjit-core_b200/libdrjit_core_b200.so with cuobjdump; per kernel prints the
instruction count and the opcodes of interest (UBLKCP = cp.async.bulk, SYNCS = mbarrier,
REDUX / CREDUX = warp reductions, ATOMS = shared atomics, RED / REDG = global reductions,
LDG.E.128 / STG.E.128 = 128-bit global accesses, MATCH)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drjit-core_b200", "libdrjit_core_b200.so")
KEYS = ("UBLKCP", "SYNCS", "REDUX", "CREDUX", "ATOMS", "ATOMG", "RED", "REDG", "LDG", "STG", "LDS", "STS",
        "SHFL", "VOTE", "MATCH", "BAR", "UTMA", "ELECT", "FENCE", "CCTL", "NANOSLEEP")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: architectures {archs}")
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    dem = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {len(kernels)} kernels; columns: instructions | opcodes of interest (count)")
    for (mangled, ops), pretty in zip(kernels.items(), dem):
        total = sum(ops.values())
        short = collections.Counter()
        wide = collections.Counter()
        for op, c in ops.items():
            base = op.split(".")[0]
            if base in KEYS:
                short[base] += c
            if base in ("LDG", "STG", "LDS", "STS") and ".128" in op:
                wide[base + ".128"] += c
            if base in ("RED", "REDG", "ATOMG") and (".128" in op or ".64" in op or "F32" in op or "F16" in op):
                wide[op] += c
        items = [f"{k}={v}" for k, v in sorted(short.items())] + [f"{k}={v}" for k, v in sorted(wide.items())]
        pretty = re.sub(r"\s+", " ", pretty)
        print(f"{pretty[:110]:110s} | {total:6d} | {' '.join(items)}")


if __name__ == "__main__":
    sys.exit(main())
