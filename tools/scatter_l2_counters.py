"""One launch of the scatter-add 2^26 -> 2^20 (random indices, Direct) of this library and
of the REFERENCE's own JIT kernel (oracle/_ref/libref_cuda.so) on the same buffers, for

    ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,\
lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,\
lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,\
lts__t_sectors_srcunit_tex_op_red.sum,l1tex__m_xbar2l1tex_read_sectors.sum \
        --clock-control none python tools/scatter_l2_counters.py

(development aid: the L2 reduction-unit traffic next to the HBM fraction)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjit_core_b200 as dr  # noqa: E402
import oracle  # noqa: E402

F32, ADD = 14, 1


def main():
    dr.jit_init()
    n2, m2 = 1 << 26, 1 << 20
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, m2, (n2,), device="cuda", dtype=torch.int32, generator=g)
    val = torch.rand(n2, device="cuda")
    tgt = torch.zeros(m2, device="cuda")
    for mode in (1, 2):  # Direct, Local
        dr.scatter_reduce(F32, ADD, tgt, val, idx, None, n2, mode=mode)
    coh = (torch.arange(n2, device="cuda", dtype=torch.int64) >> 6).to(torch.int32) % m2
    dr.scatter_reduce(F32, ADD, tgt, val, coh, None, n2, mode=2)
    torch.cuda.synchronize()
    if oracle.ref_cuda_available():
        R = oracle.RefCuda.get()
        for mode in (1, 2):
            R.scatter_reduce(F32, ADD, mode, tgt.data_ptr(), m2, val.data_ptr(), idx.data_ptr(), None, n2, 1)
        R.scatter_reduce(F32, ADD, 2, tgt.data_ptr(), m2, val.data_ptr(), coh.data_ptr(), None, n2, 1)
        R.sync()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
