#!/usr/bin/env python
"""Developer tool: throughput of the host-side int16 packing (pnpb200_pack_i16) on pinned memory by thread count and
call size, with no DMA traffic beside it -- the ceiling of the packed leg of the host-buffer pipeline."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import _lib

f = _lib.lib.pnpb200_pack_i16
f.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
chunk = 131072 * 136
src = pnp.host_buffer((8 * chunk,), torch.float64)
src.copy_(torch.randint(0, 1500, (8 * chunk,), dtype=torch.int32).to(torch.float64))
dst = pnp.host_buffer((8 * chunk,), torch.int16)
print("host cores", os.cpu_count())
for values, label in ((chunk, "one chunk (142 MB)"), (8 * chunk, "whole batch (1.14 GB)")):
    for nt in (1, 2, 4, 8, 12, 15, 16):
        f(0, src.data_ptr(), values, dst.data_ptr(), nt)
        reps = 8 if values == chunk else 3
        t = time.perf_counter()
        for r in range(reps):
            off = (r % (8 * chunk // values)) * values
            ok = f(0, src.data_ptr() + 8 * off, values, dst.data_ptr() + 2 * off, nt)
        dt = (time.perf_counter() - t) / reps
        print("%-22s %2d threads: %7.3f ms  %6.1f GB/s of FP64 read  exact=%d" % (label, nt, dt * 1e3, values * 8 / dt / 1e9, ok), flush=True)
