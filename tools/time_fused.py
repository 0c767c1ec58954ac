#!/usr/bin/env python
"""Developer tool: per-kernel device times of pnpb200_solve_report_batch (moments, iterate, fused report + residual)
on the headline workload, for the library selected by PNPB200_LIB (tools/build_variant.py)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import _lib, workload as wl, patterns as pt

B, n = 1 << 20, 68
K = pt.default_camera_matrix()
P = pt.pattern_array(pt.synthetic_pattern(n))
w = wl.synth_batch(0, B, P, K)
patd = torch.from_numpy(P).cuda()[None]
prm = pnp.default_params(flags=_lib.FLAG_PROFILE)
for _ in range(3):
    o = wl.solve_report_batch("lm", w["uv"], patd, K, w["gt"], params=prm)
torch.cuda.synchronize()
_lib.lib.pnpb200_profile_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 20
for _ in range(reps):
    o = wl.solve_report_batch("lm", w["uv"], patd, K, w["gt"], params=prm)
e1.record()
torch.cuda.synchronize()
ms = (C.c_float * 3)()
nc = C.c_int()
_lib.lib.pnpb200_profile_read(ms, C.byref(nc))
chk = float(o["report"][:, 4].nan_to_num(0.0, 0.0, 0.0).clamp(-1e3, 1e3).sum()) + float(o["res_norm"].nan_to_num(0.0, 0.0, 0.0).clamp(0, 1e3).sum())
print("%-60s total %.4f ms | moments %.4f iterate %.4f report+residual %.4f | checksum %.9e"
      % (os.path.basename(os.environ.get("PNPB200_LIB", "libpnpb200.so")), e0.elapsed_time(e1) / reps, ms[0], ms[1], ms[2], chk), flush=True)
