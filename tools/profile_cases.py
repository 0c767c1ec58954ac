#!/usr/bin/env python
"""Developer tool: run a list of solve cases once each inside a cudaProfilerStart/Stop range (after a warm-up
of every case outside it), so that one `ncu --profile-from-start off` pass captures exactly one launch of each
kernel of each case; without ncu it prints the per-kernel CUDA-event times of the same launches.

    python tools/profile_cases.py lm:1024:100000:f64:0 linear_f2:1024:100000:f64:0 qeif:1024:100000:f64:32 \
                                  lm:68:1048576:f32:0 qeif6:15:1048576:f32:0

case = method:n:B:dtype:mapping   (method `qeif6` = QEIF on the 6-landmark subset of the 15-point pattern)
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import _lib, workload as wl, patterns as pt

cases = []
for spec in sys.argv[1:]:
    method, n, B, dtype, mapping = spec.split(":")
    n, B, mapping = int(n), int(B), int(mapping)
    dt = torch.float64 if dtype == "f64" else torch.float32
    pat = pt.get_golden_pattern() if n <= 15 else pt.synthetic_pattern(n)
    P = pt.pattern_array(pat)
    K = pt.default_camera_matrix()
    w = wl.synth_batch(0, B, P, K, dtype=dt)
    idx = None
    if method == "qeif6":
        method, idx = "qeif", [list(pat).index(k) for k in pt.LM_KEY_LIST_6]
    cases.append(dict(spec=spec, method=method, uv=w["uv"], pat=torch.from_numpy(P).cuda().to(dt)[None], K=K, idx=idx,
                      prm=pnp.default_params(mapping=mapping, flags=_lib.FLAG_PROFILE), B=B, n=n, esz=8 if dtype == "f64" else 4))


def run(c):
    return pnp.solve_batch(c["method"], c["uv"], c["pat"], c["K"], point_index=c["idx"], params=c["prm"])


for c in cases:
    for _ in range(3):
        run(c)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for c in cases:
    _lib.lib.pnpb200_profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run(c)
    e1.record()
    torch.cuda.synchronize()
    ms = (C.c_float * 3)()
    nc = C.c_int()
    _lib.lib.pnpb200_profile_read(ms, C.byref(nc))
    row_gb = c["B"] * c["n"] * 2 * c["esz"] / 1e9
    print("%-34s total %.3f ms | kernels [%.3f %.3f %.3f] ms | pixel rows %.3f GB -> pass0 %.0f GB/s pass2 %.0f GB/s | mean iters %.2f"
          % (c["spec"], e0.elapsed_time(e1), ms[0], ms[1], ms[2], row_gb, row_gb / max(ms[0], 1e-9) * 1e3,
             row_gb / max(ms[2], 1e-9) * 1e3 if ms[2] > 0 else 0.0, float(out["iters"].double().mean())), flush=True)
torch.cuda.profiler.stop()
