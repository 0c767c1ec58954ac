#!/usr/bin/env python
"""Developer tool: host-side cost of one solve_batch call (tiny batch, wall clock per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl

K = pt.default_camera_matrix()
pat = pt.get_golden_pattern()
P = pt.pattern_array(pat)
idx = [list(pat).index(k) for k in pt.LM_KEY_LIST_6]
w = wl.synth_batch(0, 64, P, K)
patd = torch.from_numpy(P).cuda()[None]
for method, kw in (("qeif", dict(point_index=idx)), ("qeif", {}), ("lm", {}), ("linear_f2", {}), ("eif2", {})):
    for _ in range(20):
        pnp.solve_batch(method, w["uv"], patd, K, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        pnp.solve_batch(method, w["uv"], patd, K, **kw)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-10s %-22s host %.1f us per call (%.1f us incl. drain)" % (method, "subset6" if kw else "", (t1 - t0) / 500 * 1e6, (t2 - t0) / 500 * 1e6))
