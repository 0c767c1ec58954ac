#!/usr/bin/env python
"""Developer tool: wall-clock cost of one PNP_SOLVER.solve_pnp(dict) call (the per-sample API the reference's scripts use)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl

K = pt.default_camera_matrix()
pat = pt.get_golden_pattern("Alexander")
keys = list(pat)
solver = pnp.PNP_SOLVER(K, [pat, pt.get_golden_pattern("Holly")], [1.0, 1.0])
uv = wl.synth_batch(0, 256, pt.pattern_array(pat), K)["uv"].cpu().numpy()
dicts = [{k: np.array([[uv[b, i, 0]], [uv[b, i, 1]], [1.0]]) for i, k in enumerate(keys)} for b in range(256)]
for d in dicts[:20]:
    solver.solve_pnp(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for d in dicts:
    r = solver.solve_pnp(d)
t1 = time.perf_counter()
print("solve_pnp(dict), two patterns, QEIF on 6 landmarks: %.1f us per call" % ((t1 - t0) / len(dicts) * 1e6))
