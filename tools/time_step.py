#!/usr/bin/env python
"""Developer tool: device time of each stage of one bench step (solve / report / statistics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl

B, n = 1 << 20, 68
K = pt.default_camera_matrix()
P = pt.pattern_array(pt.synthetic_pattern(n))
w = wl.synth_batch(0, B, P, K)
patd = torch.from_numpy(P).cuda()[None]
out = pnp.solve_batch("lm", w["uv"], patd, K)
rep = wl.report_batch(P, w["uv"], K, out["R"], out["t"], out["euler"], w["gt"])


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("solve  %.3f ms" % t(lambda: pnp.solve_batch("lm", w["uv"], patd, K)))
print("report %.3f ms" % t(lambda: wl.report_batch(P, w["uv"], K, out["R"], out["t"], out["euler"], w["gt"])))
print("stats  %.3f ms" % t(lambda: wl.error_statistics(rep["report"], w["gt"], lazy=True)))
