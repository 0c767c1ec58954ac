#!/usr/bin/env python
"""Developer tool: device time of each stage of the face_variation_test flow (workloads config 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl

B = 1 << 20
K = pt.default_camera_matrix()
pat = pt.get_golden_pattern("Alexander")
keys = list(pat.keys())
P = pt.pattern_array(pat)
solver = pnp.PNP_SOLVER(K, [pat], [1.0])
fixed = keys.index("eye_c_51")


def t(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms, w = t(lambda: wl.synth_face_variation(0, B, P, K, fixed, 0.02)); print("synth_face_variation %.3f ms" % ms)
ms, _ = t(lambda: wl.synth_batch(0, B, P, K)); print("synth_batch           %.3f ms" % ms)
ms, o = t(lambda: solver.solve_pnp_batch(w["uv"])); print("solve (QEIF-6)        %.3f ms" % ms)
ms, rep = t(lambda: wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])); print("report                %.3f ms" % ms)
ms, _ = t(lambda: wl.error_statistics(rep["report"], w["gt"], lazy=True)); print("statistics            %.3f ms" % ms)
ae = rep["report"][:, :4].abs()
vals = [ae[:, q] for q in range(4)]
ms, thr = t(lambda: wl.topk_thresholds(vals, B // 10)); print("topk_thresholds       %.3f ms" % ms)
ms, _ = t(lambda: wl.fragility_analysis(vals, w["perturb"], keys=keys)); print("fragility_analysis    %.3f ms (incl. thresholds, eigh on host)" % ms)
