#!/usr/bin/env python
"""Developer tool: touch every kernel once with small, ragged shapes (for compute-sanitizer --tool memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl

K = pt.default_camera_matrix()
for n, B in ((15, 77), (68, 101), (1024, 19)):
    pat = pt.get_golden_pattern() if n == 15 else pt.synthetic_pattern(n)
    keys = list(pat.keys())
    P = pt.pattern_array(pat)
    patd = torch.from_numpy(P).cuda()[None]
    w = wl.synth_batch(3, B, P, K, want_pose=True)
    for dt in (torch.float64, torch.float32):
        uv = w["uv"].to(dt)
        for method in ("qeif", "lm", "linear_f2", "linear_f1", "eif2", "lm_plus"):
            for mapping in (0, 1, 32):
                if (mapping == 1 and n > 256) or (method == "lm_plus" and mapping != 0):
                    continue
                o = pnp.solve_batch(method, uv, patd.to(dt), K, params=pnp.default_params(mapping=mapping))
        if n == 15:
            idx = [keys.index(k) for k in pt.LM_KEY_LIST_6]
            for method in ("qeif", "lm", "linear_f2"):
                o = pnp.solve_batch(method, uv, patd.to(dt), K, point_index=idx)
            two = torch.cat([patd, patd * 1.1]).to(dt)
            o = pnp.solve_batch("qeif", uv, two, K, point_index=idx)
    o = pnp.solve_batch("qeif", w["uv"], patd, K)
    rep = wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])
    st = wl.error_statistics(rep["report"], w["gt"], distributed=False)
    st = wl.statistics([rep["report"][:, 0]], None, None, 1, distributed=False)
    big = torch.randint(0, 64, (B,), dtype=torch.int32, device="cuda")
    st = wl.statistics([rep["report"][:, q] for q in range(4)], [w["gt"][:, q] for q in range(4)], big, 64, distributed=False)
    pnp.project_batch(torch.from_numpy(P).cuda(), K, w["R_gt"], w["t_gt"], True, 0.25)
    pnp.euler_from_R_batch(w["R_gt"], True)
    pnp.R_from_euler_batch(o["euler"], True)
    if n == 15:
        f = wl.synth_face_variation(5, B, P, K, keys.index("eye_c_51"), 0.02)
        ae = rep["report"][:, :4].abs()
        wl.fragility_analysis([ae[:, q] for q in range(4)], f["perturb"], idx0=5, keys=keys)
    host = pnp.HostPipeline(torch.float64, 32, n)
    outs = {"R": torch.empty((B, 3, 3), dtype=torch.float64), "t": torch.empty((B, 3), dtype=torch.float64),
            "res_norm": torch.empty((B,), dtype=torch.float64), "iters": torch.empty((B,), dtype=torch.int32)}
    host.solve("lm", w["uv"].cpu(), torch.from_numpy(P)[None].contiguous(), K, outs)
    host.close()
torch.cuda.synchronize()
print("sanitize_smoke done")
