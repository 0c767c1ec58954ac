#!/usr/bin/env python
"""Developer tool: soak of the n = 1024 one-problem-per-warp kernels (TMA ring) over many batch sizes, every result
checked against the same problems solved in another batch position (the kernels are batch-position independent)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import workload as wl, patterns as pt
n = 1024
P = pt.pattern_array(pt.synthetic_pattern(n)); K = pt.default_camera_matrix()
w = wl.synth_batch(0, 30000, P, K)
patd = torch.from_numpy(P).cuda()[None]
ref = {m: pnp.solve_batch(m, w["uv"], patd, K) for m in ("lm", "linear_f2", "qeif", "eif2")}
torch.cuda.synchronize()
n_runs = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for B in (1, 7, 8, 9, 63, 300, 1183, 1184, 1185, 2368, 5000, 9473, 20000, 30000):
        for m in ref:
            o = pnp.solve_batch(m, w["uv"][:B].contiguous(), patd, K)
            torch.cuda.synchronize()
            for k in ("R", "t", "res_norm", "iters"):
                a, b = o[k], ref[m][k][:B]
                assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b), (m, B, k)
            n_runs += 1
print("soak ok: %d solves, all bit-identical to the same problems in the 30000-batch" % n_runs)
