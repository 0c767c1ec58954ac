#!/usr/bin/env python
"""Developer tool: does running the FP64-bound iteration kernel of one part of the batch beside the streaming
kernels of another part pay?  The headline step (solve + report, 1 Mi x 68 points, FP64 LM) as ONE call against
the same batch cut into C parts issued round-robin on S streams."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import workload as wl, patterns as pt

B, n = 1 << 20, 68
K = pt.default_camera_matrix()
P = pt.pattern_array(pt.synthetic_pattern(n))
w = wl.synth_batch(0, B, P, K)
patd = torch.from_numpy(P).cuda()[None]
prm = pnp.default_params()


def run(parts, streams, reps=20):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    cut = B // parts
    uv = [w["uv"][:, i * cut:(i + 1) * cut] if w["uv"].shape[-1] == B else w["uv"][i * cut:(i + 1) * cut] for i in range(parts)]
    gt = [w["gt"][i * cut:(i + 1) * cut] for i in range(parts)]
    def once():
        cur = torch.cuda.current_stream()
        for s in ss:
            s.wait_stream(cur)
        for i in range(parts):
            with torch.cuda.stream(ss[i % streams]):
                wl.solve_report_batch("lm", uv[i], patd, K, gt[i], params=prm)
        for s in ss:
            cur.wait_stream(s)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("uv layout", tuple(w["uv"].shape), "gt", tuple(w["gt"].shape))
for parts, streams in ((1, 1), (2, 2), (4, 2), (8, 2), (4, 4), (8, 4), (16, 4), (3, 3), (6, 3)):
    print("parts %2d streams %d: %.4f ms per 1 Mi solves + reports" % (parts, streams, run(parts, streams)), flush=True)
