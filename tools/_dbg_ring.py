import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import workload as wl, patterns as pt
n, B = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
P = pt.pattern_array(pt.synthetic_pattern(n)); K = pt.default_camera_matrix()
w = wl.synth_batch(0, B, P, K)
patd = torch.from_numpy(P).cuda()[None]
for want in (("R", "t", "iters"), ("R", "t", "euler", "res_norm", "iters", "best_pattern")):
    for rep in range(3):
        o = pnp.solve_batch("lm", w["uv"], patd, K, want=want)
        torch.cuda.synchronize()
        print("ok", want[-1], rep, float(o["R"][:, 0, 0].nan_to_num().sum()), flush=True)
