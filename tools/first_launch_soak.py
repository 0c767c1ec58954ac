#!/usr/bin/env python
"""Developer tool: the n = 1024 ring kernel as the FIRST launch of fresh processes (the situation in which ring shapes
other than the shipped one faulted on about half of the processes, profiles/r02u_ring_fault_bisect.md).

    python tools/first_launch_soak.py <processes>       # parent: starts the children one after another, counts verdicts
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import pnp_solver_test_b200 as pnp
    from pnp_solver_test_b200 import workload as wl, patterns as pt
    B, method = int(sys.argv[2]), sys.argv[3]
    P = pt.pattern_array(pt.synthetic_pattern(1024)); K = pt.default_camera_matrix()
    w = wl.synth_batch(0, B, P, K)
    patd = torch.from_numpy(P).cuda()[None]
    torch.cuda.synchronize()
    o = pnp.solve_batch(method, w["uv"], patd, K)           # first library call of the process
    torch.cuda.synchronize()
    print("ok %.9e" % float(o["R"].nan_to_num(0.0).sum()))
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bad, sums = 0, {}
for i in range(n):
    B, method = (300, 5000, 20000)[i % 3], ("lm", "linear_f2")[(i // 3) % 2]
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(B), method], capture_output=True, text=True, timeout=300)
    line = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
    if r.returncode != 0 or not line.startswith("ok"):
        bad += 1
        print("process %d (B=%d, %s) FAILED: %s" % (i, B, method, (r.stderr.strip().splitlines() or [line])[-1][:160]), flush=True)
    else:
        sums.setdefault((B, method), set()).add(line)
print("first-launch soak: %d of %d fresh processes failed; results per (B, method) identical across processes: %s"
      % (bad, n, all(len(v) == 1 for v in sums.values())))
