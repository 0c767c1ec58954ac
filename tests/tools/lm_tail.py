import os, sys
_TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(_TESTS)); sys.path.insert(0, _TESTS)
import numpy as np, torch
from gpu_util import cuda_solve
from oracle import oracle as orc
from pnp_solver_test_b200 import patterns as pt
B=32768; n=68
K = pt.default_camera_matrix(); P = pt.pattern_array(pt.synthetic_pattern(n))
w = orc.synth(0, B, P, K, orc.default_synth(seed=777 + n, is_quantized=True))
ref = orc.solve_batch("lm", w["uv"], P, K)
sens = np.zeros(B)
for sgn in (1.0, -1.0):
    o = orc.solve_batch("lm", w["uv"] * (1 + sgn * 1e-13), P, K)
    sens = np.maximum(sens, np.maximum(np.abs(o["R"] - ref["R"]).reshape(B, -1).max(1), np.abs(o["t"] - ref["t"]).max(1) / np.abs(ref["t"][:, 2])))
stable = sens < 1e-10
for mapping, name in ((1, "direct (thread)"), (2, "moment")):
    out = cuda_solve("lm", w["uv"], P, K, mapping=mapping)
    d = np.maximum(np.abs(out["R"] - ref["R"]).reshape(B, -1).max(1), np.abs(out["t"] - ref["t"]).max(1) / np.abs(ref["t"][:, 2]))
    ds = d[stable]
    print(name, "max", ds.max(), "count > 1e-9:", (ds > 1e-9).sum(), "> 1e-10:", (ds > 1e-10).sum(), "median", np.median(ds), "99.9%", np.quantile(ds, 0.999))
    worst = np.argsort(np.where(stable, d, 0))[-5:]
    print("   worst: dev", d[worst], "sens", sens[worst], "ratio dev/sens", d[worst] / sens[worst])
# ratio statistic: deviation relative to the problem's own sensitivity
    r = d[stable] / np.maximum(sens[stable], 1e-16)
    print("   dev / sens: median %.2f  99.9%% %.1f  max %.1f" % (np.median(r), np.quantile(r, 0.999), r.max()))
