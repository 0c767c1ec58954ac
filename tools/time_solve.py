#!/usr/bin/env python
"""Developer tool: per-kernel device times of pnpb200_solve_batch for a method / shape / mapping /
tuning knob, via the library's own CUDA-event profiling (PNPB200_FLAG_PROFILE).

    python tools/time_solve.py --method lm --n 68 --B 1048576 --mapping 0 --tune 2,3,4
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import _lib, workload as wl, patterns as pt

ap = argparse.ArgumentParser()
ap.add_argument("--method", default="lm")
ap.add_argument("--n", type=int, default=68)
ap.add_argument("--B", type=int, default=1 << 20)
ap.add_argument("--mapping", default="0")
ap.add_argument("--tune", default="0")
ap.add_argument("--dtype", default="f64")
ap.add_argument("--subset6", action="store_true")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--noprof", action="store_true", help="no per-kernel events (the sliced pipeline only runs without them)")
args = ap.parse_args()

dt = torch.float64 if args.dtype == "f64" else torch.float32
pat = pt.get_golden_pattern() if args.n <= 15 else pt.synthetic_pattern(args.n)
P = pt.pattern_array(pat)
K = pt.default_camera_matrix()
w = wl.synth_batch(0, args.B, P, K, dtype=dt)
patd = torch.from_numpy(P).cuda().to(dt)[None]
idx = [list(pat).index(k) for k in pt.LM_KEY_LIST_6] if args.subset6 else None
for mapping in [int(x) for x in args.mapping.split(",")]:
  for sl in [0]:
    for tune in [int(x) for x in args.tune.split(",")]:
        prm = pnp.default_params(mapping=mapping, flags=(0 if args.noprof else _lib.FLAG_PROFILE) | (tune << 8))
        for _ in range(3):
            pnp.solve_batch(args.method, w["uv"], patd, K, point_index=idx, params=prm)
        torch.cuda.synchronize()
        _lib.lib.pnpb200_profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            out = pnp.solve_batch(args.method, w["uv"], patd, K, point_index=idx, params=prm)
        e1.record()
        torch.cuda.synchronize()
        ms = (C.c_float * 3)()
        nc = C.c_int()
        _lib.lib.pnpb200_profile_read(ms, C.byref(nc))
        tot = e0.elapsed_time(e1) / args.reps
        print("method=%s n=%d B=%d dtype=%s mapping=%d tune=%d slice=%d: total %.3f ms (%.3e solves/s) kernels [%.3f %.3f %.3f] ms  mean iters %.2f"
              % (args.method, args.n, args.B, args.dtype, mapping, tune, sl, tot, args.B / tot * 1e3, ms[0], ms[1], ms[2],
                 float(out["iters"].double().mean())), flush=True)
