#!/usr/bin/env python
"""Developer tool: throughput of the host-side pixel packer (pnpb200_pack_i16) by thread count, and the
end-to-end host pipeline with and without the packed transfer."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import _lib, patterns as pt, workload as wl

print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
a = torch.from_numpy(np.random.default_rng(0).integers(0, 1500, 131072 * 136).astype(np.float64)).pin_memory()
dst = torch.empty(a.shape, dtype=torch.int16).pin_memory()
for th in (1, 2, 4, 8, 16, 32):
    t = []
    for _ in range(5):
        t0 = time.perf_counter()
        rc = _lib.lib.pnpb200_pack_i16(0, C.c_void_p(a.data_ptr()), C.c_int64(a.numel()), C.c_void_p(dst.data_ptr()), th)
        t.append(time.perf_counter() - t0)
    print("pack threads %2d exact %d  %.2f ms  %.1f GB/s in" % (th, rc, min(t) * 1e3, a.numel() * 8 / min(t) / 1e9))

if torch.cuda.is_available():
    B, n = 1 << 20, 68
    K = pt.default_camera_matrix()
    P = pt.pattern_array(pt.synthetic_pattern(n))
    w = wl.synth_batch(0, B, P, K)
    host_uv = torch.empty((B, n, 2), dtype=torch.float64).pin_memory()
    host_uv.copy_(w["uv"])
    host_pat = torch.from_numpy(P)[None].contiguous()
    outs = {"R": torch.empty((B, 3, 3), dtype=torch.float64).pin_memory(), "t": torch.empty((B, 3), dtype=torch.float64).pin_memory(),
            "euler": torch.empty((B, 3), dtype=torch.float64).pin_memory(), "res_norm": torch.empty((B,), dtype=torch.float64).pin_memory(),
            "iters": torch.empty((B,), dtype=torch.int32).pin_memory(), "best_pattern": torch.empty((B,), dtype=torch.int32).pin_memory()}
    for chunk in (1 << 17, 1 << 16):
        for th in (0, 6, 8, 12, 15):
            pipe = pnp.HostPipeline(torch.float64, chunk_problems=chunk, n_total=n, n_patterns=1, n_streams=3, pack_threads=th)
            for _ in range(2):
                pipe.solve("lm", host_uv, host_pat, K, outs)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                pipe.solve("lm", host_uv, host_pat, K, outs)
            dt = (time.perf_counter() - t0) / 5
            print("chunk %6d pack_threads %2d: %.2f ms  %.3e solves/s  packed %s" % (chunk, th, dt * 1e3, B / dt, pipe.last_packed()))
            pipe.close()
