#!/usr/bin/env python
"""Run the five workloads BASELINE.json names (`configs[0..4]`) through the batched CUDA path and
print one JSON line per case.  configs[1] is what bench.py measures; the others are parity-test
shapes, timed here for the record (profiles/).  Device-timed with CUDA events, inputs resident in HBM.

    python workloads/baseline_configs.py [--configs 0,2,3,4] [--scale 1.0]

 0  random_stress_test.py workload: Alexander 15-landmark pattern, QEIF on the 6-key subset
    (what solve_pnp() runs), integer-quantised pixels, + error report + statistics
 2  LM_noise_test.py sweep: yaw in linspace(0,90,15) x noise sigma in linspace(0,5,10) (x 30/112 px),
    fixed pose t=(0,0,1), grid 30/112 px, QEIF-6, FP64 vs FP32: yaw-error mean/std/MAE per cell
 3  large n: 1024-point pattern, 100k problems: linear stage (F2) alone, and LM
 4  64M x 68-point LM, sharded by problem index over the visible ranks (torchrun) or one GPU
 5   face_variation_test.py: perturbed-pattern workload, QEIF-6 solve, error report,
    statistics, top-10 % selection + fragile-point / perturbation-direction analysis, sharded like 4
 6   EIF2 beside LM and QEIF on the stress workload, 15 and 68 landmarks: pass rate and rate
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pnp_solver_test_b200 as pnp
from pnp_solver_test_b200 import patterns as pt, workload as wl


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def config0(scale):
    B = int((1 << 20) * scale)
    K = pt.default_camera_matrix()
    pat = pt.get_golden_pattern("Alexander")
    P = pt.pattern_array(pat)
    solver = pnp.PNP_SOLVER(K, [pat], [1.0])
    w = wl.synth_batch(0, B, P, K)
    ms_solve, out = timed(lambda: solver.solve_pnp_batch(w["uv"]))

    def full():
        o = solver.solve_pnp_batch(w["uv"])
        rep = wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])
        return rep, wl.error_statistics(rep["report"], w["gt"], lazy=True)
    ms_full, (rep, st) = timed(full)
    st = st.result()
    return {"config": 0, "workload": "random_stress_test: %d problems, Alexander 15 pts, QEIF on the 6-key subset, quantised" % B,
            "dtype": "f64", "solve_ms": ms_solve, "solves_per_s": B / ms_solve * 1e3, "solve_report_stats_ms": ms_full,
            "samples_per_s_incl_report": B / ms_full * 1e3, "mean_iters": float(out["iters"].double().mean()),
            "pass_rate_10cm_10deg": float(rep["flags"].all(dim=1).double().mean()),
            "depth_MAE_cm": 100 * float(st["depth"]["all"][5]), "roll_MAE_deg": float(st["roll"]["all"][5]),
            "pitch_MAE_deg": float(st["pitch"]["all"][5]), "yaw_MAE_deg": float(st["yaw"]["all"][5]),
            "reference_note": "reference on 200 samples, seed 42: 2 failed, depth MAE 2.0 cm, roll/pitch/yaw MAE 1.0/1.3/1.6 deg (SURVEY App. C)"}


def config2(scale):
    n_draw = int((1 << 20) * scale)
    K = pt.default_camera_matrix()
    pat = pt.get_golden_pattern("Alexander")
    P = pt.pattern_array(pat)
    idx = [list(pat).index(k) for k in pt.LM_KEY_LIST_6]
    yaws = np.linspace(0.0, 90.0, 15)
    sigmas = np.linspace(0.0, 5.0, 10)
    q = 30.0 / 112.0
    out = {"config": 2, "workload": "LM_noise_test sweep: 15 yaw x 10 sigma cells x %d noise draws, QEIF-6, grid 30/112 px" % n_draw}
    for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        patd = torch.from_numpy(P).cuda().to(dt)[None]
        mean = np.zeros((15, 10)); std = np.zeros((15, 10)); mae = np.zeros((15, 10))
        t_solve = 0.0
        torch.cuda.synchronize()
        for i, yaw in enumerate(yaws):
            for j, sg in enumerate(sigmas):
                cfg = pnp.default_synth(seed=42, angle_range_deg=0.0, yaw_center_deg=float(yaw), depth_min_m=1.0, depth_max_m=1.0,
                                        fov_max_deg=0.0, is_quantized=1, quantize_q=q, noise_sigma_px=float(sg) * q)
                w = wl.synth_batch(0, n_draw, P, K, cfg=cfg, dtype=dt)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = pnp.solve_batch("qeif", w["uv"], patd, K, point_index=idx)
                e1.record()
                err = (o["euler"][:, 1].double() - yaw)              # yaw_est - yaw (LM_noise_test.py:331-333)
                mean[i, j], std[i, j], mae[i, j] = float(err.mean()), float(err.std(unbiased=False)), float(err.abs().mean())
                torch.cuda.synchronize()
                t_solve += e0.elapsed_time(e1)
        out[name] = {"solves_per_s": 150 * n_draw / t_solve * 1e3, "solve_ms_total": t_solve,
                     "yaw_err_mean_deg": mean.round(4).tolist(), "yaw_err_std_deg": std.round(4).tolist(), "yaw_err_MAE_deg": mae.round(4).tolist()}
    d = np.abs(np.array(out["f64"]["yaw_err_MAE_deg"]) - np.array(out["f32"]["yaw_err_MAE_deg"]))
    out["f32_vs_f64_max_abs_diff_of_cell_MAE_deg"] = float(d.max())
    return out


def config3(scale):
    B = int(100000 * scale)
    K = pt.default_camera_matrix()
    P = pt.pattern_array(pt.synthetic_pattern(1024))
    w = wl.synth_batch(0, B, P, K)
    patd = torch.from_numpy(P).cuda()[None]
    res = {"config": 3, "workload": "large n: %d problems x 1024 points, FP64" % B, "bytes_per_solve": 8 * (2 * 1024 + 16) + 8}
    for method in ("linear_f2", "lm", "linear_f1", "qeif"):
        ms, o = timed(lambda: pnp.solve_batch(method, w["uv"], patd, K))
        res[method] = {"ms": ms, "solves_per_s": B / ms * 1e3, "algorithmic_GBps": B * res["bytes_per_solve"] / ms / 1e6}
    return res


def config4(scale):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    total = int((1 << 26) * scale)
    lo, hi = wl.shard_range(total, rank, world)
    B = hi - lo
    K = pt.default_camera_matrix()
    P = pt.pattern_array(pt.synthetic_pattern(68))
    w = wl.synth_batch(lo, B, P, K)
    patd = torch.from_numpy(P).cuda()[None]

    def full():
        o = pnp.solve_batch("lm", w["uv"], patd, K)
        rep = wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])
        return wl.error_statistics(rep["report"], w["gt"], lazy=True)
    ms, st = timed(full, reps=3, warm=1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = st.result()
    return {"config": 4, "workload": "%d x 68-point LM FP64 sharded over %d GPU(s), %d per GPU, + report + all-reduced statistics" % (total, world, B),
            "ms": float(t[0]), "solves_per_s": total / float(t[0]) * 1e3, "n_stat": float(st["depth"]["all"][0]),
            "yaw_MAE_deg": float(st["yaw"]["all"][5]), "hbm_gb_inputs_per_gpu": B * 68 * 16 / 1e9}


def config5(scale):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    total = int((1 << 20) * world * scale)
    lo, hi = wl.shard_range(total, rank, world)
    B = hi - lo
    K = pt.default_camera_matrix()
    pat = pt.get_golden_pattern("Alexander")
    keys = list(pat.keys())
    P = pt.pattern_array(pat)
    solver = pnp.PNP_SOLVER(K, [pat], [1.0])
    fixed = keys.index("eye_c_51")                                  # face_variation_test.py:318

    def full():
        w = wl.synth_face_variation(lo, B, P, K, fixed, 0.02)
        o = solver.solve_pnp_batch(w["uv"])
        rep = wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])
        st = wl.error_statistics(rep["report"], w["gt"], lazy=True)
        ae = rep["report"][:, :4].abs()
        fr = wl.fragility_analysis([ae[:, q] for q in range(4)], w["perturb"], idx0=lo, total=total, keys=keys)
        return st, fr, rep
    ms, (st, fr, rep) = timed(full, reps=5, warm=3)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = st.result()
    return {"config": 5, "workload": "face_variation_test: %d problems over %d GPU(s), Alexander 15 pts perturbed by a 2 cm random direction, "
                                     "QEIF-6 with the golden pattern, report, statistics, top-10 %% fragility analysis" % (total, world),
            "ms_generate_solve_report_stats_analysis": float(t[0]), "samples_per_s": total / float(t[0]) * 1e3,
            "depth_MAE_cm": 100 * float(st["depth"]["all"][5]), "yaw_MAE_deg": float(st["yaw"]["all"][5]),
            "n_selected": [f["n_selected"] for f in fr],
            "most_fragile_depth": fr[0]["fragile_point_sorted_list"][:3], "most_fragile_yaw": fr[3]["fragile_point_sorted_list"][:3],
            "top_similarity_depth": [float(x) for x in fr[0]["top_similarity"]], "value_max_depth_m": fr[0]["value_max"],
            "top_value_mean_depth_m": fr[0]["top_value_mean"]}


def config6(scale):
    B = int((1 << 18) * scale)
    K = pt.default_camera_matrix()
    res = {"config": 6, "workload": "stress workload, %d problems, FP64: EIF2 beside LM and QEIF (all landmarks)" % B}
    for n, pat in ((15, pt.get_golden_pattern("Alexander")), (68, pt.synthetic_pattern(68))):
        P = pt.pattern_array(pat)
        w = wl.synth_batch(0, B, P, K)
        patd = torch.from_numpy(P).cuda()[None]
        for method in ("eif2", "lm", "qeif"):
            ms, o = timed(lambda: pnp.solve_batch(method, w["uv"], patd, K))
            rep = wl.report_batch(P, w["uv"], K, o["R"], o["t"], o["euler"], w["gt"])
            res["%s_n%d" % (method, n)] = {"ms": ms, "solves_per_s": B / ms * 1e3, "mean_iters": float(o["iters"].double().mean()),
                                          "pass_rate_10cm_10deg": float(rep["flags"].all(dim=1).double().mean())}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="0,2,3")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    fns = {0: config0, 2: config2, 3: config3, 4: config4, 5: config5, 6: config6}
    for c in [int(x) for x in args.configs.split(",")]:
        if world > 1 and c not in (4, 5):
            continue
        r = fns[c](args.scale)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
