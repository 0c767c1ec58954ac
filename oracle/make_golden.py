#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only, from
/root/reference/scripts) on seeded inputs.  Run in the build container only; the GPU box has
no /root/reference, the committed .npz files are what travels.

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference diffs

What is stored per case (SURVEY.md section 8c recipe):
  uv [B,n,2] pixels in pattern key order, pattern [n,3], K, R, t, t3, euler (roll,yaw,pitch),
  res_norm, iters (counted by wrapping QEKF_get_hx_H / EKF2_get_hx_H on the instance; the
  linear solvers always run 3), `iters_stable` (same count under the perturbation below), and a per-problem `stable` tag: the reference is re-run with the
  pixels multiplied by (1 +/- 1e-13) and stable = max(|dR|, |dt|/|t3|) < 1e-10.  LM is not
  contractive on ~15 % of inputs; parity is only well-posed where stable.
Inputs follow random_stress_test.py:246-258 draw order from numpy default_rng(seed) and the
reference's own perspective_projection_golden_landmarks (quantised and exact variants).
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/scripts"
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference():
    sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    with contextlib.redirect_stdout(io.StringIO()):
        import PNP_SOLVER_LIB as PNPS  # noqa
        import TEST_TOOLBOX as TTBX  # noqa
    return PNPS, TTBX


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def draw_pose(rng):
    """random_stress_test.py:246-258"""
    roll = rng.uniform(-45.0, 45.0, None)
    pitch = rng.uniform(-45.0, 45.0, None)
    yaw = rng.uniform(-45.0, 45.0, None)
    depth = rng.uniform(20, 225, None) / 100.0
    fov_x = rng.uniform(-45.0, 45.0, None)
    fov_y = rng.uniform(-45.0, 45.0, None)
    t = np.zeros((3, 1))
    t[0, 0] = depth * np.tan(np.deg2rad(fov_x))
    t[1, 0] = depth * np.tan(np.deg2rad(fov_y))
    t[2, 0] = depth
    return roll, pitch, yaw, depth, t


def dict_from_uv(keys, uv):
    return {k: np.array([[uv[i, 0]], [uv[i, 1]], [1.0]]) for i, k in enumerate(keys)}


def call_method(solver, method, pts, pattern_np, key_list):
    """Returns (R, t, t3, roll, yaw, pitch, res_norm, iters)."""
    counter = {"n": 0}
    name = {"qeif": "QEKF_get_hx_H", "lm": "EKF2_get_hx_H", "eif2": "EKF2_get_hx_H"}.get(method)
    if name:
        orig = getattr(solver, name)

        def wrapped(*a, **k):
            counter["n"] += 1
            return orig(*a, **k)
        setattr(solver, name, wrapped)
    try:
        if method == "qeif":
            r = solver.solve_pnp_QEIF_single_pattern(pts, pattern_np, LM_key_list=key_list)
        elif method == "lm":
            r = solver.solve_pnp_LM_single_pattern(pts, pattern_np)
        elif method == "linear_f2":
            r = solver.solve_pnp_formulation_2_single_pattern(pts, pattern_np)
        elif method == "linear_f1":
            r = solver.solve_pnp_single_pattern(pts, pattern_np)
        elif method == "eif2":
            r = solver.solve_pnp_EIF2_single_pattern(pts, pattern_np)
        else:
            raise ValueError(method)
    finally:
        if name:
            delattr(solver, name)
    iters = counter["n"] if name else 3
    R, t, t3, roll, yaw, pitch, res = r
    return (np.array(R), np.array(t).reshape(3), float(t3), float(roll), float(yaw), float(pitch),
            float(res), iters)


def run_case(PNPS, TTBX, tag, method, pattern_dict, key_list, B, seed, quantized, with_stability=True):
    from pnp_solver_test_b200 import patterns as pt
    K = pt.default_camera_matrix()
    solver_gt = quiet(PNPS.PNP_SOLVER, K, [pattern_dict], [1.0], verbose=False)
    solver = quiet(PNPS.PNP_SOLVER, K, [pattern_dict], [1.0], verbose=False)
    keys_all = list(pattern_dict.keys())
    keys = keys_all if key_list is None else list(key_list)
    pattern_np_full = solver.np_point_3d_pretransfer_dict_list[0]
    rng = np.random.default_rng(seed)
    n = len(keys)
    uv = np.zeros((B, n, 2))
    gt = np.zeros((B, 4))
    Rg = np.zeros((B, 3, 3))
    tg = np.zeros((B, 3))
    R = np.zeros((B, 3, 3))
    t = np.zeros((B, 3))
    eul = np.zeros((B, 3))
    res = np.zeros(B)
    iters = np.zeros(B, np.int32)
    stable = np.ones(B, bool)
    iters_stable = np.ones(B, bool)
    sens = np.zeros(B)
    for b in range(B):
        roll, pitch, yaw, depth, tt = draw_pose(rng)
        R_gt = solver_gt.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)
        pts_all = quiet(solver_gt.perspective_projection_golden_landmarks, R_gt, tt, is_quantized=quantized,
                        is_pretrans_points=False, is_returning_homogeneous_vec=True)
        uv[b] = np.array([[pts_all[k][0, 0], pts_all[k][1, 0]] for k in keys])
        gt[b] = (depth, roll, pitch, yaw)
        Rg[b], tg[b] = R_gt, tt.reshape(3)
        # the solver sees exactly the n selected points, in `keys` order
        pts = dict_from_uv(keys, uv[b])
        pat = {k: pattern_np_full[k] for k in keys}
        out = quiet(call_method, solver, method, pts, pat, None)
        R[b], t[b], _, r_, y_, p_, res[b], iters[b] = out
        eul[b] = (r_, y_, p_)
        if with_stability:
            worst = 0.0
            for sgn in (+1.0, -1.0):
                pts2 = dict_from_uv(keys, uv[b] * (1.0 + sgn * 1e-13))
                o2 = quiet(call_method, solver, method, pts2, pat, None)
                d = max(np.abs(o2[0] - R[b]).max(), np.abs(o2[1] - t[b]).max() / abs(t[b, 2]))
                worst = max(worst, d)
                iters_stable[b] &= (o2[7] == iters[b])
            sens[b] = worst
            stable[b] = worst < 1e-10
    P = np.array([pattern_np_full[k].reshape(3) for k in keys])
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), method=method, K=K, pattern=P, uv=uv, gt=gt,
                        R_gt=Rg, t_gt=tg, R=R, t=t, euler=eul, res_norm=res, iters=iters,
                        stable=stable, iters_stable=iters_stable, sens=sens, quantized=quantized, seed=seed)
    return dict(method=method, K=K, pattern=P, uv=uv, R=R, t=t, euler=eul, res_norm=res, iters=iters,
                stable=stable, iters_stable=iters_stable)


def compare_with_oracle(tag, g):
    from oracle import oracle as orc
    o = orc.solve_batch(str(g["method"]), g["uv"], g["pattern"], g["K"])
    st = g["stable"]
    dR = np.abs(o["R"] - g["R"]).reshape(len(st), -1).max(axis=1)
    dt = np.abs(o["t"] - g["t"]).max(axis=1) / np.abs(g["t"][:, 2])
    de = np.abs(o["euler"] - g["euler"]).max(axis=1)
    dres = np.abs(o["res_norm"] - g["res_norm"]) / np.maximum(np.abs(g["res_norm"]), 1e-6)
    it_eq = (o["iters"] == g["iters"]) | ~g["iters_stable"]
    print("%-28s B=%4d stable=%4d | stable: dR %.2e dt %.2e deuler %.2e dres %.2e | iters equal %d/%d | unstable max dR %.2e"
          % (tag, len(st), st.sum(), dR[st].max(), dt[st].max(), de[st].max(), dres[st].max(), it_eq.sum(),
             len(st), dR[~st].max() if (~st).any() else 0.0))


def euler_fixture():
    """GT_R_t_dict.pkl -> plain npz: (roll, yaw, pitch) from the file-name labels under the sign
    rules of m1_result_analysis.py:186-194,:236-238, with the stored np_R_GT / np_t_GT_est."""
    import math
    import joblib
    d = joblib.load(os.path.join(REF, "ground_truth_R_t", "GT_R_t_dict.pkl"))
    names, rpy, Rs, ts, dist = [], [], [], [], []
    for name, v in d.items():
        s = name.split('_')
        # pkl keys are the data file names without their leading token:
        # <left|right>_<roll>_<distance cm>_pitch_<u|d>_<pitch>_yaw_<yaw>_image
        raw_roll, raw_pitch, raw_yaw = float(s[1]), float(s[5]), float(s[7])
        roll = (math.fmod(raw_roll + 180.0, 360.0) - 180.0) * -1.0
        pitch = raw_pitch * (-1.0 if s[4] == 'd' else 1.0)
        yaw = raw_yaw * (-1.0 if s[0] == 'left' else 1.0)
        names.append(name)
        rpy.append((roll, yaw, pitch))
        Rs.append(v["np_R_GT"])
        ts.append(v["np_t_GT_est"].reshape(3))
        dist.append(float(s[2]))
    np.savez_compressed(os.path.join(OUT, "euler_fixture.npz"), roll_yaw_pitch_deg=np.array(rpy),
                        R=np.array(Rs), t=np.array(ts), distance_cm=np.array(dist))
    return len(names)


def solve_pnp_case(PNPS, TTBX, tag, B, seed):
    """solve_pnp() itself with two patterns (arg-min res_norm, PNP_SOLVER_LIB.py:166-199)."""
    from pnp_solver_test_b200 import patterns as pt
    K = pt.default_camera_matrix()
    pats = [pt.get_golden_pattern("Alexander"), pt.get_golden_pattern("Holly")]
    solver_gt = quiet(PNPS.PNP_SOLVER, K, pats, [1.0, 1.0], verbose=False)
    solver = quiet(PNPS.PNP_SOLVER, K, pats, [1.0, 1.0], verbose=False)
    keys = list(pats[0].keys())
    rng = np.random.default_rng(seed)
    uv = np.zeros((B, 15, 2)); R = np.zeros((B, 3, 3)); t = np.zeros((B, 3)); eul = np.zeros((B, 3))
    res = np.zeros(B); best = np.zeros(B, np.int32); gen = np.zeros(B, np.int32)
    for b in range(B):
        roll, pitch, yaw, depth, tt = draw_pose(rng)
        gen[b] = b % 2
        solver_gt.set_golden_pattern_id(int(gen[b]))
        R_gt = solver_gt.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)
        pts = quiet(solver_gt.perspective_projection_golden_landmarks, R_gt, tt, is_quantized=True)
        uv[b] = np.array([[pts[k][0, 0], pts[k][1, 0]] for k in keys])
        Rb, tb, t3, r_, y_, p_, rn = quiet(solver.solve_pnp, pts)
        R[b], t[b], eul[b], res[b] = Rb, np.array(tb).reshape(3), (r_, y_, p_), rn
        best[b] = solver.current_golden_pattern_id
    P = np.array([[solver.np_point_3d_pretransfer_dict_list[i][k].reshape(3) for k in keys] for i in range(2)])
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), K=K, patterns=P, uv=uv, R=R, t=t, euler=eul,
                        res_norm=res, best_pattern=best, generating_pattern=gen,
                        key_index=np.array([keys.index(k) for k in pt.LM_KEY_LIST_6], np.int32))
    from oracle import oracle as orc
    idx = [keys.index(k) for k in pt.LM_KEY_LIST_6]
    o = orc.solve_batch("qeif", uv[:, idx], P[:, idx], K)
    print("%-28s B=%4d | dR %.2e dt %.2e | best_pattern equal %d/%d"
          % (tag, B, np.abs(o["R"] - R).max(), np.abs(o["t"] - t).max(), (o["best_pattern"] == best).sum(), B))

def stress_report_case(PNPS, TTBX, tag, B, seed):
    """The body of random_stress_test.py:226-410 (pose draw, quantised projection, solve_pnp,
    pass flags, compare_result_and_generate_result_dict) plus the statistics block of
    TEST_TOOLBOX.data_analysis_and_saving (:1070-1112), all by the unmodified reference."""
    from pnp_solver_test_b200 import patterns as pt
    K = pt.default_camera_matrix()
    pats = [pt.get_golden_pattern("Alexander")]
    solver_gt = quiet(PNPS.PNP_SOLVER, K, pats, [1.0], verbose=False)
    solver = quiet(PNPS.PNP_SOLVER, K, pats, [1.0], verbose=False)
    cls_param = quiet(TTBX.get_classification_parameters, drpy_class_format="drpy_expand")
    keys = list(pats[0].keys())
    rng = np.random.default_rng(seed)
    uv = np.zeros((B, 15, 2)); gt = np.zeros((B, 4))
    R = np.zeros((B, 3, 3)); t = np.zeros((B, 3)); eul = np.zeros((B, 3)); res = np.zeros(B)
    rep = np.zeros((B, 16)); flags = np.zeros((B, 4), np.int32); midx = np.zeros((B, 3), np.int32)
    depth_class = np.zeros(B, np.int32)
    result_list = []
    for b in range(B):
        roll, pitch, yaw, depth, tt = draw_pose(rng)
        R_gt0 = solver_gt.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)
        solver_gt.set_golden_pattern_id(0)
        pts = quiet(solver_gt.perspective_projection_golden_landmarks, R_gt0, tt, is_quantized=True,
                    is_pretrans_points=False, is_returning_homogeneous_vec=True)
        uv[b] = np.array([[pts[k][0, 0], pts[k][1, 0]] for k in keys])
        gt[b] = (depth, roll, pitch, yaw)
        dist_cm = depth * 100.0
        cd = {c: quiet(TTBX.classify_drpy, cls_param, v, class_name=n_)
              for c, v, n_ in (("distance", dist_cm, "depth"), ("roll", roll, "roll"), ("pitch", pitch, "pitch"), ("yaw", yaw, "yaw"))}
        depth_class[b] = cls_param["labels"]["depth"].index(cd["distance"])
        Rb, tb, t3, r_, y_, p_, rn = quiet(solver.solve_pnp, pts)
        R[b], t[b], eul[b], res[b] = Rb, np.array(tb).reshape(3), (r_, y_, p_), rn
        R_gt = solver.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)   # :365
        distance_GT = dist_cm * 0.01
        t_gt_est = (tb / t3) * distance_GT                                               # :367-368
        pass_list, pass_count = TTBX.check_if_the_sample_passed((t3 * 100.0, r_, p_, y_), (dist_cm, roll, pitch, yaw),
                                                                (10.0, 10.0, 10.0, 10.0))
        flags[b] = [int(bool(x)) for x in pass_list]
        data_idx = dict(idx=b, file_name="random_drpy_%s_%s_%s_%s" % (cd["distance"], cd["roll"], cd["pitch"], cd["yaw"]),   # random_stress_test.py:301
                        distance=dist_cm, roll=roll, pitch=pitch, yaw=yaw, **{"class": cd})
        rd, _, _ = quiet(TTBX.compare_result_and_generate_result_dict, solver, data_idx, Rb, tb, (r_, p_, y_), R_gt, t_gt_est,
                         (roll, pitch, yaw), rn, 4 - pass_count, pass_count, pass_list, np_point_image_dict=pts, verbose=False)
        result_list.append(rd)
        rep[b] = [rd["depth_err"], rd["roll_err"], rd["pitch_err"], rd["yaw_err"],
                  rd["LM_GT_error_average_normalize"], rd["LM_GT_error_max_normalize"],
                  rd["predict_LM_error_average_normalize"], rd["predict_LM_error_max_normalize"],
                  rd["predict_GT_error_average_normalize"], rd["predict_GT_error_max_normalize"],
                  rd["t3_est"], rd["distance_GT"], rd["roll_est"], rd["pitch_est"], rd["yaw_est"], 0.0]
        midx[b] = [keys.index(rd["LM_GT_error_max_key"]) if rd["LM_GT_error_max_key"] else -1,
                   keys.index(rd["predict_LM_error_max_key"]) if rd["predict_LM_error_max_key"] else -1,
                   keys.index(rd["predict_GT_error_max_key"]) if rd["predict_GT_error_max_key"] else -1]
    # statistics, unscaled (unit_scale=1): n, m_ratio, mean, stddev, max_dev, MAE_2_GT, MAE_2_mean
    def stat(lst, ek, gk):
        if len(lst) == 0:
            return [np.nan] * 7
        d = quiet(TTBX.get_statistic_of_result, lst, data_est_key=ek, data_GT_key=gk, unit="u", unit_scale=1.0, verbose=False)
        return [d["n_data"], d["m_ratio"], d["mean(u)"], d["stddev(u)"], d["max_dev(u)"], d["MAE_2_GT(u)"], d["MAE_2_mean(u)"]]
    quant = (("depth", "t3_est", "distance_GT"), ("roll", "roll_est", "roll_GT"), ("pitch", "pitch_est", "pitch_GT"),
             ("yaw", "yaw_est", "yaw_GT"))
    labels = cls_param["labels"]["depth"]
    cdict = quiet(TTBX.get_classified_result, result_list, class_name='distance', approval_func=None)
    stats_all = np.array([stat(result_list, ek, gk) for _, ek, gk in quant], dtype=np.float64)
    stats_depth = np.array([[stat(cdict.get(lb, []), ek, gk) for lb in labels] for _, ek, gk in quant], dtype=np.float64)
    # the same with the approval functions of get_classified_result (:939-970): only approved samples stay in their class
    stats_depth_appr = {}
    for nm, fn in (("small_angle", TTBX.approval_func_small_angle), ("large_angle", TTBX.approval_func_large_angle)):
        cd_a = quiet(TTBX.get_classified_result, result_list, class_name='distance', approval_func=fn)
        stats_depth_appr[nm] = np.array([[stat(cd_a.get(lb, []), ek, gk) for lb in labels] for _, ek, gk in quant], dtype=np.float64)
        assert len(cd_a["all"]) == len(result_list)
    # the writers (TEST_TOOLBOX.py:693-820), unmodified, on the same result_list.  The six ndarray fields are
    # dropped from the dicts first (NumPy's multi-line str() of a matrix in a CSV cell is not reproduced), and the
    # NumPy scalars of this generator are turned into the Python floats the script's own draws would be under NumPy 1.x
    # (NumPy 2 prints np.float64(...) inside the `drpy` tuple).
    import tempfile
    drop = ("np_R_GT", "np_R_est", "np_R_err", "np_t_GT_est", "np_t_est", "np_t_err")
    plain = []
    for rd in result_list:
        d = {k: v for k, v in rd.items() if k not in drop}
        d["drpy"] = tuple(float(x) for x in d["drpy"])
        for k, v in list(d.items()):
            if isinstance(v, (np.floating,)):
                d[k] = float(v)
            elif isinstance(v, (np.bool_,)):
                d[k] = bool(v)
        plain.append(d)
    tmpd = tempfile.mkdtemp()
    quiet(TTBX.write_result_to_csv, plain, os.path.join(tmpd, "r.csv"))
    result_csv = open(os.path.join(tmpd, "r.csv"), newline="").read()
    stat_txt, stat_csv = {}, {}
    for name, ek, gk, unit, sc in (("depth", "t3_est", "distance_GT", "cm", 100.0), ("roll", "roll_est", "roll_GT", "deg.", 1.0),
                                   ("pitch", "pitch_est", "pitch_GT", "deg.", 1.0), ("yaw", "yaw_est", "yaw_GT", "deg.", 1.0)):
        csd = {lb: quiet(TTBX.get_statistic_of_result, lst, class_name="distance", class_label=lb, data_est_key=ek, data_GT_key=gk,
                         unit=unit, unit_scale=sc, verbose=False) for lb, lst in cdict.items()}
        quiet(TTBX.write_statistic_to_txt, csd, os.path.join(tmpd, "s.txt"), class_name="distance", statistic_data_name=name)
        quiet(TTBX.write_statistic_to_csv, csd, os.path.join(tmpd, "s.csv"), class_name="distance", statistic_data_name=name, is_horizontal=True)
        stat_txt[name] = open(os.path.join(tmpd, "s.txt"), newline="").read()
        stat_csv[name] = open(os.path.join(tmpd, "s.csv"), newline="").read()
    # the (distance, roll, pitch, yaw) class-combination tables of data_analysis_and_saving (:1215-1346): the unmodified
    # get_all_class_seperated_result / get_drpy_statistic / write_drpy_2_depth_statistic_CSV on the same result_list
    dcd, d_l, r_l, p_l, y_l = quiet(TTBX.get_all_class_seperated_result, plain)
    drpy_csv = {}
    drpy_q = (("depth", "distance", "t3_est", "distance_GT", "cm", 100.0), ("roll", "roll", "roll_est", "roll_GT", "deg.", 1.0),
              ("pitch", "pitch", "pitch_est", "pitch_GT", "deg.", 1.0), ("yaw", "yaw", "yaw_est", "yaw_GT", "deg.", 1.0),
              ("LM_GT_error_average_normalize", "LM_GT_error_average_normalize", "LM_GT_error_average_normalize", None, "px_m", 1.0))
    for name, cname, ek, gk, unit, sc in drpy_q:
        sd = quiet(TTBX.get_drpy_statistic, dcd, class_name=cname, data_est_key=ek, data_GT_key=gk, unit=unit, unit_scale=sc)
        todo = [("%s_mean(%s)" % (name, unit), "mean(%s)" % unit), ("%s_stddev(%s)" % (name, unit), "stddev(%s)" % unit)]
        if name == "depth":
            todo.append(("all_n_data", "n_data"))
        for fkey, metric in todo:
            quiet(TTBX.write_drpy_2_depth_statistic_CSV, sd, os.path.join(tmpd, "d.csv"), d_l, r_l, p_l, y_l, matric_label=metric)
            drpy_csv[fkey] = open(os.path.join(tmpd, "d.csv"), newline="").read()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), K=K, pattern=pt.pattern_array(pats[0]), uv=uv, gt=gt, R=R, t=t,
                        **{"drpy_csv_" + k: np.array(v) for k, v in drpy_csv.items()},
                        **{"stats_by_depth_" + k: v for k, v in stats_depth_appr.items()},
                        euler=eul, res_norm=res, report=rep, flags=flags, max_idx=midx, depth_class=depth_class,
                        stats_all=stats_all, stats_by_depth=stats_depth, keys=np.array(keys), result_csv=np.array(result_csv),
                        **{"stat_txt_" + k: np.array(v) for k, v in stat_txt.items()},
                        **{"stat_csv_" + k: np.array(v) for k, v in stat_csv.items()},
                        key_index=np.array([keys.index(k) for k in pt.LM_KEY_LIST_6], np.int32))
    from oracle import oracle as orc
    o = orc.report_batch(pt.pattern_array(pats[0]), uv, K, R, t, eul, gt)
    print("%-28s B=%4d | report max abs diff %.2e | flags equal %s | max_idx equal %s | pass rate %.3f"
          % (tag, B, np.abs(o["report"] - rep).max(), (o["flags"] == flags).all(), (o["max_idx"] == midx).all(),
             flags.all(axis=1).mean()))
    sa = np.array([orc.stats_of(rep[:, 10 + 0], rep[:, 11])] + [orc.stats_of(rep[:, 12 + i], gt[:, 1 + i]) for i in range(3)])
    print("%-28s stats_all max rel diff %.2e" % ("", np.abs(sa - stats_all).max()))


def fragility_case(tag, B, seed):
    """The analysis block of face_variation_test.py (:631-759).  The script executes on import, so its
    function get_most_fragile_point_and_perturbation_direction is taken from the file's text at run
    time (lines from its `def` to the next top-level `def`) and exec'd unmodified; the heaps are built
    and popped exactly as :631-653 do.  Inputs: the oracle's perturbed-pattern workload and the
    errors of the (pinned) oracle QEIF solve with the golden pattern."""
    import heapq
    from oracle import oracle as orc
    from pnp_solver_test_b200 import patterns as pt
    src = open(os.path.join(REF, "face_variation_test.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if l.startswith("def get_most_fragile_point_and_perturbation_direction"))
    i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("def "))
    ns = {"np": np}
    exec("\n".join(src[i0:i1]), ns)
    ref_fn = ns["get_most_fragile_point_and_perturbation_direction"]
    pat = pt.get_golden_pattern("Alexander")
    keys = list(pat.keys())
    P, K = pt.pattern_array(pat), pt.default_camera_matrix()
    fixed = keys.index("eye_c_51")
    w = orc.synth_face_variation(0, B, P, K, fixed, 0.02, orc.default_synth(seed=seed))
    idx = [keys.index(k) for k in pt.LM_KEY_LIST_6]
    o = orc.solve_batch("qeif", w["uv"][:, idx], P[idx], K)
    err = np.stack([o["t"][:, 2] - w["gt"][:, 0], o["euler"][:, 0] - w["gt"][:, 1], o["euler"][:, 2] - w["gt"][:, 2],
                    o["euler"][:, 1] - w["gt"][:, 3]], axis=1)      # depth, roll, pitch, yaw
    result_list = [dict(np_pattern_perturbation_dict={k: w["perturb"][b, j].reshape(3, 1) for j, k in enumerate(keys)})
                   for b in range(B)]
    out = {}
    k = int(B * 0.1)
    for q, name in enumerate(("depth", "roll", "pitch", "yaw")):
        heap = []
        for b in range(B):
            heapq.heappush(heap, (-abs(err[b, q]), b))              # :555-560 of the script
        top = []
        for _ in range(k):
            e = list(heapq.heappop(heap)); e[0] *= -1; top.append(e)   # :641-653
        r = quiet(ref_fn, pat, result_list, top)
        out[name + "_count"] = np.array([r["fragile_point_count_dict"][kk] for kk in keys])
        out[name + "_sorted_keys"] = np.array([kk for _, kk in r["fragile_point_sorted_list"]])
        out[name + "_similarity"] = np.array(r["top_similarity_list"])
        out[name + "_directions"] = np.array([[d[kk].reshape(3) for kk in keys] for d in r["top_perturbation_list"]])
        out[name + "_value_max"], out[name + "_value_mean"] = r["value_max"], r["top_value_mean"]
        mine = orc.fragility_of(np.abs(err[:, q]), w["perturb"], keys)
        dsim = np.abs(mine["top_similarity"] - out[name + "_similarity"]).max()
        ddir = max(min(np.abs(mine["top_perturbation"][i] - out[name + "_directions"][i]).max(),
                       np.abs(mine["top_perturbation"][i] + out[name + "_directions"][i]).max()) for i in range(5))
        print("%-28s %-5s counts equal %s | sorted equal %s | d similarity %.2e | d direction %.2e | d mean %.2e"
              % (tag, name, (np.array([mine["fragile_point_count_dict"][kk] for kk in keys]) == out[name + "_count"]).all(),
                 [kk for _, kk in mine["fragile_point_sorted_list"]] == list(out[name + "_sorted_keys"]), dsim, ddir,
                 abs(mine["top_value_mean"] - out[name + "_value_mean"])))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), K=K, pattern=P, keys=np.array(keys), fixed_index=fixed, seed=seed,
                        uv=w["uv"], gt=w["gt"], perturb=w["perturb"], err=err, **out)


def lm_trace_case(PNPS, tag, src_tag):
    """Per-iteration outputs of the reference's LM (PNP_SOLVER_LIB.py:2642-2702) on the inputs of the golden
    case `src_tag`: what solve_pnp_LM_single_pattern would return had its loop stopped after k = 1..14
    iterations.  The loop count is a local literal (:2635), so the unmodified method is run once and
    EKF2_get_hx_H (:3718) is wrapped on the instance: its k-th call receives the state x_{k-1} (copied; the
    loop updates it in place) and returns hx(x_{k-1}), from which res_norm of a k-iteration run follows
    (:2681); R_k, t_k are the reference's own EKF2_reconstruct_R_t_m1 (:3500) of the state of call k + 1
    (k = 14: the method's return value).  The same for pixels * (1 +/- 1e-13): first_div[b] = the first k
    at which either perturbed run differs from the unperturbed one by more than 1e-10 in max(|dR|, |dt|/t3)
    (15 = never) -- up to there a 1e-9 parity statement is well-posed for problem b."""
    from pnp_solver_test_b200 import patterns as pt
    g = np.load(os.path.join(OUT, src_tag + ".npz"))
    K, P, uv = g["K"], g["pattern"], g["uv"]
    B, n = uv.shape[0], uv.shape[1]
    keys = ["p%04d" % i for i in range(n)]
    pat = {k: P[i].reshape(3, 1).copy() for i, k in enumerate(keys)}
    solver = quiet(PNPS.PNP_SOLVER, K, [{k: P[i].tolist() for i, k in enumerate(keys)}], [1.0], verbose=False)

    def run(uv_b):
        calls = []
        orig = solver.EKF2_get_hx_H

        def wrapped(x, B_x, B_y, co_P, *a, **kw):
            hx, J = orig(x, B_x, B_y, co_P, *a, **kw)
            z = np.vstack([B_x, B_y])
            calls.append((np.array(x, dtype=np.float64).copy(), float(np.linalg.norm(z - hx[:2 * n]))))
            return hx, J
        solver.EKF2_get_hx_H = wrapped
        try:
            r = quiet(solver.solve_pnp_LM_single_pattern, dict_from_uv(keys, uv_b), pat)
        finally:
            del solver.EKF2_get_hx_H
        assert len(calls) == 14
        Rk, tk, resk = np.zeros((14, 3, 3)), np.zeros((14, 3)), np.zeros(14)
        for k in range(1, 15):
            if k < 14:
                Rr, tr, _ = quiet(solver.EKF2_reconstruct_R_t_m1, calls[k][0].copy())
            else:
                Rr, tr = r[0], r[1]
            Rk[k - 1], tk[k - 1], resk[k - 1] = np.array(Rr), np.array(tr).reshape(3), calls[k - 1][1]
        assert resk[13] == float(r[6])
        return Rk, tk, resk

    R_k, t_k, res_k = np.zeros((B, 14, 3, 3)), np.zeros((B, 14, 3)), np.zeros((B, 14))
    sens_k = np.zeros((B, 14))
    for b in range(B):
        R_k[b], t_k[b], res_k[b] = run(uv[b])
        for sgn in (+1.0, -1.0):
            R2, t2, _ = run(uv[b] * (1.0 + sgn * 1e-13))
            d = np.maximum(np.abs(R2 - R_k[b]).reshape(14, -1).max(axis=1), np.abs(t2 - t_k[b]).max(axis=1) / np.abs(t_k[b][:, 2]))
            sens_k[b] = np.maximum(sens_k[b], np.nan_to_num(d, nan=np.inf))
    bad = sens_k > 1e-10
    first_div = np.where(bad.any(axis=1), bad.argmax(axis=1) + 1, 15).astype(np.int32)
    assert np.abs(R_k[:, 13] - g["R"]).max() == 0.0 and np.abs(t_k[:, 13] - g["t"]).max() == 0.0
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), source=src_tag, K=K, pattern=P, uv=uv, R_k=R_k, t_k=t_k, res_k=res_k,
                        sens_k=sens_k, first_div=first_div, stable=g["stable"])
    from oracle import oracle as orc
    worst = 0.0
    for k in range(1, 15):
        o = orc.solve_batch("lm", uv, P, K, params=orc.default_params(max_it=k))
        m = first_div > k
        d = np.maximum(np.abs(o["R"] - R_k[:, k - 1]).reshape(B, -1).max(axis=1), np.abs(o["t"] - t_k[:, k - 1]).max(axis=1) / np.abs(t_k[:, k - 1, 2]))
        worst = max(worst, d[m].max() if m.any() else 0.0)
    print("%-28s B=%4d | first_div histogram (k=1..15): %s | oracle vs reference over all k < first_div: %.2e"
          % (tag, B, np.bincount(first_div, minlength=16)[1:].tolist(), worst))


def homogeneous_case(PNPS, tag, B, seed):
    """Image points whose homogeneous coordinate is not 1: f2_get_B_xy multiplies the (3,1) vector it is given by K^-1
    (PNP_SOLVER_LIB.py:3305-3307) and perspective_projection emits (x, y, z) / |z|, i.e. w = -1, behind the camera (:4548).
    The stress workload's points with two of the fifteen vectors negated (u, v, 1) -> (-u, -v, -1) -- the same ray -- and
    one scaled by 2, solved by the unmodified reference (solve_pnp = QEIF on the 6-key subset, and linear F2 on all)."""
    from pnp_solver_test_b200 import patterns as pt
    K = pt.default_camera_matrix()
    pat = pt.get_golden_pattern("Alexander")
    solver_gt = quiet(PNPS.PNP_SOLVER, K, [pat], [1.0], verbose=False)
    solver = quiet(PNPS.PNP_SOLVER, K, [pat], [1.0], verbose=False)
    keys = list(pat.keys())
    rng = np.random.default_rng(seed)
    uvw = np.zeros((B, 15, 3))
    out = {m: dict(R=np.zeros((B, 3, 3)), t=np.zeros((B, 3)), euler=np.zeros((B, 3)), res_norm=np.zeros(B)) for m in ("qeif6", "linear_f2")}
    flip = [keys.index("eye_l_96"), keys.index("chin_t_16")]
    for b in range(B):
        roll, pitch, yaw, depth, tt = draw_pose(rng)
        R_gt = solver_gt.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)
        pts = quiet(solver_gt.perspective_projection_golden_landmarks, R_gt, tt, is_quantized=True)
        for j in flip:
            pts[keys[j]] = -pts[keys[j]]
        pts[keys[3]] = 2.0 * pts[keys[3]]
        uvw[b] = np.array([pts[k].reshape(3) for k in keys])
        for m, fn in (("qeif6", lambda: solver.solve_pnp(pts)),
                      ("linear_f2", lambda: solver.solve_pnp_formulation_2_single_pattern(pts, solver.np_point_3d_pretransfer_dict_list[0]))):
            Rb, tb, t3, r_, y_, p_, rn = quiet(fn)
            out[m]["R"][b], out[m]["t"][b], out[m]["euler"][b], out[m]["res_norm"][b] = Rb, np.array(tb).reshape(3), (r_, y_, p_), rn
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), K=K, pattern=pt.pattern_array(pat), uvw=uvw,
                        key_index=np.array([keys.index(k) for k in pt.LM_KEY_LIST_6], np.int32),
                        **{"%s_%s" % (m, k): v for m, d in out.items() for k, v in d.items()})
    print("%-28s B=%4d | points with w != 1 per problem: %d" % (tag, B, int((uvw[0, :, 2] != 1).sum())))


def main():
    os.makedirs(OUT, exist_ok=True)
    PNPS, TTBX = load_reference()
    from pnp_solver_test_b200 import patterns as pt
    alex = pt.get_golden_pattern("Alexander")
    p68, p1024 = pt.synthetic_pattern(68), pt.synthetic_pattern(1024)
    small = "--small" in sys.argv
    B = 32 if small else 192
    cases = []
    for q, qn in ((True, "q"), (False, "x")):
        cases += [
            ("qeif_n6_%s" % qn, "qeif", alex, pt.LM_KEY_LIST_6, B, 42, q),
            ("qeif_n15_%s" % qn, "qeif", alex, None, B, 43, q),
            ("qeif_n68_%s" % qn, "qeif", p68, None, B // 2, 44, q),
            ("lm_n15_%s" % qn, "lm", alex, None, B, 45, q),
            ("lm_n68_%s" % qn, "lm", p68, None, B, 46, q),
            ("linear_f2_n15_%s" % qn, "linear_f2", alex, None, B, 47, q),
            ("linear_f2_n68_%s" % qn, "linear_f2", p68, None, B // 2, 48, q),
            ("linear_f1_n15_%s" % qn, "linear_f1", alex, None, B, 49, q),
            ("linear_f1_n68_%s" % qn, "linear_f1", p68, None, B // 2, 50, q),
            ("eif2_n15_%s" % qn, "eif2", alex, None, B, 55, q),
            ("eif2_n68_%s" % qn, "eif2", p68, None, B // 2, 56, q),
        ]
    cases += [
        ("qeif_n1024_q", "qeif", p1024, None, 3, 51, True),
        ("lm_n1024_q", "lm", p1024, None, 12, 52, True),
        ("linear_f2_n1024_q", "linear_f2", p1024, None, 12, 53, True),
        ("linear_f1_n1024_q", "linear_f1", p1024, None, 12, 54, True),
        ("eif2_n1024_q", "eif2", p1024, None, 6, 57, True),
    ]
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for tag, method, pat, keys, b, seed, q in cases:
        if only and not any(o in tag for o in only):
            continue
        g = run_case(PNPS, TTBX, tag, method, pat, keys, b, seed, q)
        compare_with_oracle(tag, g)
    if not only or "solve_pnp" in only:
        solve_pnp_case(PNPS, TTBX, "solve_pnp_two_patterns", 96, 60)
    if not only or "stress_report" in only:
        stress_report_case(PNPS, TTBX, "stress_report", 256, 61)
    if not only or "fragility" in only:
        fragility_case("fragility", 600, 62)
    if not only or "lm_trace" in only:
        for src in ("lm_n15_q", "lm_n15_x", "lm_n68_q", "lm_n68_x"):
            lm_trace_case(PNPS, src.replace("lm_", "lmtrace_"), src)
    if not only or "homogeneous" in only:
        homogeneous_case(PNPS, "homogeneous_n15", 48, 63)
    if not only or "euler" in only:
        print("euler fixture:", euler_fixture(), "vectors")


if __name__ == "__main__":
    main()
