"""ctypes front-end of the CPU parity oracle (oracle/pnp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  See the header of pnp_oracle.c for what
the oracle restates (reference file:line citations live next to each C function).

Also holds `stats_of`, the NumPy restatement of TEST_TOOLBOX.get_statistic_of_result
(TEST_TOOLBOX.py:892-937), which is easier to state in NumPy than in C.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpnp_oracle.so")

METHOD_QEIF, METHOD_LM, METHOD_LINEAR_F2, METHOD_LINEAR_F1 = 0, 1, 2, 3
METHODS = {"qeif": 0, "lm": 1, "linear_f2": 2, "linear_f1": 3, "lm_plus": 4, "eif2": 5}
TRACE_DIM = {0: 6, 1: 12, 2: 11, 3: 11, 4: 12, 5: 12}


class Params(C.Structure):
    """Mirror of oracle_params_t (defaults = the reference's inline constants)."""
    _fields_ = [("max_it", C.c_int32), ("linear_it", C.c_int32), ("lm_lambda", C.c_double),
                ("exit_tol", C.c_double), ("f_weight", C.c_double), ("meas_sigma_px", C.c_double),
                ("proc_q", C.c_double), ("proc_d", C.c_double), ("omega0", C.c_double),
                ("res_old0", C.c_double)]


class Synth(C.Structure):
    """Mirror of oracle_synth_t (defaults = random_stress_test.py:246-258)."""
    _fields_ = [("seed", C.c_uint64), ("angle_range_deg", C.c_double), ("depth_min_m", C.c_double),
                ("depth_max_m", C.c_double), ("fov_max_deg", C.c_double), ("is_quantized", C.c_int32),
                ("quantize_q", C.c_double), ("noise_sigma_px", C.c_double), ("roll_center_deg", C.c_double),
                ("pitch_center_deg", C.c_double), ("yaw_center_deg", C.c_double)]


def build(force=False):
    """Compile libpnp_oracle.so in place (building the checker is not using it)."""
    src = os.path.join(_HERE, "pnp_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else os.environ.get("CC", "gcc")
    subprocess.check_call([cc, "-O2", "-fPIC", "-pthread", "-fvisibility=hidden", "-std=c99",
                           "-D_GNU_SOURCE", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.pnp_oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_params(**kw):
    p = Params()
    lib().pnp_oracle_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_synth(seed=42, is_quantized=True, quantize_q=1.0, noise_sigma_px=0.0, **kw):
    s = Synth(seed, 45.0, 0.20, 2.25, 45.0, int(is_quantized), quantize_q, noise_sigma_px, 0.0, 0.0, 0.0)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def num_threads():
    return int(lib().pnp_oracle_num_threads())


def solve_one(method, P, uv, K, params=None, want_trace=False):
    """One problem, one pattern.  Returns dict(R, t, euler, res_norm, iters[, trace])."""
    m = METHODS[method] if isinstance(method, str) else method
    P, uv, K = _f64(P), _f64(uv), _f64(K)
    n = P.shape[0]
    params = params or default_params()
    R, t, e = np.zeros(9), np.zeros(3), np.zeros(3)
    res = C.c_double(0.0)
    nit = max(params.max_it, params.linear_it)
    trace = np.full((nit, TRACE_DIM[m]), np.nan) if want_trace else None
    fn = [lib().pnp_oracle_qeif, lib().pnp_oracle_lm, lib().pnp_oracle_linear_f2,
          lib().pnp_oracle_linear_f1, lib().pnp_oracle_lm_plus, lib().pnp_oracle_eif2][m]
    fn.restype = C.c_int
    it = fn(C.c_int(n), _p(P), _p(uv), _p(K), C.byref(params), _p(R), _p(t), _p(e), C.byref(res),
            _p(trace) if want_trace else None)
    out = dict(R=R.reshape(3, 3), t=t, euler=e, res_norm=res.value, iters=int(it))
    if want_trace:
        out["trace"] = trace[:it]
    return out


def solve_batch(method, uv, patterns, K, params=None, n_threads=0):
    """uv [B,n,2]; patterns [P,n,3] (or [n,3]).  Returns dict of arrays."""
    m = METHODS[method] if isinstance(method, str) else method
    uv, patterns, K = _f64(uv), _f64(patterns), _f64(K)
    if patterns.ndim == 2:
        patterns = patterns[None]
    B, n = uv.shape[0], uv.shape[1]
    assert patterns.shape[1] == n
    params = params or default_params()
    R, t, e = np.zeros((B, 3, 3)), np.zeros((B, 3)), np.zeros((B, 3))
    res = np.zeros(B)
    iters, best = np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().pnp_oracle_solve_batch.restype = C.c_int
    rc = lib().pnp_oracle_solve_batch(C.c_int(m), C.c_int64(B), C.c_int(n), _p(uv), _p(patterns),
                                      C.c_int(patterns.shape[0]), _p(K), C.byref(params), _p(R), _p(t),
                                      _p(e), _p(res), _p(iters), _p(best), C.c_int(n_threads))
    assert rc == 0
    return dict(R=R, t=t, euler=e, res_norm=res, iters=iters, best_pattern=best)


def R_from_euler(roll, yaw, pitch, is_degree=False):
    R = np.zeros(9)
    lib().pnp_oracle_R_from_euler(C.c_double(roll), C.c_double(yaw), C.c_double(pitch),
                                  C.c_int(int(is_degree)), _p(R))
    return R.reshape(3, 3)


def euler_from_R(R, is_degree=False):
    R = _f64(R).reshape(9)
    out = np.zeros(3)
    lib().pnp_oracle_euler_from_R(_p(R), C.c_int(int(is_degree)), _p(out))
    return tuple(out)  # (roll, yaw, pitch)


def project(P, K, R, t, is_quantized=False, q=1.0):
    P, K, R, t = _f64(P), _f64(K), _f64(R).reshape(9), _f64(t).reshape(3)
    out = np.zeros((P.shape[0], 3))
    lib().pnp_oracle_project(C.c_int(P.shape[0]), _p(P), _p(K), _p(R), _p(t), C.c_int(int(is_quantized)),
                             C.c_double(q), _p(out))
    return out


def synth(b0, B, P, K, cfg=None, n_threads=0):
    """Counter-based synthetic workload, problems b0..b0+B-1 of the global stream."""
    P, K = _f64(P), _f64(K)
    n = P.shape[0]
    cfg = cfg or default_synth()
    uv, gt = np.zeros((B, n, 2)), np.zeros((B, 4))
    Rg, tg = np.zeros((B, 3, 3)), np.zeros((B, 3))
    lib().pnp_oracle_synth(C.c_int64(b0), C.c_int64(B), C.c_int(n), _p(P), _p(K), C.byref(cfg), _p(uv),
                           _p(gt), _p(Rg), _p(tg), C.c_int(n_threads))
    return dict(uv=uv, gt=gt, R_gt=Rg, t_gt=tg)


def report_batch(P, uv, K, R_est, t_est, euler_est, gt, bounds=(10.0, 10.0, 10.0, 10.0), n_threads=0):
    P, uv, K = _f64(P), _f64(uv), _f64(K)
    R_est, t_est, euler_est, gt = _f64(R_est), _f64(t_est), _f64(euler_est), _f64(gt)
    B, n = uv.shape[0], uv.shape[1]
    bounds = _f64(bounds)
    rep = np.zeros((B, 16))
    flags, midx = np.zeros((B, 4), np.int32), np.zeros((B, 3), np.int32)
    lib().pnp_oracle_report_batch(C.c_int64(B), C.c_int(n), _p(P), _p(uv), _p(K), _p(R_est), _p(t_est),
                                  _p(euler_est), _p(gt), _p(bounds), _p(rep), _p(flags), _p(midx),
                                  C.c_int(n_threads))
    return dict(report=rep, flags=flags, max_idx=midx)


def drpy_stats(report, gt, bins):
    """get_all_class_seperated_result + get_drpy_statistic (TEST_TOOLBOX.py:975-1066) on arrays:
    stats_of per (distance, roll, pitch, yaw) class combination for depth, roll, pitch, yaw and
    LM_GT_error_average_normalize.  bins: four ascending bin-edge lists (distance in cm).
    Returns {name: [nd, nr, np, ny, 7]} with zeros where the combination holds no data."""
    report, gt = np.asarray(report, np.float64), np.asarray(gt, np.float64)
    cls = [np.digitize(gt[:, q] * (100.0 if q == 0 else 1.0), np.asarray(bins[q], np.float64)) for q in range(4)]
    shape = tuple(len(b) + 1 for b in bins)
    pairs = dict(depth=(report[:, 10], report[:, 11]), roll=(report[:, 12], gt[:, 1]), pitch=(report[:, 13], gt[:, 2]),
                 yaw=(report[:, 14], gt[:, 3]), LM_GT_error_average_normalize=(report[:, 4], None))
    out = {k: np.zeros(shape + (7,)) for k in pairs}
    flat = np.ravel_multi_index(cls, shape)
    for c in np.unique(flat):
        m = flat == c
        idx = np.unravel_index(c, shape)
        for k, (e, g) in pairs.items():
            out[k][idx] = stats_of(e[m], None if g is None else g[m])
    return out


def stats_of(est, gt=None):
    """TEST_TOOLBOX.get_statistic_of_result (TEST_TOOLBOX.py:892-937) on plain vectors.

    Returns (n, m_ratio, mean, stddev, max_dev, MAE_2_GT, MAE_2_mean), population variance."""
    est = np.asarray(est, np.float64).reshape(-1, 1)
    if est.shape[0] == 0:
        return None
    if gt is not None:
        gt = np.asarray(gt, np.float64).reshape(-1, 1)
        ratio, err = est / gt, est - gt
    else:
        ratio, err = est, est
    n = err.shape[0]
    ratio_mean = np.average(ratio)
    mean = np.average(err)
    var = (np.linalg.norm(err - mean, ord=2) ** 2) / n
    std = var ** 0.5
    mae_gt = np.linalg.norm(err, ord=1) / n
    mae_mean = np.linalg.norm(err - mean, ord=1) / n
    max_dev = np.linalg.norm(err - mean, ord=np.inf)
    return (n, ratio_mean, mean, std, max_dev, mae_gt, mae_mean)


def synth_face_variation(b0, B, P, K, fixed_index, radius_m=0.02, cfg=None, n_threads=0):
    """face_variation_test.py:296-356 workload: pixels of a randomly perturbed pattern."""
    P, K = _f64(P), _f64(K)
    n = P.shape[0]
    cfg = cfg or default_synth()
    uv, gt, pert = np.zeros((B, n, 2)), np.zeros((B, 4)), np.zeros((B, n, 3))
    lib().pnp_oracle_synth_face_variation(C.c_int64(b0), C.c_int64(B), C.c_int(n), _p(P), _p(K), C.byref(cfg),
                                          C.c_double(radius_m), C.c_int(fixed_index), _p(uv), _p(gt), None, None,
                                          _p(pert), C.c_int(n_threads))
    return dict(uv=uv, gt=gt, perturb=pert)


def fragility_of(abs_err, perturb, keys, ratio=0.1, k_top_direction=5):
    """NumPy restatement of the analysis block of face_variation_test.py for ONE error quantity:
    the heap selection (:631-653: pop int(len * ratio) items off a heap of (-|err|, idx), i.e. largest
    |err| first, smaller idx on ties) and get_most_fragile_point_and_perturbation_direction (:658-728).
    abs_err [B]; perturb [B,n,3] in key order.  Returns the script's result_dict with arrays."""
    abs_err = np.asarray(abs_err, np.float64)
    perturb = np.asarray(perturb, np.float64)
    B, n = perturb.shape[0], perturb.shape[1]
    k = int(B * ratio)                                               # :632
    order = sorted(range(B), key=lambda i: (-abs_err[i], i))[:k]     # heappop order (:641-646)
    count = {key: 0 for key in keys}                                 # :661-664
    rows, vsum = [], 0.0
    for i in order:
        vsum += abs_err[i]
        norms = np.linalg.norm(perturb[i], axis=1)                   # :672
        nmax, kmax = -1.0, None
        for j, key in enumerate(keys):                               # strict '>', first wins (:674-676)
            if norms[j] > nmax:
                nmax, kmax = norms[j], key
        count[kmax] += 1
        rows.append(perturb[i].reshape(-1))                          # vstack of the (3,1) blocks (:681)
    M = np.array(rows)                                               # m x 3n (:684)
    sorted_list = sorted([(count[key], key) for key in keys], reverse=True)    # :694-695
    u, s, vh = np.linalg.svd(M)                                      # :699
    return dict(fragile_point_count_dict=count, fragile_point_sorted_list=sorted_list,
                top_perturbation=vh[:k_top_direction].reshape(k_top_direction, n, 3),
                top_similarity=s[:k_top_direction], value_max=abs_err[order[0]],
                top_value_mean=vsum / float(len(order)), selected=np.array(order))
