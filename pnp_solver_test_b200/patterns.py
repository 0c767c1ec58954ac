"""Golden landmark patterns (fixture data) and the deterministic synthetic extensions.

`get_golden_pattern` mirrors TEST_TOOLBOX.get_golden_pattern (TEST_TOOLBOX.py:67-128): two
15-landmark face models in metres, keyed by landmark name; dict insertion order is the row
order the solvers use (PNP_SOLVER_LIB.py:3269).  The reference has no 68- or 1024-point
pattern; `synthetic_pattern` builds the ones BASELINE.json's configs name, as specified in
SURVEY.md Appendix D (the 15 Alexander points first, then uniform points in the Alexander
bounding box from a seeded NumPy generator), so the 6-key subset stays addressable.
"""
import numpy as np

# The landmark subset hard-wired in PNP_SOLVER.solve_pnp (PNP_SOLVER_LIB.py:156)
LM_KEY_LIST_6 = ['eye_l_96', 'eye_r_97', 'mouse_l_76', 'mouse_r_82', 'nose_t_54', 'chin_t_16']

_COMMON_TAIL = (
    ("brow_cl_35", [0.035, -0.0228, 0.0]),
    ("brow_il_37", [0.0135, -0.017, 0.0]),
    ("brow_ir_42", [-0.0135, -0.017, 0.0]),
    ("brow_cr_44", [-0.035, -0.0228, 0.0]),
)

_ALEXANDER = (
    ("eye_l_96", [0.032, 0.0, 0.0]),
    ("eye_r_97", [-0.032, 0.0, 0.0]),
    ("eye_c_51", [0.0, 0.0, -0.015]),
    ("mouse_l_76", [0.027, 0.070, 0.0]),
    ("mouse_r_82", [-0.027, 0.070, 0.0]),
    ("nose_t_54", [-0.005, 0.0455, -0.03]),
    ("chin_t_16", [0.0, 0.12, 0.0]),
) + _COMMON_TAIL + (
    ("eye_lo_60", [0.047, 0.0, 0.005]),
    ("eye_li_64", [0.018, 0.0, 0.00]),
    ("eye_ro_72", [-0.047, 0.0, 0.005]),
    ("eye_ri_68", [-0.018, 0.0, 0.00]),
)

_HOLLY = (
    ("eye_l_96", [0.028, 0.0, 0.0]),
    ("eye_r_97", [-0.028, 0.0, 0.0]),
    ("eye_c_51", [0.0, 0.0, 0.0]),
    ("mouse_l_76", [0.025, 0.060, 0.0]),
    ("mouse_r_82", [-0.025, 0.060, 0.0]),
    ("nose_t_54", [0.00, 0.039, -0.03]),
    ("chin_t_16", [0.0, 0.098, 0.0]),
) + _COMMON_TAIL + (
    ("eye_lo_60", [0.046, 0.0, 0.0]),
    ("eye_li_64", [0.01, 0.0, 0.0]),
    ("eye_ro_72", [-0.046, 0.0, 0.0]),
    ("eye_ri_68", [-0.01, 0.0, 0.0]),
)


def get_golden_pattern(pattern_name="Alexander"):
    """{landmark key: [x, y, z]} in metres; 'Holly' or (default) 'Alexander'."""
    src = _HOLLY if pattern_name == 'Holly' else _ALEXANDER
    return {k: list(v) for k, v in src}


def synthetic_pattern(n_point):
    """the 15 Alexander points, then n_point-15 uniform points in the
    Alexander bounding box drawn one point at a time from default_rng(n_point)."""
    pat = get_golden_pattern("Alexander")
    if n_point < len(pat):
        raise ValueError("synthetic_pattern needs n_point >= 15")
    pts = np.array(list(pat.values()), dtype=np.float64)
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    rng = np.random.default_rng(n_point)
    for i in range(n_point - len(pat)):
        pat["syn_%04d" % i] = [float(v) for v in rng.uniform(lo, hi)]
    return pat


def pattern_array(point_3d_dict, key_list=None, scale=1.0):
    """Rows of theta_i^T in key order -> (n, 3) float64 (f2_get_P, PNP_SOLVER_LIB.py:3260-3272)."""
    keys = list(point_3d_dict.keys()) if key_list is None else list(key_list)
    return np.array([point_3d_dict[k] for k in keys], dtype=np.float64).reshape(len(keys), 3) * scale


def default_camera_matrix():
    """K used by all three workload scripts (random_stress_test.py:103-110)."""
    f = 225.68717584155982
    return np.array([[f, 0.0, 160.0], [0.0, f, 120.0], [0.0, 0.0, 1.0]])
