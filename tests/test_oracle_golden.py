"""Pin the CPU oracle (oracle/pnp_oracle.c) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/make_golden.py in the build container) and against the
reference's only own fixture (GT_R_t_dict.pkl -> euler_fixture.npz)."""
import numpy as np
import pytest

from conftest import LM_TRACE_GOLDENS, SOLVER_GOLDENS, compare_lm_trace, compare_solutions, load_golden
from oracle import oracle as orc


@pytest.mark.parametrize("name", SOLVER_GOLDENS)
def test_oracle_matches_reference_solver(name):
    g = load_golden(name)
    o = orc.solve_batch(str(g["method"]), g["uv"], g["pattern"], g["K"])
    stats = compare_solutions(o, g, mask=g["stable"], iters_mask=g["iters_stable"])
    # LM is chaotic on part of its inputs; the stable fraction must be what the
    # survey measured, otherwise the tag (and the contract) is meaningless.
    # (eif2_n1024 holds 6 problems, one of them sensitive at the 1e-10 cut)
    floor = {"lm": 0.7, "eif2": 0.8}.get(str(g["method"]), 0.999)
    assert g["stable"].mean() > floor, stats


@pytest.mark.parametrize("name", LM_TRACE_GOLDENS)
def test_oracle_lm_matches_reference_iteration_by_iteration(name):
    """LM on ALL inputs, incl. the 15-25 % where the 14-iteration output is chaotic: the oracle follows the
    unmodified reference to 1e-9 at every iteration up to the one where the reference's own output stops
    being reproducible (its +-1e-13 runs part by > 1e-10)."""
    g = load_golden(name)
    assert len(LM_TRACE_GOLDENS) == 4
    compare_lm_trace(lambda k: orc.solve_batch("lm", g["uv"], g["pattern"], g["K"], params=orc.default_params(max_it=k)), g, name)


def test_oracle_solve_pnp_two_patterns():
    g = load_golden("solve_pnp_two_patterns")
    idx = g["key_index"]
    o = orc.solve_batch("qeif", g["uv"][:, idx], g["patterns"][:, idx], g["K"])
    assert (o["best_pattern"] == g["best_pattern"]).all()
    assert np.abs(o["R"] - g["R"]).max() < 1e-9 and np.abs(o["t"] - g["t"]).max() < 1e-9
    assert np.abs(o["res_norm"] - g["res_norm"]).max() < 1e-12


def test_oracle_euler_fixture():
    """The reference's own fixture: 1434 (file-name Euler labels -> np_R_GT) vectors."""
    g = load_golden("euler_fixture")
    rpy, R = g["roll_yaw_pitch_deg"], g["R"]
    for i in range(len(R)):
        assert np.abs(orc.R_from_euler(*rpy[i], is_degree=True) - R[i]).max() < 1e-12
        back = np.array(orc.euler_from_R(R[i], is_degree=True))
        assert np.abs(back - rpy[i]).max() < 1e-9
    assert np.abs(g["t"][:, 2] - g["distance_cm"] / 100.0).max() < 1e-12


def test_oracle_report_and_statistics():
    g = load_golden("stress_report")
    o = orc.report_batch(g["pattern"], g["uv"], g["K"], g["R"], g["t"], g["euler"], g["gt"])
    assert np.abs(o["report"] - g["report"]).max() < 1e-10
    assert (o["flags"] == g["flags"]).all() and (o["max_idx"] == g["max_idx"]).all()
    rep, gt = g["report"], g["gt"]
    cols = [(rep[:, 10], rep[:, 11])] + [(rep[:, 12 + i], gt[:, 1 + i]) for i in range(3)]
    for q, (e, t) in enumerate(cols):
        np.testing.assert_allclose(np.array(orc.stats_of(e, t)), g["stats_all"][q], rtol=1e-12, atol=1e-14)
        for c in range(12):
            sel = g["depth_class"] == c
            ref = g["stats_by_depth"][q, c]
            if sel.sum() == 0:
                assert np.isnan(ref).all()
            else:
                np.testing.assert_allclose(np.array(orc.stats_of(e[sel], t[sel])), ref, rtol=1e-12, atol=1e-14)


def test_oracle_synth_is_shard_invariant_and_sane():
    from pnp_solver_test_b200 import patterns as pt
    P, K = pt.pattern_array(pt.synthetic_pattern(68)), pt.default_camera_matrix()
    full = orc.synth(0, 300, P, K)
    a, b = orc.synth(0, 100, P, K), orc.synth(100, 200, P, K)
    assert np.array_equal(full["uv"], np.concatenate([a["uv"], b["uv"]]))
    assert np.array_equal(full["gt"], np.concatenate([a["gt"], b["gt"]]))
    gt = full["gt"]
    assert (gt[:, 0] >= 0.2).all() and (gt[:, 0] <= 2.25).all() and (np.abs(gt[:, 1:]) <= 45).all()
    assert np.array_equal(full["uv"], np.rint(full["uv"]))          # quantised to integer pixels
    noisy = orc.synth(0, 2000, P, K, orc.default_synth(is_quantized=False, noise_sigma_px=2.0))
    clean = orc.synth(0, 2000, P, K, orc.default_synth(is_quantized=False))
    d = (noisy["uv"] - clean["uv"]).ravel()
    assert abs(d.mean()) < 0.02 and abs(d.std() - 2.0) < 0.02
    # QEIF on the synthetic stress workload passes the reference's own 10 cm / 10 deg criterion
    P15 = pt.pattern_array(pt.get_golden_pattern())
    w = orc.synth(0, 400, P15, K)
    idx = [list(pt.get_golden_pattern()).index(k) for k in pt.LM_KEY_LIST_6]
    s = orc.solve_batch("qeif", w["uv"][:, idx], P15[idx], K)
    r = orc.report_batch(P15, w["uv"], K, s["R"], s["t"], s["euler"], w["gt"])
    assert r["flags"].all(axis=1).mean() > 0.95


def test_oracle_fragility_analysis_matches_reference():
    """face_variation_test.py's heap selection + get_most_fragile_point_and_perturbation_direction
    (the script's own function, exec'd unmodified by oracle/make_golden.py) on 600 problems."""
    g = load_golden("fragility")
    keys = [str(k) for k in g["keys"]]
    for q, name in enumerate(("depth", "roll", "pitch", "yaw")):
        r = orc.fragility_of(np.abs(g["err"][:, q]), g["perturb"], keys)
        assert np.array_equal(np.array([r["fragile_point_count_dict"][k] for k in keys]), g[name + "_count"])
        assert [k for _, k in r["fragile_point_sorted_list"]] == [str(k) for k in g[name + "_sorted_keys"]]
        assert np.abs(r["top_similarity"] - g[name + "_similarity"]).max() < 1e-12
        for i in range(5):
            d, ref = r["top_perturbation"][i], g[name + "_directions"][i]
            assert min(np.abs(d - ref).max(), np.abs(d + ref).max()) < 1e-9
        assert r["value_max"] == g[name + "_value_max"] and abs(r["top_value_mean"] - g[name + "_value_mean"]) < 1e-15


def test_oracle_face_variation_workload():
    from pnp_solver_test_b200 import patterns as pt
    pat = pt.get_golden_pattern()
    keys = list(pat.keys())
    P, K = pt.pattern_array(pat), pt.default_camera_matrix()
    fixed = keys.index("eye_c_51")
    w = orc.synth_face_variation(3, 500, P, K, fixed, 0.02)
    assert np.abs(np.linalg.norm(w["perturb"].reshape(500, -1), axis=1) - 0.02).max() < 1e-15   # unit_vec * radius (:326-331)
    assert (w["perturb"][:, fixed] == 0).all() and (np.abs(w["perturb"]).sum(axis=2) > 0)[:, np.arange(15) != fixed].all()
    plain = orc.synth(3, 500, P, K)
    assert np.array_equal(plain["gt"], w["gt"])                     # same pose stream
    assert 0.2 < np.abs(plain["uv"] - w["uv"]).max() < 60.0          # a 2 cm pattern change moves pixels, not poses
    g = load_golden("fragility")
    again = orc.synth_face_variation(0, 600, g["pattern"], g["K"], int(g["fixed_index"]), 0.02, orc.default_synth(seed=int(g["seed"])))
    assert np.array_equal(again["perturb"], g["perturb"]) and np.array_equal(again["uv"], g["uv"])


def test_ekf2_and_ukf2_have_no_well_posed_parity_target():
    """the reference's own EKF2 (PNP_SOLVER_LIB.py:1775-1999) and UKF2 (:2277-2565) outputs move
    by more than the parity tolerance under a 1e-13 relative pixel perturbation on EVERY probed problem
    (tests/tools/probe_ekf2_ukf2.py, run on the unmodified reference; 24 problems x 15 / 68 landmarks), while
    EIF2 -- the same filter in information form, which IS built -- is reproducible to 1e-11.  This pins the numbers
    DESIGN.md section 4 quotes as the reason those two variants carry no 1e-9 contract."""
    g = load_golden("ekf2_ukf2_sensitivity")
    names = [str(m) for m in g["methods"]]
    for n in (15, 68):
        s = {m: g["sens_n%d" % n][i] for i, m in enumerate(names)}
        assert s["EKF2"].min() > 1e-6 and np.median(s["EKF2"]) > 1e-4      # 4+ orders of magnitude above 1e-9
        assert s["UKF2"].min() > 1e-10 and np.median(s["UKF2"]) > 5e-10     # at / above the 1e-9 tolerance itself
        assert s["UKF2"].max() < 1e-8                                       # ... so a stated 1e-7 bound would be assertable
        assert s["EIF2"].max() < 1e-11
        assert (s["LM"] < 1e-10).mean() >= 0.8                              # LM: well-posed on most inputs, chaotic on the rest


def test_oracle_on_image_points_with_homogeneous_coordinate_not_one():
    """The reference multiplies the (3,1) image vector it is given by K^-1 whatever its third entry (PNP_SOLVER_LIB.py:3307;
    its projection emits w = -1 behind the camera, :4548).  The drop-in forms the same product and solves on the normalised
    coordinates with K = I: checked here with the oracle against the unmodified reference's outputs."""
    g = load_golden("homogeneous_n15")
    nu = np.einsum("ij,bnj->bni", np.linalg.inv(g["K"]), g["uvw"])[..., :2]
    idx = g["key_index"]
    o = orc.solve_batch("qeif", nu[:, idx], g["pattern"][idx], np.eye(3))
    assert np.abs(o["R"] - g["qeif6_R"]).max() < 1e-9 and np.abs(o["t"] - g["qeif6_t"]).max() < 1e-9
    assert np.abs(o["res_norm"] - g["qeif6_res_norm"]).max() < 1e-11
    o = orc.solve_batch("linear_f2", nu, g["pattern"], np.eye(3))
    assert np.abs(o["R"] - g["linear_f2_R"]).max() < 1e-9 and np.abs(o["t"] - g["linear_f2_t"]).max() < 1e-9
