#!/usr/bin/env python
"""Why SURVEY.md 8(f1)'s EKF2 and UKF2 have no 1e-9 parity contract: the UNMODIFIED reference's own
sensitivity to a 1e-13 relative pixel perturbation, method by method.

    python tests/tools/probe_ekf2_ukf2.py            # needs /root/reference (build container only)

For 24 problems of the stress workload (random_stress_test.py:246-258 draw order, quantised pixels) and the
15- and 68-landmark patterns, every 12-state variant of the reference is run three times -- on the pixels and on
the pixels times (1 +/- 1e-13) -- and sens = max(|dR|, |dt| / |t3|) between the runs is stored, with the
10 cm / 10 deg pass flag of random_stress_test.py:376 for context.  A method whose own output moves by more
than 1e-10 under that perturbation cannot be matched to 1e-9 by ANY other implementation (LDL^T for SVD-pinv,
a different summation order): that is the criterion behind the `stable` tag of the LM goldens, applied to
whole methods.  Writes tests/golden/ekf2_ukf2_sensitivity.npz; tests/test_oracle_golden.py asserts the numbers
DESIGN.md quotes from it.

Reference entry points: solve_pnp_EKF2_single_pattern PNP_SOLVER_LIB.py:1775-1999 (covariance form, pinv of the
(2n+9)x(2n+9) innovation matrix), solve_pnp_EIF2_single_pattern :2001-2276, solve_pnp_UKF2_single_pattern
:2277-2565 (alpha = 1e-3: weights -999 999 / +41 666.7), solve_pnp_LM_single_pattern :2567-2769.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METHODS = ("EKF2", "UKF2", "EIF2", "LM")
B = 24


def main():
    if not os.path.isdir("/root/reference/scripts"):
        print("no /root/reference here: nothing to probe (the committed .npz holds the last result)")
        return 0
    import make_golden as mg
    from pnp_solver_test_b200 import patterns as pt
    PNPS, _ = mg.load_reference()
    K = pt.default_camera_matrix()
    out = {"methods": np.array(METHODS), "B": B}
    for n, pattern in ((15, pt.get_golden_pattern("Alexander")), (68, pt.synthetic_pattern(68))):
        solver_gt = mg.quiet(PNPS.PNP_SOLVER, K, [pattern], [1.0], verbose=False)
        solver = mg.quiet(PNPS.PNP_SOLVER, K, [pattern], [1.0], verbose=False)
        keys = list(pattern.keys())
        pat_np = solver.np_point_3d_pretransfer_dict_list[0]
        rng = np.random.default_rng(900 + n)
        sens = np.zeros((len(METHODS), B))
        passed = np.zeros((len(METHODS), B), bool)
        for b in range(B):
            roll, pitch, yaw, depth, tt = mg.draw_pose(rng)
            R_gt = solver_gt.get_rotation_matrix_from_Euler(roll, yaw, pitch, is_degree=True)
            pts = mg.quiet(solver_gt.perspective_projection_golden_landmarks, R_gt, tt, is_quantized=True)
            uv = np.array([[pts[k][0, 0], pts[k][1, 0]] for k in keys])
            for m, name in enumerate(METHODS):
                fn = getattr(solver, "solve_pnp_%s_single_pattern" % name)
                r0 = mg.quiet(fn, mg.dict_from_uv(keys, uv), pat_np)
                R0, t0 = np.array(r0[0]), np.array(r0[1]).reshape(3)
                passed[m, b] = (abs(t0[2] - depth) * 100 < 10 and abs(r0[3] - roll) < 10 and abs(r0[4] - yaw) < 10
                                and abs(r0[5] - pitch) < 10)
                for sgn in (+1.0, -1.0):
                    r1 = mg.quiet(fn, mg.dict_from_uv(keys, uv * (1.0 + sgn * 1e-13)), pat_np)
                    d = max(np.abs(np.array(r1[0]) - R0).max(), np.abs(np.array(r1[1]).reshape(3) - t0).max() / abs(t0[2]))
                    sens[m, b] = max(sens[m, b], d if np.isfinite(d) else np.inf)
        out["sens_n%d" % n], out["passed_n%d" % n] = sens, passed
        for m, name in enumerate(METHODS):
            s = sens[m]
            print("n=%2d %-4s sensitivity to a 1e-13 pixel perturbation: min %.1e median %.1e max %.1e | below 1e-10: %2d/%d | passes 10 cm / 10 deg: %2d/%d"
                  % (n, name, s.min(), np.median(s), s.max(), (s < 1e-10).sum(), B, passed[m].sum(), B))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ekf2_ukf2_sensitivity.npz"), **out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
