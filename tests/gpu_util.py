"""Helpers for the -m gpu tests: everything goes through the C ABI via the package."""
import numpy as np
import torch

from oracle import oracle as orc


def dev(x, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dtype)


def to_np(out):
    o = {k: v.detach().double().cpu().numpy() if v.is_floating_point() else v.detach().cpu().numpy() for k, v in out.items()}
    if "R" in o:
        o["R"] = o["R"].reshape(-1, 3, 3)
    return o


def cuda_solve(method, uv, patterns, K, mapping=0, dtype=torch.float64, point_index=None, **prm):
    import pnp_solver_test_b200 as pnp
    params = pnp.default_params(mapping=mapping, **prm)
    patterns = np.asarray(patterns)
    if patterns.ndim == 2:
        patterns = patterns[None]
    out = pnp.solve_batch(method, dev(uv, dtype), dev(patterns, dtype), K, point_index=point_index, params=params)
    torch.cuda.synchronize()
    return to_np(out)


def oracle_stability(method, uv, pattern, K, ref=None):
    """SURVEY.md 8c stability tag computed with the oracle itself: rerun with pixels * (1 +/- 1e-13)."""
    ref = ref or orc.solve_batch(method, uv, pattern, K)
    worst = np.zeros(uv.shape[0])
    it_stable = np.ones(uv.shape[0], bool)
    for sgn in (+1.0, -1.0):
        o = orc.solve_batch(method, uv * (1.0 + sgn * 1e-13), pattern, K)
        d = np.maximum(np.abs(o["R"] - ref["R"]).reshape(len(worst), -1).max(axis=1),
                       np.abs(o["t"] - ref["t"]).max(axis=1) / np.abs(ref["t"][:, 2]))
        worst = np.maximum(worst, d)
        it_stable &= (o["iters"] == ref["iters"])
    oracle_stability.last_worst = worst                     # the sensitivities themselves, for tools that want more than the tag
    return ref, worst < 1e-10, it_stable
