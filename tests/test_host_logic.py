"""Host-side logic that needs no GPU: pattern fixtures, sharding, and the two-phase statistics
combination across ranks (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle as orc
from pnp_solver_test_b200 import patterns as pt


def test_patterns():
    a, h = pt.get_golden_pattern("Alexander"), pt.get_golden_pattern("Holly")
    assert len(a) == len(h) == 15 and list(a) == list(h)
    assert a["nose_t_54"] == [-0.005, 0.0455, -0.03] and h["chin_t_16"] == [0.0, 0.098, 0.0]
    assert all(k in a for k in pt.LM_KEY_LIST_6)
    g = load_golden("qeif_n6_q")
    np.testing.assert_array_equal(pt.pattern_array(a, pt.LM_KEY_LIST_6), g["pattern"])
    p68, p1024 = pt.synthetic_pattern(68), pt.synthetic_pattern(1024)
    assert len(p68) == 68 and len(p1024) == 1024 and list(p68)[:15] == list(a)
    np.testing.assert_array_equal(pt.pattern_array(p68), load_golden("lm_n68_q")["pattern"])
    np.testing.assert_array_equal(pt.pattern_array(p1024), load_golden("lm_n1024_q")["pattern"])
    np.testing.assert_array_equal(pt.default_camera_matrix(), g["K"])


def test_shard_range_partitions_exactly():
    from pnp_solver_test_b200.workload import shard_range
    for total in (0, 1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stats_worker(rank, world_size, port, q):
    import torch.distributed as dist
    from pnp_solver_test_b200 import workload as wl
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    g = load_golden("stress_report")
    est, gt, cls = g["report"][:, 10], g["report"][:, 11], g["depth_class"]
    lo, hi = wl.shard_range(len(est), rank, world_size)
    e, t, c = est[lo:hi], gt[lo:hi], cls[lo:hi]
    n_class = 12
    # per-shard sums exactly as kernels k_stats<1>/<2> define them (here in NumPy: no GPU)
    s1 = np.zeros((n_class, 4))
    for k in range(n_class):
        m = c == k
        s1[k] = [m.sum(), (e[m] / t[m]).sum(), (e[m] - t[m]).sum(), 0.0]
    s1 = wl.reduce_phase1(torch.from_numpy(s1)).numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = np.nan_to_num(s1[:, 2] / s1[:, 0])                  # what k_stats<2> derives from the reduced sums
    s2 = np.zeros((n_class, 4))
    mx = np.zeros(n_class)
    for k in range(n_class):
        m = c == k
        d = (e[m] - t[m]) - mean[k]
        s2[k] = [(d * d).sum(), np.abs(e[m] - t[m]).sum(), np.abs(d).sum(), 0.0]
        mx[k] = np.abs(d).max() if m.any() else 0.0
    s2, mx = wl.reduce_phase2(torch.from_numpy(s2), torch.from_numpy(mx))
    out = wl.finalize_stats(torch.from_numpy(s1), s2, mx).numpy()
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_phase_statistics_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = load_golden("stress_report")["stats_by_depth"][0]          # depth quantity, 12 depth classes
    has = ~np.isnan(ref[:, 0])
    np.testing.assert_allclose(out[has], ref[has], rtol=1e-11, atol=1e-13)
    assert (out[~has, 0] == 0).all()          # empty classes report n = 0 (the reference returns None)


def test_class_bins_match_reference_classification():
    from pnp_solver_test_b200 import workload as wl
    g = load_golden("stress_report")
    cls = np.digitize(g["gt"][:, 0] * 100.0, wl.CLASS_BINS["depth"])
    np.testing.assert_array_equal(cls, g["depth_class"])
    assert wl.CLASS_LABELS["depth"][0] == "20" and wl.CLASS_LABELS["depth"][-1] == "240"
