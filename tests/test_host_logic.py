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
    # the exchange as `statistics` does it (ONE all_gather of s2 | max and a local fold) and as two all-reduces
    flat = torch.cat([torch.from_numpy(s2).flatten(), torch.from_numpy(mx)]).contiguous()
    s2m, mxm = wl.reduce_phase2(flat[:n_class * 4].view(n_class, 4), flat[n_class * 4:], flat23=flat)
    s2, mx = wl.reduce_phase2(torch.from_numpy(s2), torch.from_numpy(mx))
    assert torch.allclose(s2m, s2, rtol=1e-15, atol=0) and torch.equal(mxm, mx)
    out = wl.finalize_stats(torch.from_numpy(s1), s2m, mxm).numpy()
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


def test_float_repr_formatter_matches_python():
    """The native CSV writer's float formatter against repr(float) (what csv.DictWriter writes)."""
    import ctypes as C
    from pnp_solver_test_b200 import _lib
    rng = np.random.default_rng(0)
    vals = [0.0, -0.0, 1.0, -3.0, 100000.0, 1e-5, 0.0001234, 9.999e-5, 1.5e16, 9999999999999998.0, 123456789012345678.0, 0.1 + 0.2,
            1e22, 5e-324, 1.7976931348623157e308, float("inf"), float("-inf"), float("nan"), 2.5e-7, 123.456, 1e15, 1e16, 0.5, 1 / 3]
    vals += list(rng.normal(size=300) * 10.0 ** rng.integers(-20, 20, 300)) + list(rng.uniform(-200, 200, 300))
    buf = C.create_string_buffer(64)
    for v in vals:
        _lib.check(_lib.lib.pnpb200_format_repr(C.c_double(float(v)), buf, 64), "pnpb200_format_repr")
        assert buf.value.decode() == repr(float(v)), (v, buf.value)


def test_result_and_statistic_writers_match_reference(tmp_path):
    """write_result_csv / write_statistic_txt / write_statistic_csv against the files the unmodified
    TEST_TOOLBOX writers produced from the reference's own result dicts (tests/golden/stress_report.npz)."""
    import csv
    import io
    from pnp_solver_test_b200 import workload as wl
    g = load_golden("stress_report")
    keys = [str(k) for k in g["keys"]]
    # the reference's distance_GT is (depth * 100) * 0.01; put it where report_batch's column 11 is
    p = wl.write_result_csv(tmp_path / "r.csv", g["report"], g["flags"], g["max_idx"], g["res_norm"], g["gt"], keys, n_threads=3)
    mine = list(csv.reader(io.StringIO(open(p, newline="").read())))
    ref = list(csv.reader(io.StringIO(str(g["result_csv"]))))
    assert open(p, newline="").read().count("\r\n") == len(ref)            # csv's default line terminator
    assert mine[0] == ref[0] and len(mine) == len(ref) == 257
    num = [i for i, h in enumerate(ref[0]) if h not in ("idx", "file_name", "drpy", "class", "fail_count", "pass_count")
           and not h.startswith("is_") and not h.endswith("_key")]
    for a, b in zip(mine[1:], ref[1:]):
        for i, h in enumerate(ref[0]):
            if i in num:
                assert abs(float(a[i]) - float(b[i])) <= 1e-9 * max(1.0, abs(float(b[i]))), (h, a[i], b[i])
            elif h == "drpy":
                x, y = eval(a[i]), eval(b[i])
                assert len(x) == 4 and max(abs(u - v) for u, v in zip(x, y)) < 1e-12
            else:
                assert a[i] == b[i], (h, a[i], b[i])
    # byte-identical wherever the numbers are: rows whose floats all agree exactly
    same = sum(1 for a, b in zip(mine[1:], ref[1:]) if a == b)
    assert same >= 0                                                     # informational; numerics differ in the last digits
    # statistics files: text produced from the golden statistics rows
    for q, (name, unit, scale) in enumerate((("depth", "cm", 100.0), ("roll", "deg.", 1.0), ("pitch", "deg.", 1.0), ("yaw", "deg.", 1.0))):
        st = {"all": g["stats_all"][q], "by_depth": np.nan_to_num(g["stats_by_depth"][q])}
        d = wl.statistic_dicts(st, wl.CLASS_LABELS["depth"], unit, scale)
        txt = wl.write_statistic_txt(d, tmp_path / "s.txt", class_name="distance", statistic_data_name=name)
        assert txt == str(g["stat_txt_" + name]) and open(tmp_path / "s.txt").read() == txt
        wl.write_statistic_csv(d, tmp_path / "s.csv", is_horizontal=True)
        assert open(tmp_path / "s.csv", newline="").read() == str(g["stat_csv_" + name])


def test_pixel_packing_is_exact_or_refused():
    """pnpb200_pack_i16 (host side of the packed PCIe transfer): int16 copy and 'exact' verdict for
    FP64 / FP32, one and several threads, lengths that are not a multiple of the vector width."""
    import ctypes as C
    from pnp_solver_test_b200 import _lib
    rng = np.random.default_rng(9)

    def pack(a, threads):
        dst = np.full(a.shape, 12345, np.int16)
        rc = _lib.lib.pnpb200_pack_i16(C.c_int(0 if a.dtype == np.float64 else 1), a.ctypes.data_as(C.c_void_p), C.c_int64(a.size),
                                       dst.ctypes.data_as(C.POINTER(C.c_int16)), C.c_int(threads))
        return rc, dst

    for dt in (np.float64, np.float32):
        for n in (0, 1, 15, 16, 17, 1000, 200003):
            a = rng.integers(-32768, 32768, n).astype(dt)
            if n > 2:
                a[0], a[1], a[2] = -32768, 32767, -0.0
            for th in (1, 4):
                rc, dst = pack(a, th)
                assert rc == 1 and np.array_equal(dst, a.astype(np.int16)), (dt, n, th)
        base = rng.integers(-2000, 2000, 200003).astype(dt)
        for pos in (0, 7, 16, 100000, 200002):
            for bad in (0.5, 32768.0, -32769.0, 1e30, np.nan, np.inf, -np.inf, 1e-3):
                a = base.copy()
                a[pos] = bad
                assert pack(a, 3)[0] == 0, (dt, pos, bad)
        assert pack(base, 3)[0] == 1
    assert _lib.lib.pnpb200_pack_i16(C.c_int(7), None, C.c_int64(4), None, C.c_int(1)) < 0


def test_pack_workers_are_reused_across_calls_and_survive_a_fork():
    """The packing threads outlive a call (pnpb200_pack.cpp, PackPool): many calls with changing thread counts and sizes give
    the same bytes as one thread, concurrent callers are serialised, and a forked child -- which inherits the pool object
    but none of its threads -- packs with a pool of its own instead of waiting for workers that do not exist."""
    import ctypes as C
    import os
    import threading
    from pnp_solver_test_b200 import _lib
    rng = np.random.default_rng(10)
    a = rng.integers(-3000, 3000, 700001).astype(np.float64)
    want = a.astype(np.int16)

    def pack(src, threads):
        dst = np.zeros(src.shape, np.int16)
        rc = _lib.lib.pnpb200_pack_i16(C.c_int(0), src.ctypes.data_as(C.c_void_p), C.c_int64(src.size),
                                       dst.ctypes.data_as(C.POINTER(C.c_int16)), C.c_int(threads))
        return rc, dst

    for rep in range(40):
        th = (1, 7, 2, 16, 3)[rep % 5]
        n = (700001, 65536, 32769, 5)[rep % 4]
        rc, dst = pack(a[:n], th)
        assert rc == 1 and np.array_equal(dst, want[:n]), (rep, th, n)
    results = []
    callers = [threading.Thread(target=lambda: results.append(pack(a, 5))) for _ in range(4)]
    for t in callers:
        t.start()
    for t in callers:
        t.join()
    assert len(results) == 4 and all(rc == 1 and np.array_equal(dst, want) for rc, dst in results)
    # the fork check in a process of its own (numpy + the library, nothing else): the only threads alive at the fork are
    # the pool's sleeping workers, whatever this pytest process happens to run beside the test
    script = r"""
import ctypes as C, os, sys, time
import numpy as np
lib = C.CDLL(sys.argv[1])
a = np.random.default_rng(10).integers(-3000, 3000, 700001).astype(np.float64)
want = a.astype(np.int16)
def pack(threads):
    dst = np.zeros(a.shape, np.int16)
    rc = lib.pnpb200_pack_i16(C.c_int(0), a.ctypes.data_as(C.c_void_p), C.c_int64(a.size), dst.ctypes.data_as(C.POINTER(C.c_int16)), C.c_int(threads))
    return rc == 1 and np.array_equal(dst, want)
assert pack(6)                                   # the parent's pool exists now
pid = os.fork()
if pid == 0:
    ok = False
    try:
        ok = pack(6)
    finally:
        os._exit(0 if ok else 1)
for _ in range(400):                             # a child waiting for threads it does not have would hang: bounded wait
    done, status = os.waitpid(pid, os.WNOHANG)
    if done:
        sys.exit(0 if (os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0) else 3)
    time.sleep(0.05)
os.kill(pid, 9)
os.waitpid(pid, 0)
sys.exit(4)                                      # hung
"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", script, _lib.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, "forked child: %s" % {3: "wrong result", 4: "hung in pnpb200_pack_i16"}.get(r.returncode, r.stderr[-300:])


def test_approval_masks():
    """approval_mask = TEST_TOOLBOX.approval_func_small_angle / _large_angle (:959-970), thresholds inclusive."""
    from pnp_solver_test_b200 import workload as wl
    gt = torch.tensor([[1.0, 0.0, 0.0, 0.0], [1.0, 30.0, -30.0, 30.0], [1.0, 30.000001, 0.0, 0.0], [1.0, 0.0, -31.0, 0.0],
                       [1.0, 0.0, 0.0, 45.0], [9.0, -29.9, 29.9, -29.9]], dtype=torch.float64)
    small = wl.approval_mask(gt, "small_angle")
    assert small.tolist() == [True, True, False, False, False, True]
    assert wl.approval_mask(gt, "large_angle").tolist() == [not x for x in small.tolist()]
    with pytest.raises(ValueError):
        wl.approval_mask(gt, "medium")
    # against the reference's get_classified_result(approval_func=...) on its own result list: a sample that fails the
    # approval leaves its distance class (class id -1 for the statistics kernels)
    g = load_golden("stress_report")
    cls = g["depth_class"].astype(np.int64)
    pairs = [(g["report"][:, 10], g["report"][:, 11])] + [(g["report"][:, 12 + i], g["gt"][:, 1 + i]) for i in range(3)]
    for kind in ("small_angle", "large_angle"):
        keep = wl.approval_mask(torch.from_numpy(g["gt"]), kind).numpy()
        ref = g["stats_by_depth_" + kind]
        for q, (est, true) in enumerate(pairs):
            for c in range(12):
                m = (cls == c) & keep
                if not m.any():
                    assert np.isnan(ref[q, c, 0])
                    continue
                np.testing.assert_allclose(np.array(orc.stats_of(est[m], true[m])), ref[q, c], rtol=1e-12, atol=1e-14)


def test_drpy_tables_match_reference(tmp_path):
    """drpy_statistic_dict / write_drpy_statistic_csv against the eleven class-combination tables the
    unmodified get_all_class_seperated_result / get_drpy_statistic / write_drpy_2_depth_statistic_CSV
    wrote from the reference's result dicts (TEST_TOOLBOX.py:1215-1346)."""
    from oracle import oracle as orc
    from pnp_solver_test_b200 import workload as wl
    g = load_golden("stress_report")
    st = orc.drpy_stats(g["report"], g["gt"], [wl.CLASS_BINS[q] for q in wl.DRPY_ORDER])
    assert st["depth"].shape == wl.drpy_shape() + (7,) and int(st["depth"][..., 0].sum()) == g["gt"].shape[0]
    paths = wl.drpy_analysis_and_saving(None, None, str(tmp_path) + "/", "stat_", "data.txt", stats=st)
    assert len(paths) == 11
    for p in paths:
        key = "drpy_csv_" + os.path.basename(p)[len("stat_data_drpy_to_"):-4]
        assert open(p, newline="").read() == str(g[key]), key


def _topk_worker(rank, world, port, q):
    import torch.distributed as dist
    from pnp_solver_test_b200 import workload as wl
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("fragility")
    err = np.abs(g["err"])
    err[::7, 1] = 3.25                                              # exact ties across both shards
    lo, hi = wl.shard_range(err.shape[0], rank, world)
    vals = [err[lo:hi, qq] for qq in range(4)]
    thr = wl.topk_thresholds(vals, 60, idx0=lo, hist_fn=wl.numpy_histogram(vals, lo))
    sel = []
    for qq in range(4):
        khi, klo = wl.key_of(vals[qq], lo)
        m = (khi > np.uint64(thr[0][qq])) | ((khi == np.uint64(thr[0][qq])) & (klo >= np.uint32(thr[1][qq])))
        sel.append((np.flatnonzero(m) + lo).tolist())
    gathered = [None] * world
    dist.all_gather_object(gathered, sel)
    if rank == 0:
        q.put((gathered, err))
    dist.barrier()
    dist.destroy_process_group()


def test_topk_selection_over_gloo_world_size_2():
    """The radix select of workload.fragility_analysis across two shards (histograms all-reduced over
    gloo, NumPy standing in for the histogram kernel): the union of the per-shard selections is exactly
    the top-k of the whole batch in the order the script's heaps pop (largest |err|, then smaller idx)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_topk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for qq in range(4):
        got = sorted(gathered[0][qq] + gathered[1][qq])
        want = sorted(sorted(range(err.shape[0]), key=lambda i: (-err[i, qq], i))[:60])
        assert got == want, qq
