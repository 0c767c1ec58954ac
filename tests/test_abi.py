"""The C-ABI library loads on a CPU-only box and exports every symbol include/pnpb200.h
declares; host-only entry points behave; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pnpb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pnpb200_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pnp_solver_test_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_struct_layouts_match_header():
    from pnp_solver_test_b200 import _lib
    assert C.sizeof(_lib.Params) == 2 * 4 + 8 * 8 + 2 * 4 + 8 + 8
    assert C.sizeof(_lib.Synth) == 8 + 4 * 8 + 2 * 4 + 2 * 8 + 3 * 8
    p = _lib.default_params()
    assert (p.max_it, p.linear_it, p.lm_lambda, p.exit_tol, p.f_weight) == (14, 3, 1e-5, 1e-2, 225.68)
    assert (p.meas_sigma_px, p.proc_q, p.proc_d, p.omega0, p.res_old0, p.mapping) == (3.0, 0.1, 0.01, 1e-5, 1e-7, 0)
    assert (p.flags, p.workspace, p.workspace_bytes) == (0, None, 0)
    s = _lib.default_synth()
    assert (s.seed, s.angle_range_deg, s.depth_min_m, s.depth_max_m, s.fov_max_deg) == (42, 45.0, 0.2, 2.25, 45.0)
    assert (s.is_quantized, s.quantize_q, s.noise_sigma_px) == (1, 1.0, 0.0)
    assert _lib.lib.pnpb200_version() == 100


def test_argument_validation_needs_no_gpu():
    from pnp_solver_test_b200 import _lib
    rc = _lib.lib.pnpb200_solve_batch(0, 0, C.c_int64(4), 6, 6, None, None, 1, None, None, None,
                                      None, None, None, None, None, None, None)
    assert rc == -1                                   # PNPB200_EINVAL: null uv
    K = (C.c_double * 9)(225.7, 0, 160, 0, 225.7, 120, 0, 0, 1)
    rc = _lib.lib.pnpb200_solve_batch(0, 0, C.c_int64(4), 6, 6, C.c_void_p(0x1008), C.c_void_p(0x2000), 1, None, K, None,
                                      None, None, None, None, None, None, None)
    assert rc == -1                                   # uv not 16-byte aligned
    assert _lib.lib.pnpb200_default_params(None) == -1
    with pytest.raises(_lib.PnpB200Error):
        _lib.check(-1, "x")


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from pnp_solver_test_b200 import _lib
    d = C.c_double()
    rc = _lib.lib.pnpb200_fma_peak(0, 10, C.byref(d))
    assert rc in (-2, -3)
    with pytest.raises(_lib.PnpB200Error):
        _lib.check(rc, "pnpb200_fma_peak")


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, the workload runners or the developer
    tools may import or load it (only tests/, smoke() and bench.py's CPU arm do)."""
    for top in ("pnp_solver_test_b200", "workloads", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    for needle in ("import oracle", "from oracle", "libpnp_oracle", "pnp_oracle_", "dlopen"):
                        assert needle not in txt, (needle, os.path.join(dirpath, f))
