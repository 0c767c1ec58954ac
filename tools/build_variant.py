#!/usr/bin/env python
"""Developer tool: build a VARIANT of libpnpb200.so for A/B measurements without touching the product build.

    python tools/build_variant.py <name> --units aux,kernels_f64_g1 -DPNP_REPORT_UNROLL=4 [-Xptxas -v ...]

Compiles the named units (object names of pnp_solver_test_b200/build.py) with the extra flags into
pnp_solver_test_b200/csrc/_obj/variant_<name>/, links them with the product's other objects into
pnp_solver_test_b200/variants/libpnpb200_<name>.so (git-ignored, travels to the GPU box), and prints the path.
Select it at run time with PNPB200_LIB=<path> (read by pnp_solver_test_b200/_lib.py).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnp_solver_test_b200 import build as b


def main():
    name = sys.argv[1]
    units, flags, i, base = None, [], 2, True
    while i < len(sys.argv):
        if sys.argv[i] == "--units":
            units = sys.argv[i + 1].split(",")
            i += 2
        elif sys.argv[i] == "--no-base":                   # the product objects are known to be up to date (parallel variant builds)
            base = False
            i += 1
        else:
            flags.append(sys.argv[i])
            i += 1
    if base:
        b.build()                                           # the product objects the variant links against
    vdir = os.path.join(b.OBJ, "variant_" + name)
    os.makedirs(vdir, exist_ok=True)
    outdir = os.path.join(os.path.dirname(b.LIB_PATH), "variants")
    os.makedirs(outdir, exist_ok=True)
    objs, todo = [], []
    for uname, src, defs in b.UNITS:
        if units is None or uname in units:
            obj = os.path.join(vdir, uname + ".o")
            todo.append([b._nvcc()] + b._ccbin() + b.NVCC_FLAGS + flags + defs + ["-c", src, "-o", obj])
        else:
            obj = os.path.join(b.OBJ, uname + ".o")
        objs.append(obj)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for r in ex.map(lambda c: subprocess.run(c, cwd=b.CSRC, capture_output=True, text=True), todo):
            if r.returncode != 0 or "-v" in flags:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                return 1
    lib = os.path.join(outdir, "libpnpb200_%s.so" % name)
    subprocess.check_call([b._nvcc()] + b._ccbin() + ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", lib] + objs,
                          cwd=b.CSRC)
    print(lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
