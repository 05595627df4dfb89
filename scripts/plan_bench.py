"""Host planner only (no kernel): stage times of gk_phi_plan on a workload, GK_PLAN_TIMING=1 prints them to stderr.
    GK_PLAN_TIMING=1 python scripts/plan_bench.py c4 [repeats]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geneakit_b200 import _lib  # noqa: E402
from geneakit_b200._lib import lib, check  # noqa: E402

ped, pro, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
r = ped.ranks(pro)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    opts = _lib.make_opts(device=0)
    job = C.c_void_p()
    t0 = time.time()
    check(lib.gk_phi_plan(len(ped), ped.sire, ped.dam, len(pro), r, C.byref(opts), C.byref(job)))
    print("gk_phi_plan %.1f ms" % ((time.time() - t0) * 1e3), file=sys.stderr, flush=True)
    lib.gk_phi_free(job)
