"""End-to-end timing of gk_phi_f64 (host arrays -> pinned host matrix) under a few settings of the result copy.
    python scripts/e2e_bench.py [workload] [gpus]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geneakit_b200 import _lib  # noqa: E402
from geneakit_b200._lib import lib, check  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ped, pro, _ = bench.make_workload(name)
m = len(pro)
ptr = lib.gk_host_alloc(m * m * 8)
r = ped.ranks(pro)
mu = C.c_double()
settings = [{"GK_D2H_MIRROR_ROWS": x} for x in ("0.5", "0.6", "0.7", "0.8", "1.0")] + [{"GK_D2H_SYMMETRIC": "0"}, {}]
for env in settings:
    for k in ("GK_D2H_DEBUG_NOMIRROR", "GK_D2H_SYMMETRIC", "GK_D2H_MIRROR_ROWS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for it in range(2):
        opts = _lib.make_opts(device=0, gpus=gpus)
        t0 = time.time()
        check(lib.gk_phi_f64(len(ped), ped.sire, ped.dam, m, r, ptr, C.byref(mu), C.byref(opts)))
        dt = time.time() - t0
        tm = np.zeros(6)
        lib.gk_phi_last_timing(tm)
        print("%-36s call %d: %.3f s  (prefix %.3f, alloc %.3f, kernels %.3f, d2h %.3f)  mean %.6e"
              % (env or "default", it, dt, tm[0], tm[1], tm[2], tm[3], mu.value), flush=True)
lib.gk_host_free(ptr)
