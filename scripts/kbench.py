#!/usr/bin/env python
"""Kernel experiments: build one workload, then time the step kernel under several environment settings
(GK_* variables are read when a job is planned / uploaded / run).

    python scripts/kbench.py c4 "GK_STEP_KERNEL=g4" "GK_STEP_KERNEL=g4 GK_RING=4" "GK_STEP_KERNEL=ldgsts"
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import geneakit_b200 as gk  # noqa: E402
from geneakit_b200._lib import lib, check  # noqa: E402


def main():
    wl = sys.argv[1]
    ped, pro, _ = bench.make_workload(wl)
    for cfg in sys.argv[2:] or [""]:
        env = dict(kv.split("=", 1) for kv in cfg.split())
        for k, v in env.items():
            os.environ[k] = v
        job = gk.cgeneakit.DevicePhi(ped, pro).upload()
        for _ in range(2):
            job.run()
        job.sync()
        job.run().sync()
        t = job.step_times_ms()
        b = job.step_bytes()
        n = job.stats()["n_steps"]
        prof = np.zeros(16, dtype=np.uint64)
        check(lib.gk_phi_debug_prof(job._h, prof))
        step_ms = sum(t[:n])
        line = "%-44s step avg %.3f ms (%.0f GB/s alg)  last3 %s  expand %.2f ms  mean %.6e" % (
            cfg, step_ms / n, sum(b[:n]) / step_ms / 1e6, ["%.2f" % x for x in t[n - 3:n]], t[n], job.mean())
        if prof.any():
            p = prof.astype(np.float64)
            line += ("\n    producer: loop %.2e cyc, idle polls %.0f (stage-bound %.0f, ring-bound %.0f) | consumers: %.2e cyc, "
                     "%.0f%% waiting for TMA | phase 2: %.2e cyc, %.0f%% waiting for phase 1, %.0f%% publishing"
                     % (p[0], p[1], p[2], p[3], p[4], 100 * p[5] / max(p[4], 1), p[7], 100 * p[8] / max(p[7], 1), 100 * p[10] / max(p[7], 1)))
        print(line, flush=True)
        job.close()
        for k in env:
            os.environ.pop(k, None)


if __name__ == "__main__":
    main()
