#!/usr/bin/env python
"""One device-resident pass of a workload (for ncu captures): python scripts/prof_one.py c4"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import geneakit_b200 as gk  # noqa: E402

ped, pro, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
print("mean", job.mean())
job.close()
