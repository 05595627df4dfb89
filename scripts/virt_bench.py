"""Step times of the ranks of a sharded job emulated on ONE GPU (no NVLink: what the sharded kernels cost by themselves).
    python scripts/virt_bench.py c4 2 ["GK_DEBUG_NOWAIT=4" ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geneakit_b200.distributed import VirtualShardedPhi  # noqa: E402

wl, world = sys.argv[1], int(sys.argv[2])
ped, pro, _ = bench.make_workload(wl)
for cfg in sys.argv[3:] or [""]:
    env = dict(kv.split("=", 1) for kv in cfg.split())
    os.environ.update(env)
    sp = VirtualShardedPhi(ped, pro, world=world).upload()
    for _ in range(2):
        sp.run()
    for q, j in enumerate(sp.jobs):
        j.sync()
        t = j.step_times_ms()
        n = j.stats()["n_steps"]
        print("%-30s rank %d/%d: steps %.2f ms (avg %.3f), expansion %.2f ms, remote row GB %.1f" % (
            cfg, q, world, sum(t[:n]), sum(t[:n]) / n, t[n], j.stats()["rank_remote_row_bytes"] / 1e9), flush=True)
    sp.close()
    for k in env:
        os.environ.pop(k, None)
