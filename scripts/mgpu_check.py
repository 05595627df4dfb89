"""torchrun -n N scripts/mgpu_check.py: the sharded job over real NCCL / NVLink against the one-GPU job."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geneakit_b200 as gk                      # noqa: E402
from geneakit_b200.distributed import ShardedPhi   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ok = True
    # the last case is a 30-generation bottleneck population: RANKED mode (mirror tiles stored into peer rows)
    for (n, gens, m, seed, founders) in ((3000, 6, 200, 1, None), (60000, 12, 3000, 5, None), (200000, 12, 10000, 3, None),
                                         (40000, 30, 1500, 5, 40)):
        ped, pro = gk.synthetic_pedigree(n, gens, m, seed=seed, founders=founders)
        sp = ShardedPhi(ped, pro, rank=rank, world=world, device=local, stream=stream.cuda_stream).upload()
        for _ in range(2):
            sp.run()
        first = np.asarray(sp.to_host()).copy()
        sp.run_pass()                                  # planned again, cut by cut, while the device runs
        mine = np.asarray(sp.to_host())
        assert mine.tobytes() == first.tobytes(), "run_pass differs from run"
        mean = sp.mean()
        one = gk.cgeneakit.DevicePhi(ped, pro, device=local).upload().run()
        want = np.asarray(one.to_host())
        same = mine.tobytes() == np.ascontiguousarray(want[sp.row0:sp.row1]).tobytes()
        same_mean = abs(mean - one.mean()) <= 1e-15 * max(1.0, abs(mean))
        print("rank %d/%d n=%d m=%d rows [%d,%d): matrix %s mean %s" % (rank, world, n, m, sp.row0, sp.row1,
              "bit-exact" if same else "DIFFERS", "ok" if same_mean else "DIFFERS (%r vs %r)" % (mean, one.mean())), flush=True)
        ok = ok and same and same_mean
        one.close()
        sp.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if float(t[0]) == 1.0 else 1)


if __name__ == "__main__":
    main()
