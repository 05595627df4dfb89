"""Host-side logic of the multi-GPU path, on the CPU with the gloo backend and world_size 2:
row-block sharding, identical plans on every rank, and the phiMean all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import geneakit_b200 as gk
        from geneakit_b200.distributed import shard_rows, combine_mean
        from conftest import load_golden
        from oracle.api import phi_mean
        g = load_golden("overlap400_s1")
        phi = g["phi"]
        m = phi.shape[0]
        r0, r1 = shard_rows(m, rank, world)
        blocks = [None] * world
        dist.all_gather_object(blocks, (r0, r1))
        assert blocks[0][0] == 0 and blocks[-1][1] == m
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))      # contiguous, disjoint
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
        # every rank builds the same plan (the recurrence is replicated)
        r = g["rows"]
        ped = gk.cgeneakit.load_pedigree_from_vectors(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
        job = gk.cgeneakit.DevicePhi(ped, [int(x) for x in g["pro"]], rank=rank, world=world)
        assert (job.row0, job.row1) == (r0, r1)
        plans = [None] * world
        dist.all_gather_object(plans, (job.cut_sizes(), job.class_sizes(), job.stats()["cells_device"]))
        assert all(p == plans[0] for p in plans)
        job.close()
        # partial sums of this rank's row block -> all-reduce -> gen.phiMean
        mine = phi[r0:r1]
        s = float(mine.sum())
        d = float(sum(phi[i, i] for i in range(r0, r1)))
        mean = combine_mean(s, d, m)
        assert abs(mean - phi_mean(phi)) <= 1e-15
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, "fail: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_mean_on_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
