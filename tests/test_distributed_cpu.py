"""Host-side logic of the multi-GPU path, on the CPU with the gloo backend and world_size 2:
every rank builds the same plan, the panels (rows of every cut's class matrix) and the result rows
are partitioned without gaps, a rank needs nothing but ITS rows of the new matrix plus (remote)
rows of the previous one, and the phiMean all-reduce.  The arithmetic is the numpy interpreter of
the plan (tests/plan_emulator.py): no kernel is launched here."""
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
        from geneakit_b200.distributed import shard_rows, combine_mean, shard_ancestors
        cols = [None] * world
        dist.all_gather_object(cols, shard_ancestors(7399, rank, world))
        assert cols[0][0] == 0 and cols[-1][1] == 7399 and all(cols[i][1] == cols[i + 1][0] for i in range(world - 1))
        from conftest import load_golden
        from oracle.api import phi_mean
        from plan_emulator import step_view, emulate_step_rows, SINGLETON
        for name in ("overlap400_s1", "geneaJi_dup_anc"):
            g = load_golden(name)
            phi = g["phi"]
            m = phi.shape[0]
            r0, r1 = shard_rows(m, rank, world)
            blocks = [None] * world
            dist.all_gather_object(blocks, (r0, r1))
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))      # contiguous, disjoint
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
            r = g["rows"]
            ped = gk.cgeneakit.load_pedigree_from_vectors(r[:, 0], r[:, 1], r[:, 2], r[:, 3])
            job = gk.cgeneakit.DevicePhi(ped, [int(x) for x in g["pro"]], rank=rank, world=world)
            assert (job.row0, job.row1) == (r0, r1)
            st = job.stats()
            G = st["n_cuts"]
            plans = [None] * world
            dist.all_gather_object(plans, (job.cut_sizes(), job.class_sizes(), st["cells_device"], st["corrections"]))
            assert all(p == plans[0] for p in plans)                                    # identical plans
            # cut 0: founders' matrix, each rank its own rows
            v0 = step_view(job, 0)
            A = np.zeros((v0["Kpad"], v0["Kpad"]))
            for c in range(v0["K"]):
                if v0["flags"][c] & SINGLETON:
                    A[c, c] = 0.5
            d = np.full(v0["Kpad"], 0.5)
            for s in range(1, G):
                v = step_view(job, s)
                ranges = [None] * world
                dist.all_gather_object(ranges, (v["panel_begin"], v["panel_end"]))
                assert ranges[0][0] == 0 and ranges[-1][1] == v["np"]
                assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))  # panels partitioned
                c0, c1, rows, dn = emulate_step_rows(v, A, d)
                parts = [None] * world
                dist.all_gather_object(parts, (c0, c1, rows, dn))                        # stands for the peer reads + d all-gather
                A = np.zeros((v["Kpad"], v["Kpad"]))
                d = np.full(v["Kpad"], 0.5)
                for (a, b, rr, dd) in parts:
                    A[a:b] = rr
                    d[a:b] = dd
                assert np.array_equal(A, A.T)
            import ctypes as C
            from geneakit_b200._lib import lib, check
            cls_p, ind_p, klast = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64()
            check(lib.gk_phi_final_view(job._h, C.byref(cls_p), C.byref(ind_p), C.byref(klast)))
            fcls = np.ctypeslib.as_array(cls_p, shape=(m,)).copy()
            find = np.ctypeslib.as_array(ind_p, shape=(m,)).copy()
            mine = A[np.ix_(fcls[r0:r1], fcls)]
            same = find[r0:r1, None] == find[None, :]
            mine = np.where(same, d[fcls[r0:r1]][:, None] * np.ones((1, m)), mine)
            assert mine.tobytes() == np.ascontiguousarray(phi[r0:r1]).tobytes()          # this rank's rows, bit-exact
            job.close()
            if m > 1:
                s_all = float(mine.sum())
                s_diag = float(sum(mine[i - r0, i] for i in range(r0, r1)))
                mean = combine_mean(s_all, s_diag, m)
                assert abs(mean - phi_mean(phi)) <= 1e-15
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: %r\n%s" % (e, traceback.format_exc())))
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
