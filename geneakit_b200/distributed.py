"""Multi-GPU driver: one process per GPU (torchrun), torch.distributed for the plumbing.

Sharding (DESIGN.md "Multi-GPU"): the final m x m kinship matrix is sharded by contiguous blocks
of result rows (= columns: the matrix is symmetric), one block per rank -- each rank holds
m/P x m doubles in its own HBM and copies only that block to the host.  The per-generation
recurrence runs replicated on every rank (with sibship compression it is 3-4x smaller than the
result matrix, and sharding it needs, every generation, the other parent's rows from remote
shards: ~1/2 shard of NVLink traffic per generation against ~2 shards of HBM traffic, i.e.
NVLink-bound at 8 GPUs -- SURVEY.md 8e).  The only data-path collective is therefore the
all-reduce of the two phiMean partial sums.
"""
import ctypes as C

import numpy as np

from . import cgeneakit
from ._lib import lib


def shard_rows(m, rank, world):
    """Result rows [row0, row1) owned by `rank` (gk_shard_rows in the C ABI)."""
    r0, r1 = C.c_int64(), C.c_int64()
    lib.gk_shard_rows(int(m), int(rank), int(world), C.byref(r0), C.byref(r1))
    return r0.value, r1.value


def combine_mean(sum_all, sum_diag, m, group=None):
    """All-reduce this rank's partial (sum of all entries, sum of the diagonal) of its row block
    and apply gen.phiMean's formula (reference compute.py:131-132).  Works on NCCL (device
    tensors) and on gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([sum_all, sum_diag], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    total, diag = float(t[0]), float(t[1])
    return (total - diag) / (float(m) * float(m) - float(m))


class ShardedPhi:
    """gen.phi across the ranks of the default process group: `.local` is this rank's row block
    of the kinship matrix (device resident), `.mean()` the global gen.phiMean."""

    def __init__(self, pedigree, proband_ids=(), rank=None, world=None, **gk):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.job = cgeneakit.DevicePhi(pedigree, proband_ids, rank=self.rank, world=self.world, **gk)
        self.m = self.job.m
        self.row0, self.row1 = self.job.row0, self.job.row1

    def upload(self):
        self.job.upload(); return self

    def run(self):
        self.job.run(); return self

    def mean(self):
        s, d = self.job.sums()
        return combine_mean(s, d, self.m)

    def to_host(self, out=None):
        return self.job.to_host(out)

    def gather(self):
        """Full m x m matrix on every rank (host), for tests: all_gather of the row blocks."""
        import torch
        import torch.distributed as dist
        mine = torch.from_numpy(np.ascontiguousarray(self.to_host()))
        parts = [None] * self.world
        dist.all_gather_object(parts, mine.numpy())
        return np.concatenate(parts, axis=0)

    def close(self):
        self.job.close()


def bench_sharded(args, rank, world, local):
    """bench.py body for N > 1 (launched by torch.distributed.run, one rank per GPU)."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from bench import WORKLOADS, ClockSampler, make_workload, peaks

    ped, pro, gen_s = make_workload(args.workload)
    m = len(pro)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = ShardedPhi(ped, pro, rank=rank, world=world, compress=args.compress, device=local,
                    stream=stream.cuda_stream)
    sp.upload()
    st = sp.job.stats()
    for _ in range(max(args.warmup, 3)):
        sp.run()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        sp.run()
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)             # device time, max over ranks
    ms = float(t[0])
    mean = sp.mean()
    sampler.stop.set()
    pairs = m * (m + 1) / 2

    # end to end: plan + upload + kernels + device->host copy of this rank's rows, every rank in parallel
    e2e = None
    if args.e2e_steps > 0:
        rows = sp.row1 - sp.row0
        ptr = lib.gk_host_alloc(rows * m * 8)
        host = np.frombuffer((C.c_double * (rows * m)).from_address(ptr), dtype=np.float64).reshape(rows, m) if ptr \
            else np.empty((rows, m), dtype=np.float64)
        dts = []
        for it in range(args.e2e_steps + 1):
            dist.barrier()
            t0 = time.time()
            j = ShardedPhi(ped, pro, rank=rank, world=world, compress=args.compress, device=local)
            j.upload().run()
            j.to_host(host)
            mu = j.mean()
            j.close()
            dist.barrier()
            if it > 0:
                dts.append(time.time() - t0)
        assert mu == mean, (mu, mean)
        dt = torch.tensor([sum(dts) / len(dts)], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs / float(dt[0]), "unit": "pairs/s", "seconds_per_step": float(dt[0]),
               "h2d_bytes_per_step": int(st["plan_bytes"]) * world, "d2h_bytes_per_step": int(m) * int(m) * 8 + 16 * world,
               "host_buffer": "pinned" if ptr else "pageable", "steps": args.e2e_steps}
        if ptr:
            lib.gk_host_free(ptr)

    if rank == 0:
        sampler.join()
        t_ms, b = sp.job.step_times_ms(), sp.job.step_bytes()
        nstep = st["n_steps"]
        step_ms, step_bytes = sum(t_ms[:nstep]), sum(b[:nstep])
        peak, peak_src = peaks()
        achieved = step_bytes / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
        w = WORKLOADS[args.workload]
        line = {"metric": "kinship pairs/sec for gen.phi", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": w["desc"], "individuals": len(ped), "generations": w["gens"], "probands": m,
                           "seed": w["seed"], "sharding": "result rows sharded %d-way; recurrence replicated per rank; "
                           "collective = all-reduce of 2 doubles (phiMean)" % world,
                           "rows_per_rank": sp.row1 - sp.row0, "compressed": bool(st["compressed"]),
                           "l2": "inputs larger than L2", "phi_mean": mean},
                "roofline": {"bound": "hbm", "kernel": "phi_step_kernel", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                             "note": "per GPU (rank 0); every rank runs the same recurrence launches"},
                "cpu_baseline": None, "e2e": e2e, "clocks": sampler.summary(),
                "gpu_launches": int(st["kernel_launches"]) * args.steps * world}
        print(json.dumps(line))
    sp.close()
    dist.destroy_process_group()
