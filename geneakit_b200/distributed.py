"""Multi-GPU driver: one process per GPU (torchrun), torch.distributed for the plumbing.

Sharding (DESIGN.md "Multi-GPU"): the class matrix of EVERY cut is sharded by blocks of rows
(= columns: the matrices are symmetric), in panels of 32 classes, one block per rank; the final
m x m matrix is sharded the same way by result rows.  A rank computes, for its own panels,
phase 1 from the parent rows -- read straight out of the owners' HBM by the step kernel (CUDA-IPC
mappings, peer loads over NVLink inside the kernel: the exchange is fused with the gather-average,
there is no separate pack / send / unpack) -- and phase 2 against every tile, storing only its own
rows.  The one collective per cut is the all-gather of the new classes' self kinships (8 bytes per
class: the "diagonal inputs" of the next cut), which also is the barrier that makes a cut's rows
readable by the peers' next step.  phiMean adds one all-reduce of two doubles.
"""
import ctypes as C
import os

import numpy as np

from . import cgeneakit
from ._lib import lib, check


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


class _DevArray:
    """A raw device pointer as a __cuda_array_interface__ object (so torch can wrap it, zero copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _dvec(job, g):
    p, chunk, total = C.c_void_p(), C.c_int64(), C.c_int64()
    check(lib.gk_phi_dvec(job._h, g, C.byref(p), C.byref(chunk), C.byref(total)))
    return p.value, chunk.value, total.value


class ShardedPhi:
    """gen.phi across the ranks of the default process group: `.local` rows stay in this rank's
    HBM, `.mean()` is the global gen.phiMean, `.to_host()` copies this rank's row block.

    The job's kernels and the per-cut collectives run on ONE torch stream owned by this object (the caller may pass its
    own with stream=<cudaStream_t handle of a non-default torch stream>): the all-gather after a cut is what orders the
    peers' reads of this rank's rows, so it must be enqueued behind the cut's kernels."""

    def __init__(self, pedigree, proband_ids=(), rank=None, world=None, stream=None, **gk):
        import torch.distributed as dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self._tstream = None
        if self.world > 1:
            import torch
            if stream is None:
                self._tstream = torch.cuda.Stream()
                stream = self._tstream.cuda_stream
            else:
                self._tstream = torch.cuda.ExternalStream(int(stream))
        self.job = cgeneakit.DevicePhi(pedigree, proband_ids, rank=self.rank, world=self.world, stream=stream, **gk)
        self.m = self.job.m
        self.row0, self.row1 = self.job.row0, self.job.row1
        self.G = self.job.stats()["n_cuts"]
        self._ranked = bool(self.job.stats()["ranked"])
        self._dv = None
        self._fence = None
        self._shared_plan = False

    def upload(self):
        """Allocate, then make every rank's matrix buffers addressable by every rank (CUDA IPC)."""
        import torch
        import torch.distributed as dist
        self.job.upload()
        if self.world > 1:
            mine = (C.c_ubyte * 128)()
            check(lib.gk_phi_ipc_export(self.job._h, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine))
            for q in range(self.world):
                if q != self.rank:
                    buf = (C.c_ubyte * 128).from_buffer_copy(handles[q])
                    check(lib.gk_phi_ipc_open(self.job._h, q, buf))
            check(lib.gk_phi_peers_ready(self.job._h))
            # self-kinship vectors of both parities as torch views (all-gathered in place after every cut)
            self._dv = []
            for g in range(self.G):
                p, chunk, total = _dvec(self.job, g)
                self._dv.append((torch.as_tensor(_DevArray(p, total), device="cuda"), chunk))
            self._fence = torch.zeros(1, dtype=torch.int32, device="cuda")
            # ONE planner per box: rank 0 maps every rank's plan segments (CUDA IPC) so that run_pass() can plan once and
            # copy each rank's view straight into that rank's HBM (GK_SHARED_PLAN=0: every rank plans for itself)
            self._shared_plan = os.environ.get("GK_SHARED_PLAN", "1") != "0"
            if self._shared_plan:
                nseg = C.c_int32()
                check(lib.gk_phi_seg_count(self.job._h, C.byref(nseg)))
                mine = []
                for g in range(nseg.value):
                    h, cap = (C.c_ubyte * 64)(), C.c_int64()
                    check(lib.gk_phi_seg_ipc_export(self.job._h, g, h, C.byref(cap)))
                    mine.append((bytes(h), cap.value))
                everybody = [None] * self.world
                dist.all_gather_object(everybody, mine)
                if self.rank == 0:
                    for q in range(1, self.world):
                        for g, (h, cap) in enumerate(everybody[q]):
                            check(lib.gk_phi_seg_ipc_open(self.job._h, q, g, (C.c_ubyte * 64).from_buffer_copy(h), cap))
            dist.barrier()
        return self

    def replan(self):
        """Plan again on the host + copy the plan to the device again (every rank plans the same pedigree)."""
        self.job.replan()
        return self

    def run_pass(self):
        """One whole pass from the flattened pedigree: the host plans the cuts in order while the device already runs
        the earlier ones (every rank plans the same pedigree, on its share of the host threads)."""
        return self.run(replan=True)

    def run(self, replan=False):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            if replan:
                self.job.run_pass()
            else:
                self.job.run()
            return self
        shared = replan and self._shared_plan
        with torch.cuda.stream(self._tstream):
            if shared:
                # rank 0 plans for everybody.  Its copies of cut g + 1 (into every rank's segments) are enqueued BEFORE the
                # all-gather that ends cut g, so that collective -- no rank leaves it before rank 0 has entered it -- orders
                # them before the peers' step g + 1; a fence collective does the same for cut 0.
                if self.rank == 0:
                    check(lib.gk_phi_pass_begin_all(self.job._h))
                    check(lib.gk_phi_pass_step(self.job._h, 0))
                dist.all_reduce(self._fence)
            elif replan:
                check(lib.gk_phi_pass_begin(self.job._h))
            for g in range(self.G):
                if replan and not shared:
                    check(lib.gk_phi_pass_step(self.job._h, g))     # cut g is planned and its copy enqueued
                check(lib.gk_phi_run_step(self.job._h, g))
                if shared and self.rank == 0 and g + 1 < self.G:
                    check(lib.gk_phi_pass_step(self.job._h, g + 1))
                full, chunk = self._dv[g]
                # the only per-cut collective: 8 bytes per class, and the barrier between cuts
                dist.all_gather_into_tensor(full, full[self.rank * chunk:(self.rank + 1) * chunk])
                if self._ranked and g > 0:
                    # deep pedigrees (rank-ordered rounding): tiles J <= I only, every rank stored into its own rows; now pull
                    # the mirror tiles out of the higher ranks' rows, and fence before anybody reads this cut's rows
                    check(lib.gk_phi_run_pull(self.job._h, g))
                    dist.all_reduce(self._fence)
            check(lib.gk_phi_finish(self.job._h))
            # the expansion read the peers' rows: nobody may start the next pass (or free) before everybody is done
            dist.all_reduce(self._fence)
            if replan and (not shared or self.rank == 0):
                check(lib.gk_phi_pass_end(self.job._h))
        return self

    def mean(self):
        s, d = self.job.sums()
        if self.world == 1:
            return (s - d) / (float(self.m) * float(self.m) - float(self.m))
        return combine_mean(s, d, self.m)

    def to_host(self, out=None):
        return self.job.to_host(out)

    def gather(self):
        """Full m x m matrix on every rank (host), for tests: all_gather of the row blocks."""
        import torch.distributed as dist
        mine = np.ascontiguousarray(self.to_host())
        parts = [None] * self.world
        dist.all_gather_object(parts, mine)
        return np.concatenate(parts, axis=0)

    def close(self):
        if self.world > 1 and self._dv is not None:
            import torch
            import torch.distributed as dist
            self.job.sync()
            torch.cuda.synchronize()
            dist.barrier()                 # a peer may still be reading this rank's buffers
        self._dv = None
        self.job.close()


class VirtualShardedPhi:
    """The sharded algorithm with all `world` ranks emulated in ONE process on ONE GPU: every
    rank's job runs its step back to back on ONE stream (rank 0's), peers are plain pointers, the all-gather
    is a device copy.  Used by the GPU tests (a fresh box has one GPU) and to study the sharded
    kernels without NVLink."""

    def __init__(self, pedigree, proband_ids=(), world=2, **gk):
        self.world = world
        self.jobs = [cgeneakit.DevicePhi(pedigree, proband_ids, rank=q, world=world, **gk) for q in range(world)]
        self.m = self.jobs[0].m
        self.G = self.jobs[0].stats()["n_cuts"]

    def upload(self):
        # ONE stream for every rank's job: the persistent step kernels of different ranks must run one after the other
        self.jobs[0].upload()
        st = C.c_void_p()
        check(lib.gk_phi_stream(self.jobs[0]._h, C.byref(st)))
        for j in self.jobs[1:]:
            j._opts.stream = st
            check(lib.gk_phi_set_stream(j._h, st))
            j.upload()
        bufs = []
        for j in self.jobs:
            b0, b1 = C.c_void_p(), C.c_void_p()
            check(lib.gk_phi_local_buffers(j._h, C.byref(b0), C.byref(b1)))
            bufs.append((b0, b1))
        for q, j in enumerate(self.jobs):
            for p in range(self.world):
                if p != q:
                    check(lib.gk_phi_set_peer(j._h, p, bufs[p][0], bufs[p][1]))
            check(lib.gk_phi_peers_ready(j._h))
        return self

    def run(self):
        for g in range(self.G):
            for j in self.jobs:
                check(lib.gk_phi_run_step(j._h, g))
            for j in self.jobs:
                j.sync()
            for j in self.jobs:
                for src in self.jobs:
                    check(lib.gk_phi_dvec_copy_from(j._h, src._h, g))
            for j in self.jobs:
                j.sync()
            for j in self.jobs:                # ranked jobs: the mirror tiles of the own rows, out of the higher ranks' rows
                check(lib.gk_phi_run_pull(j._h, g))
            for j in self.jobs:
                j.sync()
        for j in self.jobs:
            check(lib.gk_phi_finish(j._h))
        return self

    def mean(self):
        s = d = 0.0
        for j in self.jobs:
            a, b = j.sums()
            s += a
            d += b
        return (s - d) / (float(self.m) * float(self.m) - float(self.m))

    def to_host(self):
        return np.concatenate([np.asarray(j.to_host()) for j in self.jobs], axis=0)

    def close(self):
        for j in reversed(self.jobs):          # rank 0 owns the stream the others enqueue on: it goes last
            j.close()


def shard_ancestors(n_anc, rank, world):
    """gen.gc shards by ANCESTOR COLUMN with no exchange at all (every ancestor's propagation is
    independent): columns [a0, a1) of the m x a matrix for `rank`."""
    return n_anc * rank // world, n_anc * (rank + 1) // world


def sharded_gc(pedigree, proband_ids, ancestor_ids, rank, world, **gk):
    """This rank's column block of gen.gc (rows = probands, columns = its slice of the ancestors),
    computed by the same C-ABI call on the sliced ancestor list."""
    a0, a1 = shard_ancestors(len(ancestor_ids), rank, world)
    if a1 <= a0:
        return np.zeros((len(proband_ids), 0), dtype=np.float64), (a0, a1)
    block = cgeneakit.compute_genetic_contributions(pedigree, proband_ids, list(ancestor_ids[a0:a1]), **gk)
    return block, (a0, a1)
