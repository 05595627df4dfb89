#!/usr/bin/env python
"""Benchmark of the kinship hot path: gen.phi (+ fused gen.phiMean) on BASELINE.json's synthetic
pedigrees.  One "step" = one full pass of the hot path (all generations + the final expansion)
over one pedigree / proband list.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c5s|tiny] [--impl reference]

Prints ONE JSON line (contract in the task statement): metric = kinship pairs/s, with
`roofline` (dominant kernel, algorithmic bytes / CUDA-event time / measured HBM peak),
`cpu_baseline` (the unmodified reference on a bounded sample of the same workload),
`e2e` (the same metric through the host-buffer C-ABI call gk_phi_f64, copies included),
`clocks` and `gpu_launches`.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (individuals, generations, probands, seed, founders)  -- SURVEY.md 8(d)
    "c4": dict(n=2_000_000, gens=20, m=100_000, seed=4, founders=None,
               desc="synthetic 2M-individual, 20-generation pedigree, 100k probands (BASELINE configs[3])"),
    "c3": dict(n=200_000, gens=12, m=10_000, seed=3, founders=None,
               desc="synthetic 200k-individual, 12-generation pedigree, 10k probands (BASELINE configs[2])"),
    "c5": dict(n=5_000_000, gens=30, m=50_000, seed=5, founders=2000,
               desc="synthetic 5M-individual, 30-generation founder-population pedigree, 50k probands (BASELINE configs[4]; needs 8 GPUs)"),
    "c5s": dict(n=1_250_000, gens=30, m=12_500, seed=5, founders=500,
                desc="1/4-scale founder-population pedigree, 30 generations (BASELINE configs[4] shape)"),
    "tiny": dict(n=20_000, gens=8, m=1_500, seed=3, founders=None, desc="tiny synthetic (debug)"),
}
SAMPLE_DIV = 20   # the CPU reference runs a 1/20-width replica of the workload (same generator, same depth)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], threading.Event()

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 7 and r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_workload(name, div=1):
    import geneakit_b200 as gk
    w = WORKLOADS[name]
    n, m = w["n"] // div, w["m"] // div
    founders = None if w["founders"] is None else max(4, w["founders"] // div)
    t0 = time.time()
    ped, pro = gk.synthetic_pedigree(n, w["gens"], m, seed=w["seed"], founders=founders)
    return ped, pro, time.time() - t0


def reference_handle(ped):
    """The same pedigree inside the unmodified reference (oracle/_ref) or, if that library was not
    built in this tree, the C port of it (oracle/libgkoracle.so)."""
    from oracle.api import Oracle, Reference
    f = np.where(ped.sire >= 0, ped.ids[np.maximum(ped.sire, 0)], 0)
    d = np.where(ped.dam >= 0, ped.ids[np.maximum(ped.dam, 0)], 0)
    if Reference.available():
        R = Reference()
        return "reference", R, R.genealogy(ped.ids, f, d, ped.sex, sorted=True), R.L.gkref_max_threads()
    O = Oracle()
    return "port", O, O.genealogy(ped.ids, f, d, ped.sex, sorted=True), O.L.gko_max_threads()


def cpu_sample(workload, repeats=1):
    ped, pro, _ = make_workload(workload, SAMPLE_DIV)
    kind, eng, h, cores = reference_handle(ped)
    best = None
    for _ in range(repeats):
        t0 = time.time()
        eng.phi(h, pro)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    m = len(pro)
    w = WORKLOADS[workload]
    return {"value": m * (m + 1) / 2 / best, "unit": "pairs/s", "cores": int(cores), "kind": kind,
            "seconds": best,
            "sample": "1/%d-width replica of %s: %d individuals, %d generations, %d probands, full gen.phi"
                      % (SAMPLE_DIV, workload, w["n"] // SAMPLE_DIV, w["gens"], m)}


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ped, pro, _ = make_workload(args.workload, SAMPLE_DIV)
    kind, eng, h, cores = reference_handle(ped)
    m = len(pro)
    for _ in range(args.warmup):
        eng.phi(h, pro)
    t0 = time.time()
    for _ in range(args.steps):
        eng.phi(h, pro)
    dt = (time.time() - t0) / max(args.steps, 1)
    val = m * (m + 1) / 2 / dt
    w = WORKLOADS[args.workload]
    sample = ("1/%d-width replica of %s: %d individuals, %d generations, %d probands, full gen.phi"
              % (SAMPLE_DIV, args.workload, w["n"] // SAMPLE_DIV, w["gens"], m))
    line = {"impl": "reference", "metric": "kinship pairs/sec for gen.phi", "value": val, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": w["desc"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": int(cores), "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out.write(json.dumps(line) + "\n")
    out.flush()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else a library prints there (NCCL's version banner,
    torchrun notices) goes to stderr.  Returns a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compress", type=int, default=-1)
    ap.add_argument("--order", type=int, default=-1)
    args = ap.parse_args()
    out = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import geneakit_b200 as gk
    from geneakit_b200 import _lib
    from geneakit_b200._lib import lib, check
    from geneakit_b200.distributed import ShardedPhi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ped, pro, gen_s = make_workload(args.workload)
    m = len(pro)
    stream = torch.cuda.Stream()          # a real stream: the default stream's handle is NULL
    torch.cuda.set_stream(stream)
    sp = ShardedPhi(ped, pro, rank=rank, world=world, compress=args.compress, order=args.order, device=local,
                    stream=stream.cuda_stream)
    sp.upload()                                    # plan + buffers resident in HBM before timing
    job = sp.job
    st = job.stats()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        sp.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        sp.run()
    e1.record(stream)
    barrier()
    sampler.stop.set()
    ms = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)             # device time, max over ranks
        ms = float(t[0])
    mean = sp.mean()
    pairs = m * (m + 1) / 2
    value = pairs / (ms * 1e-3)

    # dominant kernel: phi_step_kernel, one launch per generation; CUDA events recorded by the
    # library on this same stream around every launch of the LAST timed pass (this rank's launches)
    t_ms = job.step_times_ms()
    b = job.step_bytes()
    nstep = st["n_steps"]
    step_ms, step_bytes = sum(t_ms[:nstep]), sum(b[:nstep]) / world      # a rank produces 1/world of every cut's matrix
    peak, peak_src = peaks()
    achieved = step_bytes / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "phi_step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "launches": nstep, "bytes_per_launch_avg": step_bytes / max(nstep, 1),
                "ms_per_launch_avg": step_ms / max(nstep, 1), "share_of_step": step_ms / ms,
                "expand_ms": t_ms[nstep] if len(t_ms) > nstep else None,
                "expand_gbs": (b[nstep] / world / (t_ms[nstep] * 1e-3) / 1e9) if len(t_ms) > nstep and t_ms[nstep] > 0 else None}
    rank_ms = None
    if dist is not None:
        # every rank's step-kernel and expansion time of the last pass (imbalance shows up here)
        mine = torch.tensor([step_ms, t_ms[nstep] if len(t_ms) > nstep else 0.0], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [[round(float(x[0]), 3), round(float(x[1]), 3)] for x in allr]
    if world > 1:
        roofline["rank_step_expand_ms"] = rank_ms
        roofline["note"] = "per GPU (rank 0): 1/%d of every cut's algorithmic bytes over this rank's launch time" % world
        roofline["nvlink_bytes_per_pass_rank0"] = int(st["rank_remote_row_bytes"])
        roofline["nvlink_gbs_rank0"] = st["rank_remote_row_bytes"] / (step_ms * 1e-3) / 1e9 if step_ms > 0 else None
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tr) and world == 1:
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
        except Exception:
            pass

    # end to end through the host-buffer entry points (plan + H2D of the plan + kernels + D2H of the
    # result rows inside the timed region).  One GPU: the one-shot C-ABI call gk_phi_f64.
    e2e = None
    if args.e2e_steps > 0:
        rows = job.row1 - job.row0
        nbytes = rows * m * 8
        ptr = lib.gk_host_alloc(nbytes)
        pinned = bool(ptr)
        if not pinned:
            hostbuf = np.empty((rows, m), dtype=np.float64)
            ptr = hostbuf.ctypes.data
        host = np.frombuffer((C.c_double * (rows * m)).from_address(ptr), dtype=np.float64).reshape(rows, m)
        r = ped.ranks(pro)
        mu = C.c_double()
        dts = []
        for it in range(args.e2e_steps + 1):
            barrier()
            t0 = time.time()
            if world == 1:
                opts = _lib.make_opts(device=local, compress=args.compress, order=args.order)
                check(lib.gk_phi_f64(len(ped), ped.sire, ped.dam, m, r, ptr, C.byref(mu), C.byref(opts)))
                got = mu.value
            else:
                j2 = ShardedPhi(ped, pro, rank=rank, world=world, compress=args.compress, order=args.order,
                                device=local, stream=stream.cuda_stream)
                j2.upload().run()
                j2.to_host(host)
                got = j2.mean()
                j2.close()
            barrier()
            if it > 0:
                dts.append(time.time() - t0)
        assert got == mean, (got, mean)
        dt = sum(dts) / len(dts)
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {"value": pairs / dt, "unit": "pairs/s", "seconds_per_step": dt,
               "h2d_bytes_per_step": int(st["plan_bytes"]) * world,
               "d2h_bytes_per_step": int(m) * int(m) * 8 + 16 * world, "host_buffer": "pinned" if pinned else "pageable",
               "plan_seconds": st["plan_seconds"], "steps": args.e2e_steps}
        if pinned:
            lib.gk_host_free(ptr)

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cpu = cpu_sample(args.workload)

    if rank == 0:
        sampler.join()
        w = WORKLOADS[args.workload]
        sharding = ("one GPU" if world == 1 else
                    "rows of every cut's class matrix and of the result sharded %d-way; parent rows read from the owners' "
                    "HBM over NVLink inside the step kernel; collectives: all-gather of the self-kinship vector per cut "
                    "+ all-reduce of 2 doubles (phiMean)" % world)
        line = {"metric": "kinship pairs/sec for gen.phi", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "individuals": len(ped), "generations": w["gens"], "probands": m,
                           "seed": w["seed"], "cuts": job.cut_sizes(), "classes": job.class_sizes(),
                           "compressed": bool(st["compressed"]), "ranked": bool(st["ranked"]), "panel": st["tile"],
                           "sharding": sharding, "rows_per_rank": job.row1 - job.row0,
                           "l2": "inputs larger than L2 (%.1f GB of matrices per pass)" % (st["algorithmic_bytes_device"] / 1e9),
                           "cells_reference": st["cells_reference"], "cells_device": st["cells_device"],
                           "parent_row_bytes_per_pass": st["panel_row_bytes"], "corrections": st["corrections"],
                           "phi_mean": mean, "pedigree_build_s": gen_s},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": sampler.summary(),
                "gpu_launches": int(st["kernel_launches"]) * args.steps * world,
                "reference_algorithmic_gbs": st["algorithmic_bytes_reference"] / (ms * 1e-3) / 1e9}
        out.write(json.dumps(line) + "\n")
        out.flush()
    sp.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
