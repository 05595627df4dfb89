#!/usr/bin/env python
"""Benchmark of the kinship hot path: gen.phi (+ fused gen.phiMean) on BASELINE.json's configurations.
One "step" = one full pass of the hot path from the flattened pedigree arrays ON THE HOST to the m x m matrix and its
mean resident in HBM (SURVEY.md 8d's t_phi): host planning, the copy of the plan to the device, all generations, the
final expansion.  The buffers are allocated once, before the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c1|c2|c5|c5s|tiny] [--op phi|gc] [--impl reference]

Prints ONE JSON line (contract in the task statement): metric = kinship pairs/s, with
  `resident`     the same pass with the plan already in HBM (kernels only),
  `roofline`     dominant kernel: algorithmic bytes / CUDA-event time / measured HBM peak,
  `parity_check` after the timed region every rank compares result rows with an independent one-GPU job on a sub-list,
  `cpu_baseline` the unmodified reference on a bounded sample of the same workload, all host threads,
  `e2e`          the same metric through the host-buffer C-ABI call gk_phi_f64 (N > 1: ONE process driving the N GPUs),
  `clocks`, `gpu_launches`.
"""
import argparse
import ctypes as C
import importlib.util
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
    # the reference's bundled datasets, from the fixtures the unmodified reference produced (tests/golden)
    "c1": dict(fixture="genea140", desc="genea140 bundled dataset: gen.phi(ped, pro=gen.pro(ped)) + gen.phiMean (BASELINE configs[0])"),
    "c2": dict(fixture="genea140_pop140", desc="pop140 bundled dataset: gen.phi on the 140 probands of pop140 + gen.gc of the founders (BASELINE configs[1])"),
}
SAMPLE_DIV = 10   # the CPU reference runs a 1/10-width replica of the synthetic workloads (same generator, same depth)


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


def _synth():
    """geneakit_b200/synth.py loaded by path: numpy only, so the reference arm never loads the product library."""
    spec = importlib.util.spec_from_file_location("gk_synth_numpy", os.path.join(ROOT, "geneakit_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_arrays(name, div=1):
    """(ids, fathers, mothers, sexes, proband ids, ancestor ids for gc) -- numpy / lists only."""
    w = WORKLOADS[name]
    if "fixture" in w:
        z = np.load(os.path.join(ROOT, "tests", "golden", w["fixture"] + ".npz"))
        r = z["rows"]
        return r[:, 0], r[:, 1], r[:, 2], r[:, 3], [int(x) for x in z["pro"]], [int(x) for x in z["anc"]], bool(z["sorted"])
    s = _synth()
    n, m = w["n"] // div, w["m"] // div
    founders = None if w["founders"] is None else max(4, w["founders"] // div)
    ids, f, mo, sx, offs = s.synthetic_arrays(n, w["gens"], w["seed"], founders)
    pro = s.synthetic_probands(ids, offs, m, w["seed"])
    anc = [int(x) for x in ids[:offs[1]][:2000]]          # gc: (up to 2000 of) the founders
    return ids, f, mo, sx, pro, anc, True


def make_workload(name, div=1):
    import geneakit_b200 as gk
    t0 = time.time()
    ids, f, mo, sx, pro, _, srt = workload_arrays(name, div)
    ped = gk.cgeneakit.load_pedigree_from_vectors(ids, f, mo, sx, srt)
    return ped, pro, time.time() - t0


def reference_engine():
    """The unmodified reference (oracle/_ref) or, if that library was not built in this tree, the C port of it
    (oracle/libgkoracle.so) -- with ALL host threads, whatever OMP_NUM_THREADS the launcher exported."""
    from oracle.api import Oracle, Reference
    threads = os.cpu_count() or 1
    if Reference.available():
        R = Reference()
        R.set_threads(threads)
        return "reference", R, int(R.L.gkref_max_threads())
    O = Oracle()
    O.set_threads(threads)
    return "port", O, int(O.L.gko_max_threads())


def sample_text(name, div, n_ind, m):
    w = WORKLOADS[name]
    if "fixture" in w:
        return "%s in full: %d individuals, %d probands, full gen.phi" % (name, n_ind, m)
    return ("1/%d-width replica of %s: %d individuals, %d generations, %d probands, full gen.phi"
            % (div, name, n_ind, w["gens"], m))


def cpu_sample(name):
    kind, eng, cores = reference_engine()
    div = 1 if "fixture" in WORKLOADS[name] else SAMPLE_DIV
    ids, f, mo, sx, pro, _, srt = workload_arrays(name, div)
    h = eng.genealogy(ids, f, mo, sx, sorted=srt)
    t0 = time.time()
    eng.phi(h, pro)
    dt = time.time() - t0
    m = len(pro)
    return {"value": m * (m + 1) / 2 / dt, "unit": "pairs/s", "cores": cores, "kind": kind, "seconds": dt,
            "sample": sample_text(name, div, len(ids), m)}


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample.  Never imports the product."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind, eng, cores = reference_engine()
    div = 1 if "fixture" in WORKLOADS[args.workload] else SAMPLE_DIV
    ids, f, mo, sx, pro, _, srt = workload_arrays(args.workload, div)
    h = eng.genealogy(ids, f, mo, sx, sorted=srt)
    t0 = time.time()
    eng.phi(h, pro)
    first = time.time() - t0
    if first > 15.0 and div == SAMPLE_DIV:        # a small host: keep the whole run within a few minutes
        div = 2 * SAMPLE_DIV
        ids, f, mo, sx, pro, _, srt = workload_arrays(args.workload, div)
        h = eng.genealogy(ids, f, mo, sx, sorted=srt)
    m = len(pro)
    for _ in range(max(args.warmup - 1, 0)):
        eng.phi(h, pro)
    t0 = time.time()
    for _ in range(args.steps):
        eng.phi(h, pro)
    dt = (time.time() - t0) / max(args.steps, 1)
    val = m * (m + 1) / 2 / dt
    sample = sample_text(args.workload, div, len(ids), m)
    line = {"impl": "reference", "metric": "kinship pairs/sec for gen.phi", "value": val, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOADS[args.workload]["desc"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_library_loaded": "geneakit_b200" in sys.modules}
    out.write(json.dumps(line) + "\n")
    out.flush()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else a library prints there (NCCL's version banner,
    torchrun notices) goes to stderr.  Returns a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def parity_check(sp, ped, pro, local, rows_per_rank=48):
    """Result rows of THIS rank (after the timed passes) against an independent one-GPU job on a sub-list of the probands:
    kinship of a sub-list = the sub-matrix (other cuts, other classes, other kernels' tiles), bit for bit."""
    import geneakit_b200 as gk
    m = len(pro)
    r0, r1 = sp.row0, sp.row1
    own = np.unique(np.linspace(r0, r1 - 1, num=min(rows_per_rank, r1 - r0)).astype(np.int64))
    cols = np.unique(np.linspace(0, m - 1, num=min(rows_per_rank, m)).astype(np.int64))
    sub = np.unique(np.concatenate([own, cols]))
    small = gk.cgeneakit.compute_kinships(ped, [pro[i] for i in sub], False, device=local)
    pos = {int(i): k for k, i in enumerate(sub)}
    ok = True
    for i in own:
        row = sp.job.rows(int(i), 1)[0]
        ok = ok and row[sub].tobytes() == small[pos[int(i)]].tobytes()
    return int(len(own)), int(len(sub)), bool(ok)


def run_gc(args, out, torch, dist, rank, world, local):
    """--op gc: gen.gc of the workload's founders (up to 2000) to its probands, ancestor columns sharded over the ranks
    (no exchange), each rank through the one-shot C-ABI call gk_gc_f64 (host buffers, copies included)."""
    import geneakit_b200 as gk
    from geneakit_b200 import _lib
    from geneakit_b200.distributed import shard_ancestors
    ped, pro, _ = make_workload(args.workload)
    anc = workload_arrays(args.workload)[5]
    a0, a1 = shard_ancestors(len(anc), rank, world)
    mine = anc[a0:a1]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    st = _lib.gk_gc_stats()
    for _ in range(max(args.warmup, 1)):
        gk.cgeneakit.compute_genetic_contributions(ped, pro, mine, device=local)
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        block = gk.cgeneakit.compute_genetic_contributions(ped, pro, mine, device=local)
    barrier()
    dt = (time.time() - t0) / args.steps
    _lib.lib.gk_gc_last_stats(C.byref(st))
    vals = torch.tensor([dt, st.kernel_ms, float(block.sum())], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, kms, total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        kms, total = st.kernel_ms, float(block.sum())
    if rank == 0:
        peak, peak_src = peaks()
        cells = len(pro) * len(anc)
        ach = st.algorithmic_bytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        line = {"metric": "genetic contributions/sec for gen.gc", "value": cells / dt, "unit": "proband-ancestor cells/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload]["desc"], "op": "gen.gc", "probands": len(pro), "ancestors": len(anc),
                           "sharding": "ancestor columns, no exchange", "sum_of_contributions": total,
                           "timed": "gk_gc_f64 from host arrays to a host matrix (plan, H2D, kernels, D2H)"},
                "roofline": {"bound": "hbm", "kernel": "gc_step_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": None, "peak_source": peak_src, "note": "rank 0: 24 B x live members x its ancestor columns / kernel time",
                             "kernel_ms": kms, "column_blocks": int(st.column_blocks), "launches": int(st.kernel_launches)},
                "cpu_baseline": None, "e2e": {"value": cells / dt, "unit": "proband-ancestor cells/s",
                                              "h2d_bytes_per_step": int(8 * len(ped) + 4 * len(pro)), "d2h_bytes_per_step": int(8 * cells)},
                "gpu_launches": int(st.kernel_launches) * args.steps * world}
        out.write(json.dumps(line) + "\n")
        out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--op", default="phi", choices=["phi", "gc"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
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
    dist = host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")        # host-side barriers: no kernel may sit on a GPU another process drives
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    if args.op == "gc":
        run_gc(args, out, torch, dist, rank, world, local)
        if dist is not None:
            dist.destroy_process_group()
        return

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=host_group)

    ped, pro, gen_s = make_workload(args.workload)
    m = len(pro)
    stream = torch.cuda.Stream()          # a real stream: the default stream's handle is NULL
    torch.cuda.set_stream(stream)
    sp = ShardedPhi(ped, pro, rank=rank, world=world, compress=args.compress, order=args.order, device=local,
                    stream=stream.cuda_stream)
    sp.upload()                                    # buffers allocated, peers mapped: not part of a pass
    job = sp.job
    warm = max(args.warmup, 3)
    for _ in range(warm):
        sp.run_pass()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        sp.run_pass()                              # host planning + H2D of the plan + every kernel of the pass (overlapped inside the pass)
        stream.synchronize()                       # a pass ends before the next one is planned (no overlap across passes)
    e1.record(stream)
    barrier()
    sampler.stop.set()
    ms = e0.elapsed_time(e1) / args.steps
    st = job.stats()
    # the same pass with the plan resident in HBM (kernels only)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0.record(stream)
    for _ in range(3):
        sp.run()
    r1.record(stream)
    barrier()
    ms_res = r0.elapsed_time(r1) / 3
    if dist is not None:
        t = torch.tensor([ms, ms_res], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)             # device time, max over ranks
        ms, ms_res = float(t[0]), float(t[1])
    mean = sp.mean()
    pairs = m * (m + 1) / 2
    value = pairs / (ms * 1e-3)

    # dominant kernel: the step kernel, one launch per generation; CUDA events recorded by the library on this same
    # stream around every launch of the LAST pass (this rank's launches)
    t_ms = job.step_times_ms()
    b = job.step_bytes()
    nstep = st["n_steps"]
    step_ms, step_bytes = sum(t_ms[:nstep]), sum(b[:nstep]) / world      # a rank produces 1/world of every cut's matrix
    peak, peak_src = peaks()
    achieved = step_bytes / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
    kernel = "phi_step_g4_kernel" if max(job.class_sizes()) >= 8192 else ("phi_step_tma_kernel" if world > 1 else "phi_step_kernel")
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "launches": nstep, "bytes_per_launch_avg": step_bytes / max(nstep, 1),
                "ms_per_launch_avg": step_ms / max(nstep, 1), "share_of_step": step_ms / ms_res,
                "expand_ms": t_ms[nstep] if len(t_ms) > nstep else None,
                "expand_gbs": (b[nstep] / world / (t_ms[nstep] * 1e-3) / 1e9) if len(t_ms) > nstep and t_ms[nstep] > 0 else None}
    if dist is not None:
        # every rank's step-kernel and expansion time of the last pass (imbalance shows up here) and its NVLink rate
        mine = torch.tensor([step_ms, t_ms[nstep] if len(t_ms) > nstep else 0.0,
                             st["rank_remote_row_bytes"] / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        roofline["rank_step_expand_ms"] = [[round(float(x[0]), 3), round(float(x[1]), 3)] for x in allr]
        roofline["nvlink_gbs"] = [round(float(x[2]), 1) for x in allr]
        roofline["note"] = "per GPU (rank 0): 1/%d of every cut's algorithmic bytes over this rank's launch time" % world
        roofline["nvlink_bytes_per_pass_rank0"] = int(st["rank_remote_row_bytes"])
    tr = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tr) and world == 1:
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            roofline["traffic_source"] = "profiles/traffic_%s.json (ncu --set full of this kernel, per launch)" % args.workload
        except Exception:
            pass

    # correctness of what was just timed, on every rank
    pc = None
    if not args.no_parity_check:
        n_rows, n_cols, ok = parity_check(sp, ped, pro, local)
        flags = torch.tensor([n_rows, 1 if ok else 0], dtype=torch.int64, device="cuda")
        if dist is not None:
            allf = [torch.zeros_like(flags) for _ in range(world)]
            dist.all_gather(allf, flags)
        else:
            allf = [flags]
        pc = {"rows": int(sum(int(x[0]) for x in allf)), "columns_per_row": n_cols, "ranks": world,
              "bit_exact": bool(all(int(x[1]) == 1 for x in allf)),
              "against": "independent one-GPU job on the sub-list of probands (sub-list kinship = sub-matrix)"}
    plan_bytes = int(st["plan_bytes"])
    launches = int(st["kernel_launches"])
    cuts, classes = job.cut_sizes(), job.class_sizes()
    rows_per_rank = job.row1 - job.row0
    sp.close()
    lib.gk_cache_release()                          # the end-to-end leg allocates for itself (rank 0 drives every GPU)
    host_barrier()

    # end to end through the host-buffer entry point (plan + H2D of the plan + kernels + D2H of the m x m result inside
    # the timed region): the one-shot C-ABI call gk_phi_f64 from ONE process -- rank 0 drives all `world` GPUs
    e2e = None
    if args.e2e_steps > 0:
        if rank == 0:
            nbytes = m * m * 8
            ptr = lib.gk_host_alloc(nbytes)
            pinned = bool(ptr)
            if not pinned:
                hostbuf = np.empty((m, m), dtype=np.float64)
                ptr = hostbuf.ctypes.data
            r = ped.ranks(pro)
            mu = C.c_double()
            dts = []
            for it in range(args.e2e_steps + 1):
                opts = _lib.make_opts(device=local, compress=args.compress, order=args.order, gpus=world)
                t0 = time.time()
                check(lib.gk_phi_f64(len(ped), ped.sire, ped.dam, m, r, ptr, C.byref(mu), C.byref(opts)))
                if it > 0:
                    dts.append(time.time() - t0)
            got = mu.value
            tm = np.zeros(6)
            lib.gk_phi_last_timing(tm)
            assert abs(got - mean) <= 1e-12 * abs(mean), (got, mean)
            dt = sum(dts) / len(dts)
            e2e = {"value": pairs / dt, "unit": "pairs/s", "seconds_per_step": dt,
                   "h2d_bytes_per_step": plan_bytes * world,
                   "d2h_bytes_per_step": int(m) * int(m) * 8 + 16 * world, "host_buffer": "pinned" if pinned else "pageable",
                   "driver": "one process, gk_phi_f64 with opts.world = %d" % world if world > 1 else "gk_phi_f64",
                   "phi_mean": got, "steps": args.e2e_steps,
                   "last_call_seconds": {"plan": tm[0], "alloc_and_h2d": tm[1], "kernels": tm[2], "d2h": tm[3], "total": tm[4]}}
            lib.gk_cache_release()
            if pinned:
                lib.gk_host_free(ptr)
        host_barrier()

    cpu = None
    if not args.no_cpu_baseline and rank == 0:
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
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic" if "fixture" not in w else "bundled dataset (reference-generated fixture)",
                "config": {"workload": w["desc"], "individuals": len(ped), "generations": w.get("gens"), "probands": m,
                           "seed": w.get("seed"), "cuts": cuts, "classes": classes,
                           "compressed": bool(st["compressed"]), "ranked": bool(st["ranked"]), "panel": st["tile"],
                           "sharding": sharding, "rows_per_rank": rows_per_rank,
                           "timed_region": "per step: host planning + H2D of the plan + all cuts + expansion + mean (SURVEY 8d t_phi); buffers pre-allocated",
                           "l2": "inputs larger than L2 (%.1f GB of matrices per pass)" % (st["algorithmic_bytes_device"] / 1e9),
                           "cells_reference": st["cells_reference"], "cells_device": st["cells_device"],
                           "parent_row_bytes_per_pass": st["panel_row_bytes"], "corrections": st["corrections"],
                           "phi_mean": mean, "plan_seconds": st["plan_seconds"], "pedigree_build_s": gen_s},
                "resident": {"value": pairs / (ms_res * 1e-3), "unit": "pairs/s", "ms_per_step": ms_res,
                             "note": "plan already in HBM: kernels of one pass only"},
                "roofline": roofline, "parity_check": pc, "cpu_baseline": cpu, "e2e": e2e, "clocks": sampler.summary(),
                "gpu_launches": launches * (warm + args.steps + 3) * world,
                "reference_algorithmic_gbs": st["algorithmic_bytes_reference"] / (ms_res * 1e-3) / 1e9}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
