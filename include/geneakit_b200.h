/* geneakit_b200 -- C ABI of the B200-native dense-kinship hot path.
 *
 * This is the drop-in boundary behind the reference's `cgeneakit` binding.  Every entry point
 * takes plain pointers and sizes (no C++/torch types) and names the reference interface it
 * replaces (paths relative to the reference repository, geneakit/...).
 *
 * Index space: RANK space.  Individual k is the k-th entry of Pedigree::ids, which the
 * reference keeps topologically ordered with Individual::rank == k (include/pedigree.hpp:34-44).
 * sire[k] / dam[k] = rank of the father / mother, or -1 when unknown.  The binding flattens
 * its Pedigree once and validates ids (so IndexError behaviour is unchanged).
 *
 * Errors: no exception crosses this boundary.  0 = GK_OK, otherwise a GK_ERR_* code; the
 * message is available from gk_last_error() (thread-local).  There is NO CPU fallback:
 * without a usable CUDA device every compute entry point returns GK_ERR_CUDA.
 */
#ifndef GENEAKIT_B200_H
#define GENEAKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GK_OK            0
#define GK_ERR_ARG       1   /* bad argument (null pointer, negative size, parent not before child) */
#define GK_ERR_INDEX     2   /* proband / ancestor rank out of range (reference: std::out_of_range, lib/compute.cpp:649) */
#define GK_ERR_CUDA      3   /* CUDA runtime error or no device */
#define GK_ERR_OOM       4   /* device (or host) memory exhausted */
#define GK_ERR_STATE     5   /* call order violated on a job handle */

#define GK_ABI_VERSION   2

#define GK_FLAG_DIAGONAL_ONLY 1   /* job without the m x m result: only the self kinships / inbreeding coefficients (gen.f) */

typedef struct gk_opts {
    int32_t struct_size;   /* sizeof(gk_opts); lets the ABI grow */
    int32_t device;        /* CUDA device ordinal, -1 = current device */
    int32_t verbose;       /* 1: print the cut sizes / "Cut k out of n" lines of lib/compute.cpp:662-676 */
    int32_t compress;      /* -1 auto (default), 0 one matrix row per individual, 1 one row per sibship */
    int32_t order;         /* -1/1 auto: classes ordered by a locality walk; 0: sorted by parent rows (debug) */
    int32_t rank;          /* row shard owned by this process (multi-GPU), 0 .. world-1; -1 with world > 1: the one-shot
                              entry points drive GPUs 0 .. world-1 from this process (also: environment GENEAKIT_GPUS=N) */
    int32_t world;         /* number of row shards (<= 8), 0/1 = single GPU */
    int32_t flags;         /* GK_FLAG_* bits */
    void   *stream;        /* cudaStream_t to enqueue on, NULL = the library's own stream */
} gk_opts;

/* Defaults: current device, automatic compression and class order, one GPU.  (A zero-filled gk_opts is NOT the default:
 * compress = 0 and order = 0 are the uncompressed / debug-order settings.) */
void gk_opts_init(gk_opts *opts);

/* ---- one-shot entry points: what the binding lambdas call -------------------------------- */

/* Replaces compute_kinships(Pedigree<>&, std::vector<int> proband_ids, bool verbose)
 * (include/compute.hpp:114-116, lib/compute.cpp:638-700) behind the lambda at
 * lib/cgeneakit.cpp:411-431.  pro[0..m) = proband ranks in CALLER order (duplicates allowed);
 * out = m*m doubles, row-major, caller-owned HOST memory (pageable or pinned).
 * mean_or_null, when non-NULL, receives gen.phiMean of the result (geneakit/compute.py:124-132)
 * from the fused device reduction: (sum(all) - sum(diag)) / (m*m - m). */
int gk_phi_f64(int32_t n, const int32_t *sire, const int32_t *dam,
               int32_t m, const int32_t *pro,
               double *out, double *mean_or_null, const gk_opts *opts);

/* Where the last gk_phi_f64 call of this process spent its time: seconds of host planning, allocation + H2D of the plan,
 * kernels, D2H of the result, the whole call (before the buffers go back to the cache); GPUs used. */
void gk_phi_last_timing(double *out6);

/* Replaces compute_genetic_contributions(Pedigree<>&, proband_ids, ancestor_ids)
 * (include/compute.hpp:130-131, lib/compute.cpp:832-862) behind lib/cgeneakit.cpp:491-511.
 * out = m*a doubles, row-major: rows = probands (caller order), columns = ancestors. */
int gk_gc_f64(int32_t n, const int32_t *sire, const int32_t *dam,
              int32_t m, const int32_t *pro, int32_t a, const int32_t *anc,
              double *out, const gk_opts *opts);

/* Replaces compute_inbreedings(Pedigree<>&, proband_ids) (lib/compute.cpp:725-814) behind gen.f (geneakit/compute.py:255-282):
 * out[i] = inbreeding coefficient of proband i = 2 phi(i, i) - 1, caller order.  The recurrence runs without the expansion. */
int gk_f_f64(int32_t n, const int32_t *sire, const int32_t *dam,
             int32_t m, const int32_t *pro, double *out, const gk_opts *opts);

/* Timing / traffic of the last gk_gc_f64 call of this process (the benchmark's --op gc line). */
typedef struct gk_gc_stats {
    double  kernel_ms;           /* CUDA-event time of the propagation + gather kernels, all column blocks */
    double  total_seconds;       /* the whole call: plan, H2D, kernels, D2H */
    int64_t algorithmic_bytes;   /* 24 B x live members (summed over the cuts) x ancestors */
    int64_t member_rows;         /* sum over the cuts of their members */
    int64_t column_blocks;       /* blocks of ancestor columns the call was split into (HBM budget) */
    int64_t block_columns;
    int64_t kernel_launches;
    int32_t n_cuts, reserved;
} gk_gc_stats;
int gk_gc_last_stats(gk_gc_stats *stats);

/* Replaces get_required_memory_for_kinships (include/compute.hpp:110-111,
 * lib/compute.cpp:591-635): GB of the two largest adjacent cut matrices.  Host only.
 * Returns a negative value on error. */
double gk_phi_required_gb(int32_t n, const int32_t *sire, const int32_t *dam,
                          int32_t m, const int32_t *pro);

/* Replace get_proband_ids / get_founder_ids (lib/identify.cpp:55-65, :28-38) in rank space:
 * ranks of childless / parentless individuals, ascending rank; the binding maps to ids and
 * sorts by id.  Return the count; out may be NULL to query it.  Host only. */
int32_t gk_childless(int32_t n, const int32_t *sire, const int32_t *dam, int32_t *out);
int32_t gk_parentless(int32_t n, const int32_t *sire, const int32_t *dam, int32_t *out);

/* Replaces the DFS topological sort of the loader (lib/create.cpp:121-155,
 * convert_pedigree :158-199 with sorted=false).  fidx/midx = FILE position of each parent or
 * -1; order[r] = file position of the individual that gets rank r.  Host only. */
int gk_dfs_order(int32_t n, const int32_t *fidx, const int32_t *midx, int32_t *order);

/* Sizes of the vertex cuts (lib/compute.cpp:119-135) for these probands; returns the number
 * of cuts G (sizes[0..min(G,cap))) or a negative GK_ERR_*.  Host only. */
int32_t gk_cut_sizes(int32_t n, const int32_t *sire, const int32_t *dam,
                     int32_t m, const int32_t *pro, int32_t *sizes, int32_t cap);

/* Multi-GPU: the final m x m matrix is sharded by contiguous blocks of result ROWS (== columns,
 * the matrix is symmetric), one block per rank; opts->rank / opts->world select the block a job
 * produces (the class matrices of the cuts are sharded the same way, by panels of 32 rows).  Host only. */
void gk_shard_rows(int32_t m, int32_t rank, int32_t world, int64_t *row0, int64_t *row1);

const char *gk_last_error(void);
int32_t gk_abi_version(void);

/* Pinned host memory for result matrices (the binding owns it through its capsule deleter,
 * lib/cgeneakit.cpp:420-423); device->host copies into it run at full PCIe rate. */
void *gk_host_alloc(uint64_t bytes);
void  gk_host_free(void *p);
/* The library keeps freed device buffers for the next call (allocating ~110 GB costs more than the kernels of a c4
 * job); this returns them to the driver.  GENEAKIT_CACHE=0 disables the cache. */
void  gk_cache_release(void);

/* ---- device-resident job API: plan once, keep everything in HBM ---------------------------
 * Used by the benchmark, by phiMean-only callers that never need the matrix on the host, and
 * by the multi-GPU driver.  Lifetime: gk_phi_plan -> gk_phi_upload -> gk_phi_run (repeatable)
 * -> gk_phi_mean / gk_phi_copy_out / gk_phi_device_result -> gk_phi_free. */
typedef struct gk_phi_job gk_phi_job;

typedef struct gk_phi_stats {
    int32_t  n_cuts;            /* G */
    int32_t  n_steps;           /* G - 1 recurrence steps */
    int32_t  compressed;        /* 1 if sibship classes are used */
    int32_t  ranked;            /* 1 if the rank-ordered rounding rule is applied per pair (G-1 > 26) */
    int32_t  tile;
    int32_t  m;                 /* probands (caller entries) */
    int64_t  max_cut;           /* largest |cut_g| */
    int64_t  max_classes;       /* largest matrix dimension actually stored */
    int64_t  cells_reference;   /* sum_{g>=1} |cut_g|^2 : cells the reference's loop nest updates */
    int64_t  cells_device;      /* sum_{g>=1} K_g^2 (+ m*m for the final expansion) */
    int64_t  algorithmic_bytes_reference;  /* 8*sum(P_g*|cut_{g-1}| + |cut_g|^2 - k_g^2) */
    int64_t  algorithmic_bytes_device;     /* 8*(sum_g reads(K_{g-1} rows needed) + K_g^2) + 8*m*m */
    int64_t  device_bytes;      /* HBM allocated for matrices + plan */
    int64_t  kernel_launches;   /* launches one gk_phi_run enqueues */
    double   plan_seconds;      /* host planning time */
    int64_t  plan_bytes;        /* int32 plan arrays copied host->device by gk_phi_upload */
    int64_t  panel_row_bytes;   /* bytes of previous-matrix rows phase 1 reads: 8 * sum_g sum_panels (distinct parent rows) * Kpad_{g-1} */
    int64_t  corrections;       /* "same individual" (i, j) corrections applied by the diagonal kernel, all steps */
    int64_t  rank_row_bytes;    /* parent-row bytes THIS rank's phase 1 reads (its own panels) */
    int64_t  rank_remote_row_bytes; /* ... of which from rows owned by another rank: NVLink bytes per pass */
} gk_phi_stats;

int gk_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam,
                int32_t m, const int32_t *pro, const gk_opts *opts, gk_phi_job **job);
int gk_phi_upload(gk_phi_job *job);                 /* allocate HBM, copy the plan (H2D) */
/* Plan again from the job's inputs and copy the plan to the device again, on the buffers already allocated: with
 * gk_phi_run this is one "flattened arrays on the host -> matrix resident in HBM" pass (what the benchmark times).
 * gk_phi_reupload only repeats the copy (ranks of a one-process multi-GPU job that share rank 0's plan). */
int gk_phi_replan(gk_phi_job *job);
int gk_phi_reupload(gk_phi_job *job);
/* One PASS of a one-GPU job from its inputs with the three stages overlapped: the host plans the cuts in order on its
 * threads, each cut's arrays are copied (pinned staging, own copy stream) and its step launched as soon as they are
 * ready, while the later cuts are still being planned.  Asynchronous like gk_phi_run.  The one-shot entry points work
 * the same way.  Sharded jobs driven cut by cut: gk_phi_pass_begin, then gk_phi_pass_step(g) before every
 * gk_phi_run_step(g), and gk_phi_pass_end after gk_phi_finish. */
int gk_phi_pass(gk_phi_job *job);
int gk_phi_pass_begin(gk_phi_job *job);
int gk_phi_pass_step(gk_phi_job *job, int32_t step);
int gk_phi_pass_end(gk_phi_job *job);
/* One process per GPU with ONE planner: rank 0 calls gk_phi_pass_begin_all instead of gk_phi_pass_begin and plans every
 * pass for all the ranks; gk_phi_pass_step(g) then also copies every other rank's view of cut g straight into that
 * rank's device segments (mapped once with gk_phi_seg_ipc_export on the owner / gk_phi_seg_ipc_open on rank 0, one
 * 64-byte handle per segment, gk_phi_seg_count segments).  The other ranks only enqueue kernels; the per-cut collective
 * of the driver orders a copy before the kernel that reads it (geneakit_b200/distributed.py). */
int gk_phi_pass_begin_all(gk_phi_job *job);
int gk_phi_seg_count(gk_phi_job *job, int32_t *n_segments);
int gk_phi_seg_ipc_export(gk_phi_job *job, int32_t segment, void *handle /* 64 bytes */, int64_t *capacity_ints);
int gk_phi_seg_ipc_open(gk_phi_job *job, int32_t peer_rank, int32_t segment, const void *handle, int64_t capacity_ints);
/* Another rank of `job` in the same process (shares its plan); device = CUDA ordinal for that rank. */
int gk_phi_clone_rank(gk_phi_job *job, int32_t rank, int32_t device, gk_phi_job **out);
int gk_phi_run(gk_phi_job *job);                    /* enqueue init + all steps + expansion; async */
int gk_phi_sync(gk_phi_job *job);                   /* wait for the job's stream */
int gk_phi_mean(gk_phi_job *job, double *sum_all, double *sum_diag, double *mean);
int gk_phi_copy_out(gk_phi_job *job, double *host_out);            /* (row1-row0)*m doubles: this shard's rows */
/* rows [row_begin, row_begin + nrows) of the result (caller order; inside this job's block) -> nrows*m doubles */
int gk_phi_copy_rows(gk_phi_job *job, int64_t row_begin, int64_t nrows, double *host_out);
int gk_phi_device_result(gk_phi_job *job, double **dev, int64_t *ld); /* this shard's rows of the m x m matrix, in HBM */
int gk_phi_get_stats(gk_phi_job *job, gk_phi_stats *stats);
int gk_phi_cut_sizes(gk_phi_job *job, int32_t *sizes, int32_t cap);   /* |cut_g| as the reference counts them */
int gk_phi_class_sizes(gk_phi_job *job, int32_t *sizes, int32_t cap); /* K_g actually stored */
/* Average device time of the dominant kernel over the last gk_phi_run, from CUDA events
 * recorded on the job's stream around every step launch (ms per step, n_steps entries). */
int gk_phi_step_times(gk_phi_job *job, float *ms, int32_t cap);
/* ALGORITHMIC bytes of each launch timed by gk_phi_step_times (same indexing): for step g
 * 8*(P_g*K_{g-1} + K_g^2 - k_g^2) -- SURVEY 8(d)'s formula on the matrices actually stored
 * (16 B per updated cell on discrete-generation pedigrees) -- and 8*(m^2 + min(m,K)^2) for the
 * final expansion. */
int gk_phi_step_bytes(gk_phi_job *job, int64_t *bytes, int32_t cap);
/* ---- device-side consumers of a finished job: the m x m matrix need not cross PCIe ------------- */
int gk_phi_diag(gk_phi_job *job, double *self_kinship /* m, caller order */);
int gk_phi_f(gk_phi_job *job, double *inbreeding /* m: 2 phi(i,i) - 1 */);
/* gen.phiOver (geneakit/compute.py:135-190): pairs i > j with kinship > threshold, reference order; indices are positions
 * in the caller's proband list.  gk_phi_over leaves the pairs in HBM, gk_phi_over_fetch copies `count` triples out. */
int gk_phi_over(gk_phi_job *job, double threshold, int64_t *count);
int gk_phi_over_fetch(gk_phi_job *job, int32_t *pro1_idx, int32_t *pro2_idx, double *kinship);
/* the CSR gen.phi(sparse=True) returns (lib/compute.cpp:272-487 behind lib/cgeneakit.cpp:433-486): lower triangle with
 * the diagonal, float32, |2 phi| > 1e-9; indptr has m + 1 entries. */
int gk_phi_sparse(gk_phi_job *job, int64_t *nnz);
int gk_phi_sparse_fetch(gk_phi_job *job, int64_t *indptr, int32_t *indices, float *data);
/* gen.phiCI's bootstrap statistic (geneakit/compute.py:226-229) for b resamples of n indices each: mean of the entries
 * < 0.5 of phi[idx][:, idx]. */
int gk_phi_ci_stats(gk_phi_job *job, const int32_t *idx /* b*n */, int32_t b, int32_t n, double *stats /* b */);

/* GK_PROF=1 only: cycle counters of the gather4 step kernel's warp roles, summed over warps and steps of the last run:
 * [0] producer loop, [1] its idle polls, [2] polls blocked on a stage, [3] on the ring slot; [4] phase-1 consumers,
 * [5] of which waiting for TMA data; [7] phase-2 warps, [8] of which waiting for phase 1, [10] publishing (fence). */
int gk_phi_debug_prof(gk_phi_job *job, uint64_t *out16);
void gk_phi_free(gk_phi_job *job);

/* ---- multi-GPU: one job per rank, the class matrices sharded by rows (= columns: they are symmetric) --
 * Every rank plans the same pedigree with opts->rank / opts->world set, uploads, exchanges the CUDA-IPC
 * handles of its two matrix buffers (gk_phi_ipc_export -> all-gather -> gk_phi_ipc_open for every other
 * rank -> gk_phi_peers_ready), then drives the cuts in lock step:
 *     for g in 0 .. G-1:  gk_phi_run_step(job, g);  all-gather of the self-kinship vector (gk_phi_dvec)
 *     gk_phi_finish(job)
 * The step kernel reads the parent rows of its own panels straight from the owners' HBM over NVLink
 * (peer loads inside the kernel); the all-gather of d (8 bytes per class) is the only collective and
 * doubles as the barrier between cuts; it must be enqueued on the job's stream (or ordered against it), and
 * gk_phi_finish reads the peers' rows once more: a stream-ordered collective after it must precede the next
 * gk_phi_run_step(job, 0) and gk_phi_free.  gk_phi_set_peer wires jobs that live in ONE process (several ranks emulated
 * on one GPU -- all on ONE stream: persistent kernels of different ranks must never be co-scheduled on a GPU). */
int gk_phi_ipc_export(gk_phi_job *job, void *handles /* 2 x 64 bytes */);
int gk_phi_ipc_open(gk_phi_job *job, int32_t peer_rank, const void *handles);
int gk_phi_set_stream(gk_phi_job *job, void *stream);     /* before gk_phi_upload: enqueue on this cudaStream_t */
int gk_phi_stream(gk_phi_job *job, void **stream);        /* the cudaStream_t the job enqueues on (after gk_phi_upload) */
int gk_phi_local_buffers(gk_phi_job *job, void **buf0, void **buf1);
int gk_phi_set_peer(gk_phi_job *job, int32_t peer_rank, void *buf0, void *buf1);
int gk_phi_peers_ready(gk_phi_job *job);
int gk_phi_run_step(gk_phi_job *job, int32_t step);      /* step 0 = founders' matrix, 1 .. G-1 = recurrence */
/* RANKED (deeper than 26 generations) sharded jobs evaluate the tiles J <= I only and every rank stores into its own rows;
 * after the collective that follows gk_phi_run_step(job, g) every rank calls gk_phi_run_pull(job, g) -- its tiles (J own,
 * I of a higher rank) are read out of the owners' rows over NVLink and transposed -- and a second stream-ordered collective
 * must follow before the next step.  A no-op for every other job. */
int gk_phi_run_pull(gk_phi_job *job, int32_t step);
int gk_phi_finish(gk_phi_job *job);                      /* expansion to caller order + phiMean partial sums */
/* Self kinship per class of cut `step` (device pointer, replicated on every rank): rank r writes
 * doubles [r*chunk, (r+1)*chunk) and the driver all-gathers the `total` = world*chunk doubles. */
int gk_phi_dvec(gk_phi_job *job, int32_t step, void **dev, int64_t *chunk_doubles, int64_t *total_doubles);
/* one-process emulation of that all-gather (tests): copy rank src's chunk of cut `step` into job's replica */
int gk_phi_dvec_copy_from(gk_phi_job *job, gk_phi_job *src, int32_t step);

/* Plan introspection for tests (host arrays, valid until gk_phi_free): the per-step plan the kernels
 * consume.  See DESIGN.md "Plan layout". */
typedef struct gk_step_view {
    int64_t K, Kpad, K_prev, Kpad_prev;   /* classes of this cut / the previous one (padded to 32) */
    int64_t panel_rows;                   /* sum over panels of the distinct parent rows */
    int32_t np, panel, ranked, ngroups;   /* panels, classes per panel (32), ranked mode, correction groups */
    int32_t panel_begin, panel_end;       /* panels owned by this job's rank */
    const int32_t *pF, *pM;               /* Kpad: parent rows in the previous class matrix, -1 = none */
    const int32_t *flags;                 /* Kpad: 1 singleton, 2 kept from the previous cut, 4 both parents known */
    const int32_t *cls_ind, *cls_size;    /* K: lowest-ranked member, number of members */
    /* corrections: group k = (pg_i[k] >= pg_j[k]); terms pg_off[k] .. pg_off[k+1]:
     * B[i][j] += 1/4 * pt_mult * (d_prev[pt_cls] - A_prev[pt_cls][pt_cls]) */
    const int32_t *pg_i, *pg_j, *pg_off, *pt_cls, *pt_mult;
} gk_step_view;
int gk_phi_step_view(gk_phi_job *job, int32_t step, gk_step_view *view);  /* step = 0 .. G-1 */
/* Final expansion map: for caller entry i the class row in the last matrix and the individual. */
int gk_phi_final_view(gk_phi_job *job, const int32_t **cls, const int32_t **ind, int64_t *K_last);

#ifdef __cplusplus
}
#endif
#endif /* GENEAKIT_B200_H */
