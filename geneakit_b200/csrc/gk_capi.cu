// C ABI of libgeneakit_b200.so (see include/geneakit_b200.h).  Host orchestration only:
// plan on the host (gk_plan.hpp), keep matrices in HBM, enqueue the sm_100a kernels
// (gk_kernels.cu) on one stream.  There is no CPU compute path in this file.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/geneakit_b200.h"
#include "gk_gcplan.hpp"
#include "gk_kernels.cuh"
#include "gk_plan.hpp"

namespace {

thread_local std::string g_err;
std::mutex g_call_mu;              // the one-shot entry points run one call at a time (SURVEY 8b "internal mutex")
gk_gc_stats g_gc_stats{};
double g_phi_timing[6] = {};       // last gk_phi_f64: plan, upload, kernels, D2H, total seconds; GPUs used
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorMemoryAllocation) ? GK_ERR_OOM : GK_ERR_CUDA;
}
#define GK_CUDA(call)                                            \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);      \
    } while (0)

// ---- device memory cache (SURVEY 8b "lazily creates CUDA contexts/streams once per process and caches them") -------
// cudaMalloc / cudaFree of the ~110 GB a c4 job needs cost more than the kernels.  Freed blocks are kept per device and
// handed out again to the next job that asks for a similar size (best fit within 25 %); on an allocation failure the
// cache of that device is dropped and the request retried.  gk_cache_release() / GENEAKIT_CACHE=0 switch it off.
struct PoolBlock { void *p; size_t bytes; int dev; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool_free, g_pool_live;
bool pool_enabled() {
    static const bool on = [] { const char *e = std::getenv("GENEAKIT_CACHE"); return !(e && e[0] == '0'); }();
    return on;
}
void pool_drop(int dev) {       // caller holds the lock
    for (size_t i = 0; i < g_pool_free.size();)
        if (dev < 0 || g_pool_free[i].dev == dev) {
            cudaSetDevice(g_pool_free[i].dev);
            cudaFree(g_pool_free[i].p);
            g_pool_free[i] = g_pool_free.back(); g_pool_free.pop_back();
        } else i++;
}
cudaError_t pool_alloc(void **out, size_t bytes, int dev) {
    bytes = std::max<size_t>(bytes, 256);
    std::lock_guard<std::mutex> lock(g_pool_mu);
    int best = -1;
    for (size_t i = 0; i < g_pool_free.size(); i++) {
        const PoolBlock &b = g_pool_free[i];
        if (b.dev != dev || b.bytes < bytes || b.bytes > bytes + bytes / 4 + (1 << 20)) continue;
        if (best < 0 || b.bytes < g_pool_free[best].bytes) best = (int) i;
    }
    if (best >= 0) {
        *out = g_pool_free[best].p;
        g_pool_live.push_back(g_pool_free[best]);
        g_pool_free[best] = g_pool_free.back(); g_pool_free.pop_back();
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        pool_drop(dev);
        cudaSetDevice(dev);
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) g_pool_live.push_back({ *out, bytes, dev });
    return e;
}
void pool_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (size_t i = 0; i < g_pool_live.size(); i++)
        if (g_pool_live[i].p == p) {
            PoolBlock b = g_pool_live[i];
            g_pool_live[i] = g_pool_live.back(); g_pool_live.pop_back();
            if (pool_enabled()) g_pool_free.push_back(b);
            else { cudaSetDevice(b.dev); cudaFree(p); }
            return;
        }
    cudaFree(p);
}
template <class T> cudaError_t pool_alloc_t(T **out, size_t bytes, int dev) { void *p = nullptr; cudaError_t e = pool_alloc(&p, bytes, dev); *out = static_cast<T *>(p); return e; }

// Pinned host staging of the plan (one block per cut and job): the copies to the device are truly asynchronous, so the
// planner threads, the copy engine and the step kernels of earlier cuts all run at once.  Cached like the device blocks.
struct HostBlock { void *p; size_t bytes; };
std::vector<HostBlock> g_hpool_free;
void *host_pool_alloc(size_t bytes, size_t *cap) {
    bytes = std::max<size_t>(bytes, 4096);
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_hpool_free.size(); i++) {
            const HostBlock &b = g_hpool_free[i];
            if (b.bytes < bytes || b.bytes > bytes + bytes / 2 + (1 << 16)) continue;
            if (best < 0 || b.bytes < g_hpool_free[best].bytes) best = (int) i;
        }
        if (best >= 0) {
            HostBlock b = g_hpool_free[best];
            g_hpool_free[best] = g_hpool_free.back(); g_hpool_free.pop_back();
            *cap = b.bytes;
            return b.p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *cap = bytes;
    return p;
}
void host_pool_free(void *p, size_t cap) {
    if (!p) return;
    if (!pool_enabled()) { cudaFreeHost(p); return; }
    std::lock_guard<std::mutex> lock(g_pool_mu);
    g_hpool_free.push_back({ p, cap });
}

// One plan segment (the arrays of one cut, this rank's view) in pinned host memory.
struct PlanBlob {
    int32_t *p = nullptr;
    size_t cap = 0;      // bytes
    size_t n = 0;        // int32 entries in use
    bool packed = false;
};

// Collects the arrays of a segment (every array 16-byte aligned), then writes them into the blob in one go.
struct Packer {
    struct Piece { const void *src; size_t bytes, at; };
    std::vector<Piece> pieces;
    std::list<std::vector<int32_t>> keep_;      // temporaries that must live until write()
    size_t total = 0;                           // int32 entries
    size_t add_raw(const void *src, size_t bytes) {
        const size_t o = total;
        pieces.push_back({ src, bytes, o });
        total += (bytes / 4 + 3) / 4 * 4;
        return o;
    }
    size_t add(const std::vector<int32_t> &v) { return add_raw(v.data(), v.size() * 4); }
    size_t add(const std::vector<int64_t> &v) { return add_raw(v.data(), v.size() * 8); }
    std::vector<int32_t> &keep(std::vector<int32_t> &&v) { keep_.push_back(std::move(v)); return keep_.back(); }
    bool write(PlanBlob &b) {
        const size_t bytes = std::max<size_t>(total, 4) * 4;
        if (b.cap < bytes) {
            host_pool_free(b.p, b.cap);
            b.p = static_cast<int32_t *>(host_pool_alloc(bytes, &b.cap));
            if (!b.p) { b.cap = 0; return false; }
        }
        for (const Piece &q : pieces) {
            if (q.bytes) std::memcpy(b.p + q.at, q.src, q.bytes);
            const size_t end = q.at + (q.bytes / 4 + 3) / 4 * 4;
            for (size_t k = q.at + q.bytes / 4; k < end; k++) b.p[k] = 0;
        }
        b.n = std::max<size_t>(total, 4);
        b.packed = true;
        return true;
    }
};

struct StepOffsets {
    size_t prow = 0, pcnt = 0, slots = 0, prow8 = 0, pcnt8 = 0, slots8 = 0, first_use = 0;
    size_t pF = 0, pM = 0, flags = 0, pg_i = 0, pg_j = 0, pg_off = 0, pt_cls = 0, pt_mult = 0, grp_off = 0, grp_info = 0;
    size_t done = 0;              // offset into the counter arena: done1 at +0, done2 at +np
    int32_t ngroups = 0, npatch = 0;
    int64_t total_items = 0;
    int32_t pb = 0, pe = 0, cpp = 0;      // own panels [pb, pe), panels per rank
    size_t grp_off2 = 0; int64_t total1 = 0, total2 = 0; int32_t n1pp = 0, ring_tma = 4;   // multi-GPU kernel: two item streams
    int32_t lag = 1, ring_n = 3;          // phase 2 trails phase 1 by `lag` panels; panel buffers in the ring
    // one-GPU TMA gather4 kernel (phi_step_g4_kernel): padded row lists per panel, item prefix, its own ring depth
    size_t prow4 = 0, pcnt4 = 0, slots4 = 0, item_off = 0;
    int64_t total_g4 = 0;
    int32_t use_g4 = 0, ring_g4 = 4;
};

}  // namespace

struct gk_plan_inputs {             // what the plan was built from (kept for gk_phi_replan / gk_phi_pass)
    int32_t n = 0, m = 0;
    std::vector<int32_t> sire, dam, pro;
};

struct gk_phi_job {
    std::shared_ptr<gk::PhiPlan> planp = std::make_shared<gk::PhiPlan>();   // shared by the ranks of a one-process multi-GPU job
    std::shared_ptr<gk_plan_inputs> inputs;
    gk_opts opts{};
    int device = 0;
    int rank = 0, world = 1;
    bool device_ready = false;        // device selected and configured, streams and events created
    bool buffers = false;             // matrices, vectors, ring allocated
    bool uploaded = false, ran = false;
    bool no_expand = false;           // opts.flags & GK_FLAG_DIAGONAL_ONLY: no m x m result (gen.f only needs the self kinships)
    int steps_done = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;               // H2D of the plan segments (overlaps the step kernels of earlier cuts)
    cudaEvent_t pass_ev = nullptr;                    // everything enqueued before the current pass (the copies must not overtake it)
    std::vector<cudaEvent_t> seg_ev;                  // segment g is on the device
    std::vector<cudaEvent_t> fetch_ev;                // sharded jobs: chunk k of the expansion's class rows has been fetched
    double *buf[2] = { nullptr, nullptr };            // own rows of the class matrices, ping-pong by step parity
    int64_t buf_doubles[2] = { 0, 0 };
    double *peer[gk::kMaxShards][2] = {};             // every rank's buffers (own entry = buf); IPC mappings for the others
    bool peer_ipc[gk::kMaxShards][2] = {};
    bool peers_set = false;
    double *dvec[2] = { nullptr, nullptr };           // self kinship per class (full length, replicated)
    double *dvec_peer[gk::kMaxShards][2] = {};        // one-process multi-GPU: the other ranks' replicas (written by the diagonal kernel)
    int n_dvec_peer = 0;
    int64_t dvec_len = 0;
    double *ring = nullptr;
    int64_t ring_doubles = 0;
    double *zero_row = nullptr;
    int use_tma = 0;                  // step kernel of small cuts / ranked sharded jobs: 0 = warp-autonomous LDGSTS, 1 = TMA phase 1 (gk::KernelChoice)
    int g4_min_k = 0;                 // smallest previous-cut matrix the default kernel is used for (tests force 32)
    int use_g4 = 0;                   // default kernel: TMA bulk-copy phase 1 + L2-direct phase 2
    int full_rows = 0;                // phase 2 evaluates every tile J and stores only the own rows (multi-GPU; GK_FULL_ROWS=1 forces it)
    int lag_env = 0, ring_env = 0, lag_scale_pct = 100;   // tuning overrides (GK_LAG, GK_RING, GK_LAG_PCT)
    double *out = nullptr;            // result rows; single GPU: aliases the buffer the last step does not use
    double *out_own = nullptr;        // separately allocated result (world > 1)
    int64_t out_own_doubles = 0;
    double *gloc = nullptr;           // world > 1: class rows needed by the expansion, fetched from their owners
    int64_t gloc_doubles = 0;
    double *partials = nullptr;
    int64_t partials_blocks = 0;
    double *sums_dev = nullptr;
    double *sums_host = nullptr;      // pinned
    // the plan on the device: one segment per cut (g = 0 .. G-1) + the final expansion maps (segment G), this rank's view
    std::vector<int32_t *> seg;
    std::vector<size_t> seg_cap;      // int32 entries allocated
    std::vector<PlanBlob> blob;       // pinned host image of every segment
    int32_t *counters = nullptr;
    size_t counters_count = 0, counters_cap = 0;
    std::vector<StepOffsets> off;     // offsets inside the segment of each cut
    size_t off_final_cls = 0, off_final_ind = 0, off_cact = 0, off_coff = 0, off_cmem = 0;
    int64_t row0 = 0, row1 = 0;       // result rows owned by this shard (caller order)
    int32_t nact = 0, nchunks = 0;
    int32_t fetch_first = 0;          // world > 1: position in the list of active classes of the first row the NEXT rank owns
    int64_t nblocks_expand = 0;
    double *scratch = nullptr;
    unsigned long long *prof = nullptr;   // GK_PROF=1: role cycle counters of the gather4 step kernel, summed over the steps
    std::vector<cudaEvent_t> ev;      // start / end events of every step (2g, 2g+1); 0, 1: the final expansion
    int64_t device_bytes = 0;
    int64_t launches = 0;
    std::vector<int64_t> cut_row_bytes, cut_remote_bytes;   // bytes of parent rows this rank's phase 1 reads / reads from peers, per cut
    std::string pack_error;
    // a pass in flight (gk_phi_pass_begin .. gk_phi_pass_end): the cuts are planned while the device runs the earlier ones
    std::unique_ptr<gk::PhiPlanner> planner;
    std::shared_ptr<gk::PhiPlan> pass_plan;
    double pass_t0 = 0.0, pass_prefix_s = 0.0;
    // one process per GPU, rank 0 plans for everybody (gk_phi_pass_begin_all): host-only jobs of the other ranks (their
    // views of the plan are packed here) and CUDA-IPC mappings of those ranks' plan segments (the copies go straight there)
    bool shadow = false;
    std::vector<gk_phi_job *> shadows;
    std::vector<std::vector<int32_t *>> peer_seg;     // [rank][segment]
    std::vector<std::vector<size_t>> peer_seg_cap;    // int32 entries
};

namespace {

void destroy(gk_phi_job *j) {
    if (!j) return;
    j->planner.reset();               // joins the planner threads (they pack into this job's blobs)
    for (gk_phi_job *sh : j->shadows) destroy(sh);
    j->shadows.clear();
    if (!j->peer_seg.empty() && j->device_ready) {
        cudaSetDevice(j->device);
        for (auto &v : j->peer_seg) for (int32_t *p : v) if (p) cudaIpcCloseMemHandle(p);
    }
    if (j->device_ready) cudaSetDevice(j->device);
    for (int q = 0; q < gk::kMaxShards; q++)
        for (int p = 0; p < 2; p++)
            if (j->peer_ipc[q][p] && j->peer[q][p]) cudaIpcCloseMemHandle(j->peer[q][p]);
    if (j->stream && cudaStreamSynchronize(j->stream) != cudaSuccess) cudaGetLastError();   // nothing of this job may still run when its blocks go back to the cache
    if (j->copy_stream && cudaStreamSynchronize(j->copy_stream) != cudaSuccess) cudaGetLastError();
    for (int p = 0; p < 2; p++) { pool_free(j->buf[p]); pool_free(j->dvec[p]); }
    pool_free(j->ring);
    pool_free(j->zero_row);
    pool_free(j->out_own);
    pool_free(j->gloc);
    pool_free(j->partials);
    pool_free(j->scratch);
    pool_free(j->prof);
    pool_free(j->sums_dev);
    if (j->sums_host) cudaFreeHost(j->sums_host);
    for (int32_t *p : j->seg) pool_free(p);
    for (PlanBlob &b : j->blob) host_pool_free(b.p, b.cap);
    pool_free(j->counters);
    for (cudaEvent_t e : j->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : j->seg_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : j->fetch_ev) cudaEventDestroy(e);
    if (j->pass_ev) cudaEventDestroy(j->pass_ev);
    if (j->copy_stream) cudaStreamDestroy(j->copy_stream);
    if (j->own_stream && j->stream) cudaStreamDestroy(j->stream);
    delete j;
}

int select_device(const gk_opts *o, int *dev_out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_err = std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                "); geneakit_b200 has no CPU fallback";
        return GK_ERR_CUDA;
    }
    int dev = 0;
    if (o && o->device >= 0) dev = o->device;
    else GK_CUDA(cudaGetDevice(&dev));
    if (dev >= count) return fail(GK_ERR_ARG, "device ordinal out of range");
    GK_CUDA(cudaSetDevice(dev));
    *dev_out = dev;
    return GK_OK;
}

// function attributes (dynamic shared memory) are per device: configure each device the first time a job uploads to it
std::mutex g_cfg_mu;
int g_cfg_state[64] = {};            // 0 = not configured, 1 = ok, 2 = failed
cudaError_t g_cfg_err = cudaSuccess;
cudaError_t configure_device(int dev) {
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (g_cfg_state[dev] == 0) {
        // GK_L2_PERSIST_MB: set aside L2 for the lines the step kernel marks evict_last (the ring of transposed panels).
        // Measured at c4: 64 MB cut the DRAM traffic per step from 35.0 to 30.2 GB (27.8 with a 7-deep ring: 1.11 x the
        // algorithmic bytes) but the step went from 6.9 to 9.5 ms -- the streams need the L2 more than the ring does.  Off.
        int mb = 0;
        if (const char *e = std::getenv("GK_L2_PERSIST_MB")) mb = std::atoi(e);
        if (mb > 0) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0)
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t) mb << 20, (size_t) prop.persistingL2CacheMaxSize));
            cudaGetLastError();
        }
        g_cfg_err = gk::configure_kernels();
        g_cfg_state[dev] = g_cfg_err == cudaSuccess ? 1 : 2;
    }
    return g_cfg_state[dev] == 1 ? cudaSuccess : g_cfg_err;
}

// panels of step g owned by each rank: equal blocks of cpp = ceil(np / world) panels
void panel_range(int np, int rank, int world, int32_t *pb, int32_t *pe, int32_t *cpp) {
    const int c = (np + world - 1) / world;
    *cpp = c;
    *pb = std::min(np, rank * c);
    *pe = std::min(np, (rank + 1) * c);
}

gk::RowTable row_table(const gk_phi_job *j, int g) {
    gk::RowTable t{};
    const gk::StepPlan &S = j->planp->steps[g];
    const int cpp = j->off[g].cpp;
    t.n = j->world;
    for (int q = 0; q < gk::kMaxShards; q++) {
        t.base[q] = q < j->world ? j->peer[q][g & 1] : nullptr;
        t.row0[q] = (int32_t) std::min<int64_t>(S.Kpad, (int64_t) q * cpp * gk::kPanel);
    }
    t.row0[gk::kMaxShards] = (int32_t) S.Kpad;
    for (int q = j->world; q <= gk::kMaxShards; q++) t.row0[q] = (int32_t) S.Kpad;
    return t;
}

// Host threads the planner of this process may use: all of them, or its share when torchrun runs one process per GPU.
int planner_threads() {
    int hw = (int) std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) { const int w = std::atoi(e); if (w > 1) hw = std::max(2, hw / w); }
    return hw;
}

}  // namespace

// Everything of a job that follows from the SIZES of the cuts (known once the planner has found the classes: PhiPlanner::
// begin) and from (rank, world): kernel choice, panel ranges, result rows, ring depths, completion counters.
static void configure_job(gk_phi_job *j) {
    const gk::PhiPlan &P = *j->planp;
    j->no_expand = (j->opts.flags & GK_FLAG_DIAGONAL_ONLY) != 0;
    // a rank of a multi-GPU job evaluates every tile J and stores only its own rows; RANKED jobs (rounding order
    // fixed by rank) must keep J <= I and store the mirror tile into the rows of J's owner instead
    j->full_rows = (j->world > 1 && !P.ranked) ? 1 : 0;
    const gk::KernelChoice kc = gk::KernelChoice::resolve(j->world, P.ranked);
    j->use_tma = kc.use_tma; j->use_g4 = kc.use_g4; j->g4_min_k = kc.g4_min_k;
    if (const char *e = std::getenv("GK_FULL_ROWS")) if (e[0] == '1' && !P.ranked) j->full_rows = 1;
    if (const char *e = std::getenv("GK_LAG")) j->lag_env = std::atoi(e);
    if (const char *e = std::getenv("GK_RING")) j->ring_env = std::atoi(e);
    if (const char *e = std::getenv("GK_LAG_PCT")) j->lag_scale_pct = std::max(1, std::atoi(e));
    const int G = P.G;
    j->off.assign(G, StepOffsets{});
    j->cut_row_bytes.assign(G, 0); j->cut_remote_bytes.assign(G, 0);
    gk_shard_rows(P.m, j->rank, j->world, &j->row0, &j->row1);
    size_t ncount = 0;
    for (int g = 0; g < G; g++) {
        const gk::StepPlan &S = P.steps[g];
        StepOffsets &o = j->off[g];
        panel_range(S.np, j->rank, j->world, &o.pb, &o.pe, &o.cpp);
        if (g == 0) continue;
        // ring of transposed panels / lag of phase 2 behind phase 1 (warp-autonomous kernel): the resident warps work on a
        // window of ~nwarps consecutive queue items; phase 2 of a panel should start only after that many items have been
        // dealt since its phase 1 ended
        const int ncb = (int) ((S.Kpad_prev + gk::kCB - 1) / gk::kCB);
        const int64_t nwarps = (int64_t) gk::step_grid() * gk::kStepWarps;
        int lag = (int) ((nwarps * j->lag_scale_pct / 100 + ncb - 1) / ncb);
        lag = std::max(1, std::min(lag, (int) gk::kRingMax - 2));
        if (j->lag_env > 0) lag = std::min(j->lag_env, (int) gk::kRingMax - 1);
        o.lag = lag;
        o.ring_n = j->ring_env > 0 ? std::max(lag + 1, std::min(j->ring_env, (int) gk::kRingMax)) : lag + 2;
        const int ncb256 = (int) ((S.Kpad_prev + gk::kP1Cols - 1) / gk::kP1Cols);
        o.n1pp = (gk::kPanel / gk::kP1Classes) * ncb256;
        o.total1 = (int64_t) (o.pe - o.pb) * o.n1pp;
        o.ring_tma = j->ring_env > 0 ? std::max(2, std::min(j->ring_env, (int) gk::kRingMax)) : 4;
        if (j->use_tma) o.ring_n = o.ring_tma;
        o.use_g4 = kc.cut_uses_g4(S.Kpad_prev) ? 1 : 0;       // small cuts: the warp-autonomous kernel has less fixed cost
        if (o.use_g4) {
            const int64_t panel_bytes = S.Kpad_prev * gk::kPanel * 8;
            int rn = (int) std::max<int64_t>(2, std::min<int64_t>(12, ((int64_t) 104 << 20) / panel_bytes));   // half of a slot is written on average (rows some tile J <= I reads)
            if (j->ring_env > 0) rn = std::max(2, std::min(j->ring_env, (int) gk::kRingMax));
            o.ring_g4 = rn;
            o.ring_n = rn;
        }
        o.done = ncount; ncount += 2 * (size_t) S.np + 4;      // + the 8-byte aligned phase-2 queue counter
        ncount = (ncount + 1) / 2 * 2;
    }
    j->counters_count = std::max<size_t>(ncount, 1);
    j->nchunks = (int32_t) ((P.m + gk::kExpandCols - 1) / gk::kExpandCols);
    if ((int) j->blob.size() != G + 1) {
        for (PlanBlob &b : j->blob) host_pool_free(b.p, b.cap);
        j->blob.assign((size_t) G + 1, PlanBlob());
    }
    for (PlanBlob &b : j->blob) b.packed = false;
}

// The arrays of cut g as this rank's kernels read them, packed into the pinned blob of segment g (and, with the last cut,
// the final expansion maps into segment G).  Runs on a planner thread as soon as the cut is planned.
static bool pack_cut(gk_phi_job *j, int g) {
    gk::PhiPlan &P = *j->planp;
    const int G = P.G;
    const gk::StepPlan &S = P.steps[g];
    StepOffsets &o = j->off[g];
    Packer pk;
    if (g == 0) o.flags = pk.add(S.flags);
    else {
        o.pF = pk.add(S.pF); o.pM = pk.add(S.pM); o.flags = pk.add(S.flags);
        o.first_use = pk.add(S.first_use);
        o.pg_i = pk.add(S.pg_i); o.pg_j = pk.add(S.pg_j); o.pg_off = pk.add(S.pg_off);
        o.pt_cls = pk.add(S.pt_cls); o.pt_mult = pk.add(S.pt_mult);
        o.npatch = (int32_t) S.pg_i.size();
        const bool tma = !o.use_g4 && j->use_tma, ldgsts = !o.use_g4 && !j->use_tma;
        if ((o.use_g4 && !S.has_units) || (!o.use_g4 && !S.has_panel) || (tma && !S.has_sub)) {
            j->pack_error = "internal: the plan was built for another step kernel (GK_STEP_KERNEL / GK_G4_MIN_K changed between plan and upload)";
            return false;
        }
        if (!o.use_g4) { o.prow = pk.add(S.prow); o.pcnt = pk.add(S.pcnt); o.slots = pk.add(S.slots); }
        if (tma) { o.prow8 = pk.add(S.prow8); o.pcnt8 = pk.add(S.pcnt8); o.slots8 = pk.add(S.slots8); }
        if (ldgsts) {
            // work queue of the own panels: phase 2 trails phase 1 by `lag` panels: P1(s), P2(s - lag), ...
            const int ncb = (int) ((S.Kpad_prev + gk::kCB - 1) / gk::kCB);
            std::vector<int64_t> goff(1, 0);
            std::vector<int32_t> ginfo;
            auto add = [&](int phase, int panel) {
                const int64_t items = phase == 1 ? ncb : (j->full_rows ? S.np : panel + 1);
                ginfo.push_back(panel | ((phase - 1) << 30));
                goff.push_back(goff.back() + items);
            };
            for (int s = o.pb; s < o.pe + o.lag; s++) {
                if (s < o.pe) add(1, s);
                if (s - o.lag >= o.pb) add(2, s - o.lag);
            }
            o.ngroups = (int32_t) ginfo.size();
            o.total_items = goff.back();
            std::vector<int32_t> g32(goff.size() * 2);
            std::memcpy(g32.data(), goff.data(), goff.size() * 8);
            o.grp_off = pk.add(pk.keep(std::move(g32))); o.grp_info = pk.add(pk.keep(std::move(ginfo)));
        }
        {   // phase-2 items (one tile each) of the own panels: first item of each panel
            std::vector<int32_t> g32b(((size_t) (o.pe - o.pb) + 1) * 2);
            int64_t *goff2 = reinterpret_cast<int64_t *>(g32b.data());
            goff2[0] = 0;
            // (a rank of a sharded job evaluates every tile J for its panels -- except, with the default kernel, the tiles of
            // its OWN panels above the diagonal: inside its own block it works like one GPU, J <= I and both pieces stored)
            for (int s = o.pb; s < o.pe; s++)
                goff2[s - o.pb + 1] = goff2[s - o.pb] + (j->full_rows ? (o.use_g4 ? S.np - (o.pe - 1 - s) : S.np) : s + 1);
            o.total2 = goff2[o.pe - o.pb];
            o.grp_off2 = pk.add(pk.keep(std::move(g32b)));
        }
        if (o.use_g4) {
            // default kernel: phase-1 items = one unit of 16 classes x W columns (W = 512 or 256 per unit)
            const int nu = 2 * S.np;
            std::vector<int32_t> i32(((size_t) nu + 1) * 2);
            int64_t *ioff = reinterpret_cast<int64_t *>(i32.data());
            ioff[0] = 0;
            for (int u = 0; u < nu; u++) {
                const int W = (S.ucnt[u] >> 16) ? 512 : 256;
                const int p = u >> 1;
                ioff[u + 1] = ioff[u] + (p >= o.pb && p < o.pe ? (S.Kpad_prev + W - 1) / W : 0);
            }
            o.total_g4 = ioff[nu];
            o.prow4 = pk.add(S.urow); o.pcnt4 = pk.add(S.ucnt); o.slots4 = pk.add(S.uslots); o.item_off = pk.add(pk.keep(std::move(i32)));
        }
        // rows of the previous matrix this rank's panels read, and how many of them live on a peer (NVLink)
        {
            const int64_t rps = (int64_t) j->off[g - 1].cpp * gk::kPanel;
            std::vector<int32_t> stamp((size_t) std::max<int64_t>(S.K_prev, 1), -1);
            const int unit = o.use_g4 ? 16 : (j->use_tma ? gk::kSub : gk::kPanel);   // classes whose parent rows are staged together
            int64_t rb = 0, rr = 0;
            for (int p = o.pb * gk::kPanel / unit; p < o.pe * gk::kPanel / unit; p++)
                for (int i = p * unit; i < (p + 1) * unit; i++)
                    for (int q = 0; q < 2; q++) {
                        const int32_t r = q ? S.pM[i] : S.pF[i];
                        if (r < 0 || stamp[r] == p) continue;
                        stamp[r] = p;
                        rb += 8 * S.Kpad_prev;
                        if (r / rps != j->rank) rr += 8 * S.Kpad_prev;
                    }
            j->cut_row_bytes[g] = rb; j->cut_remote_bytes[g] = rr;
        }
    }
    if (!pk.write(j->blob[g])) { j->pack_error = "cudaHostAlloc failed (pinned staging of the plan)"; return false; }
    if (g != G - 1) return true;
    // ---- segment G: the final expansion maps; the caller entries of this shard's result rows grouped by class ----
    Packer pf;
    j->off_final_cls = pf.add(P.final_cls);
    j->off_final_ind = pf.add(P.final_ind);
    const int64_t Kl = P.steps[G - 1].K;
    std::vector<int32_t> cnt((size_t) Kl + 1, 0), cact, coff(1, 0), cmem;
    if (!P.single_cut) {
        for (int64_t i = j->row0; i < j->row1; i++) cnt[P.final_cls[i] + 1]++;
        for (int64_t c = 0; c < Kl; c++) if (cnt[c + 1] > 0) { cact.push_back((int32_t) c); coff.push_back(coff.back() + cnt[c + 1]); }
        std::vector<int32_t> slot((size_t) Kl, -1), w(coff.begin(), coff.end() - 1);
        for (size_t a = 0; a < cact.size(); a++) slot[cact[a]] = (int32_t) a;
        cmem.assign((size_t) (j->row1 - j->row0), 0);
        for (int64_t i = j->row0; i < j->row1; i++) cmem[w[slot[P.final_cls[i]]]++] = (int32_t) i;
    }
    j->nact = (int32_t) cact.size();
    {
        const int64_t next_row0 = (int64_t) ((j->rank + 1) % j->world) * j->off[G - 1].cpp * gk::kPanel;
        j->fetch_first = (int32_t) (std::lower_bound(cact.begin(), cact.end(), (int32_t) std::min<int64_t>(next_row0, INT32_MAX)) - cact.begin());
        if (j->fetch_first >= j->nact) j->fetch_first = 0;
    }
    j->off_cact = pf.add(cact); j->off_coff = pf.add(coff); j->off_cmem = pf.add(cmem);
    j->nblocks_expand = P.single_cut ? 1 : (int64_t) j->nact * j->nchunks;
    if (!pf.write(j->blob[G])) { j->pack_error = "cudaHostAlloc failed (pinned staging of the plan)"; return false; }
    return true;
}

extern "C" {

const char *gk_last_error(void) { return g_err.c_str(); }

void gk_opts_init(gk_opts *o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->struct_size = (int32_t) sizeof *o;
    o->device = -1; o->compress = -1; o->order = -1; o->rank = 0; o->world = 1;
}
int32_t gk_abi_version(void) { return GK_ABI_VERSION; }

void *gk_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        g_err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
void gk_host_free(void *p) { if (p) cudaFreeHost(p); }

void gk_cache_release(void) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    pool_drop(-1);
    for (const HostBlock &b : g_hpool_free) cudaFreeHost(b.p);
    g_hpool_free.clear();
}

int32_t gk_childless(int32_t n, const int32_t *sire, const int32_t *dam, int32_t *out) {
    if (n < 0 || (n > 0 && (!sire || !dam))) { g_err = "bad argument"; return -GK_ERR_ARG; }
    std::vector<uint8_t> has((size_t) n + 1, 0);
    for (int32_t k = 0; k < n; k++) {
        if (sire[k] >= 0 && sire[k] < n) has[sire[k]] = 1;
        if (dam[k] >= 0 && dam[k] < n) has[dam[k]] = 1;
    }
    int32_t c = 0;
    for (int32_t k = 0; k < n; k++) if (!has[k]) { if (out) out[c] = k; c++; }
    return c;
}

int32_t gk_parentless(int32_t n, const int32_t *sire, const int32_t *dam, int32_t *out) {
    if (n < 0 || (n > 0 && (!sire || !dam))) { g_err = "bad argument"; return -GK_ERR_ARG; }
    int32_t c = 0;
    for (int32_t k = 0; k < n; k++) if (sire[k] < 0 && dam[k] < 0) { if (out) out[c] = k; c++; }
    return c;
}

int gk_dfs_order(int32_t n, const int32_t *fidx, const int32_t *midx, int32_t *order) {
    if (n < 0 || (n > 0 && (!fidx || !midx || !order))) return fail(GK_ERR_ARG, "bad argument");
    // Post-order DFS (father subtree, mother subtree, self), roots in file order:
    // lib/create.cpp:121-155.  Explicit stack of (node, next phase).
    std::vector<uint8_t> done((size_t) n, 0), open((size_t) n, 0);
    std::vector<std::pair<int32_t, uint8_t>> st;
    int32_t r = 0;
    for (int32_t root = 0; root < n; root++) {
        if (done[root]) continue;
        st.emplace_back(root, 0);
        open[root] = 1;
        while (!st.empty()) {
            auto &top = st.back();
            const int32_t x = top.first;
            if (top.second < 2) {
                const int32_t p = top.second == 0 ? fidx[x] : midx[x];
                top.second++;
                if (p >= 0) {
                    if (p >= n) return fail(GK_ERR_ARG, "parent position out of range");
                    if (!done[p]) {
                        if (open[p]) return fail(GK_ERR_ARG, "pedigree contains a cycle");
                        open[p] = 1;
                        st.emplace_back(p, 0);
                    }
                }
            } else {
                order[r++] = x;
                done[x] = 1;
                st.pop_back();
            }
        }
    }
    return GK_OK;
}

void gk_shard_rows(int32_t m, int32_t rank, int32_t world, int64_t *row0, int64_t *row1) {
    if (world < 1) world = 1;
    if (rank < 0) rank = 0;
    if (rank >= world) rank = world - 1;
    // contiguous blocks of result rows, sizes differing by at most one
    if (row0) *row0 = (int64_t) m * rank / world;
    if (row1) *row1 = (int64_t) m * (rank + 1) / world;
}

int32_t gk_cut_sizes(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
                     int32_t *sizes, int32_t cap) {
    std::vector<int64_t> s;
    std::string err;
    int G = gk::cut_sizes_only(n, sire, dam, m, pro, s, err);
    if (G < 0) { g_err = err; return G == -2 ? -GK_ERR_INDEX : -GK_ERR_ARG; }
    for (int i = 0; i < G && i < cap; i++) sizes[i] = (int32_t) s[i];
    return G;
}

double gk_phi_required_gb(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro) {
    std::vector<int64_t> s;
    std::string err;
    int G = gk::cut_sizes_only(n, sire, dam, m, pro, s, err);
    if (G < 0) { g_err = err; return -1.0; }
    if (G < 2) return 0.0;                                   // lib/compute.cpp:609-611
    int64_t best = -1, s1 = 0, s2 = 0;
    for (int i = 0; i + 1 < G; i++)
        if (s[i] + s[i + 1] > best) { best = s[i] + s[i + 1]; s1 = s[i]; s2 = s[i + 1]; }   // first maximum, :620-631
    return (1.0 * s1 * s1 + 1.0 * s2 * s2) * sizeof(double) / 1e9;
}

// ---------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

gk_phi_job *new_job(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro, const gk_opts *opts) {
    gk_phi_job *j = new gk_phi_job();
    if (opts) std::memcpy(&j->opts, opts, std::min<size_t>(sizeof(gk_opts), (size_t) std::max(0, opts->struct_size)));
    else { j->opts.device = -1; j->opts.compress = -1; j->opts.order = -1; }
    if (j->opts.world < 1) j->opts.world = 1;
    if (j->opts.world > gk::kMaxShards) { delete j; g_err = "world > 8 shards"; return nullptr; }
    if (j->opts.rank < 0 || j->opts.rank >= j->opts.world) { delete j; g_err = "rank outside [0, world)"; return nullptr; }
    j->rank = j->opts.rank; j->world = j->opts.world;
    if (n > 0 && m > 0 && sire && dam && pro) {
        j->inputs = std::make_shared<gk_plan_inputs>();
        j->inputs->n = n; j->inputs->m = m;
        j->inputs->sire.assign(sire, sire + n); j->inputs->dam.assign(dam, dam + n); j->inputs->pro.assign(pro, pro + m);
    }
    return j;
}

int plan_error(const gk::PhiPlan &P) { return fail(P.error_code == 2 ? GK_ERR_INDEX : GK_ERR_ARG, P.error); }

// device selected and configured, streams and events of the job created (once)
int prepare_device(gk_phi_job *j) {
    if (j->device_ready) { GK_CUDA(cudaSetDevice(j->device)); return GK_OK; }
    int rc = select_device(&j->opts, &j->device);
    if (rc) return rc;
    { const cudaError_t ce = configure_device(j->device); if (ce != cudaSuccess) return cuda_fail(ce, "kernel configuration (is this an sm_100 device?)"); }
    if (j->opts.stream) j->stream = (cudaStream_t) j->opts.stream;
    else { GK_CUDA(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking)); j->own_stream = true; }
    GK_CUDA(cudaStreamCreateWithFlags(&j->copy_stream, cudaStreamNonBlocking));
    GK_CUDA(cudaEventCreateWithFlags(&j->pass_ev, cudaEventDisableTiming));
    j->device_ready = true;
    return GK_OK;
}

// Matrices, vectors, ring, counters: everything whose size follows from the sizes of the cuts (configure_job).  Blocks that
// are large enough already are kept (a job plans the same inputs again and again in the benchmark).
int alloc_buffers(gk_phi_job *j) {
    gk::PhiPlan &P = *j->planp;
    const int G = P.G;
    GK_CUDA(cudaSetDevice(j->device));
    auto ensure = [&](double **p, int64_t *have, int64_t need) -> cudaError_t {
        need = std::max<int64_t>(need, 32);
        if (*p && *have >= need) return cudaSuccess;
        pool_free(*p); *p = nullptr; *have = 0;
        const cudaError_t e = pool_alloc_t(p, (size_t) need * sizeof(double), j->device);
        if (e == cudaSuccess) { *have = need; j->device_bytes += need * 8; }
        return e;
    };
    if (j->counters_cap < j->counters_count) {
        pool_free(j->counters); j->counters = nullptr;
        GK_CUDA(pool_alloc_t(&j->counters, j->counters_count * sizeof(int32_t), j->device));
        j->counters_cap = j->counters_count;
    }
    // ---- matrices: two ping-pong buffers holding the own rows of each cut's class matrix --------
    int64_t need[2] = { 0, 0 };
    int64_t dlen = 1;
    for (int g = 0; g < G; g++) {
        const gk::StepPlan &S = P.steps[g];
        const int64_t rows = (int64_t) (j->off[g].pe - j->off[g].pb) * gk::kPanel;
        need[g & 1] = std::max(need[g & 1], rows * S.Kpad);
        dlen = std::max(dlen, (int64_t) j->off[g].cpp * gk::kPanel * j->world);
    }
    const int64_t mm = j->no_expand ? 0 : (j->row1 - j->row0) * (int64_t) P.m;
    const int out_buf = P.single_cut ? 0 : 1 - ((G - 1) & 1);
    if (j->world == 1) need[out_buf] = std::max(need[out_buf], mm);
    for (int p = 0; p < 2; p++) {
        const cudaError_t e = ensure(&j->buf[p], &j->buf_doubles[p], need[p]);
        if (e != cudaSuccess) {
            cudaGetLastError();
            char msg[256];
            std::snprintf(msg, sizeof msg, "out of device memory: kinship buffers need %.2f GB + %.2f GB of HBM",
                          need[0] * 8 / 1e9, need[1] * 8 / 1e9);
            return fail(GK_ERR_OOM, msg);
        }
        j->peer[j->rank][p] = j->buf[p];
    }
    if (j->world == 1) { j->out = j->buf[out_buf]; j->peers_set = true; }
    else {
        GK_CUDA(ensure(&j->out_own, &j->out_own_doubles, mm));
        j->out = j->out_own;
    }
    if (j->dvec_len < dlen) {
        for (int p = 0; p < 2; p++) { pool_free(j->dvec[p]); j->dvec[p] = nullptr; GK_CUDA(pool_alloc_t(&j->dvec[p], (size_t) dlen * sizeof(double), j->device)); j->device_bytes += dlen * 8; }
        j->dvec_len = dlen;
    }
    int64_t ring_doubles = 32;
    for (int g = 1; g < G; g++) ring_doubles = std::max(ring_doubles, (int64_t) j->off[g].ring_n * P.steps[g].Kpad_prev * gk::kPanel);
    GK_CUDA(ensure(&j->ring, &j->ring_doubles, ring_doubles));
    if (!j->zero_row) {
        GK_CUDA(pool_alloc_t(&j->zero_row, 1024 * sizeof(double), j->device));
        GK_CUDA(cudaMemsetAsync(j->zero_row, 0, 1024 * sizeof(double), j->stream));
        GK_CUDA(pool_alloc_t(&j->scratch, 2 * 256 * sizeof(double), j->device));
        GK_CUDA(pool_alloc_t(&j->sums_dev, 2 * sizeof(double), j->device));
        if (const char *e = std::getenv("GK_PROF")) if (e[0] == '1') GK_CUDA(pool_alloc_t(&j->prof, 16 * sizeof(unsigned long long), j->device));
        GK_CUDA(cudaHostAlloc(&j->sums_host, 2 * sizeof(double), cudaHostAllocDefault));
    }
    // per-CTA partial sums of the expansion: at most one block row per class that has a caller entry among the own rows
    int64_t pblocks = P.single_cut ? 1 : std::min<int64_t>(P.steps[G - 1].K, std::max<int64_t>(j->row1 - j->row0, 1)) * std::max(j->nchunks, 1);
    if (!P.single_cut && j->world == 1) pblocks = std::max(pblocks, gk::expand_sym_blocks((int32_t) P.steps[G - 1].K, P.m));
    if (j->partials_blocks < pblocks) {
        pool_free(j->partials); j->partials = nullptr;
        GK_CUDA(pool_alloc_t(&j->partials, (size_t) pblocks * 2 * sizeof(double), j->device));
        j->partials_blocks = pblocks; j->device_bytes += pblocks * 16;
    }
    if ((int) j->ev.size() != 2 * (G + 1)) {
        for (cudaEvent_t e : j->ev) cudaEventDestroy(e);
        j->ev.assign(2 * (size_t) (G + 1), nullptr);
        for (auto &e : j->ev) GK_CUDA(cudaEventCreate(&e));
    }
    if ((int) j->seg.size() != G + 1) {
        for (int32_t *p : j->seg) pool_free(p);
        for (cudaEvent_t e : j->seg_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : j->fetch_ev) cudaEventDestroy(e);
        j->seg.assign((size_t) G + 1, nullptr); j->seg_cap.assign((size_t) G + 1, 0);
        j->seg_ev.assign((size_t) G + 1, nullptr);
        for (auto &e : j->seg_ev) GK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    j->buffers = true;
    return GK_OK;
}

// The copies of a pass must not overtake the kernels of the previous pass that still read the segments.
int begin_copies(gk_phi_job *j) {
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaEventRecord(j->pass_ev, j->stream));
    GK_CUDA(cudaStreamWaitEvent(j->copy_stream, j->pass_ev, 0));
    return GK_OK;
}

// Segment g (g = G: the final maps) of this job's plan: H2D on the copy stream; the job's stream waits for it.
int feed_segment(gk_phi_job *j, int g) {
    PlanBlob &b = j->blob[g];
    if (!b.packed) return fail(GK_ERR_STATE, j->pack_error.empty() ? "internal: plan segment not packed" : j->pack_error);
    GK_CUDA(cudaSetDevice(j->device));
    if (j->seg_cap[g] < b.n) {
        pool_free(j->seg[g]); j->seg[g] = nullptr;
        GK_CUDA(pool_alloc_t(&j->seg[g], b.n * sizeof(int32_t), j->device));
        j->seg_cap[g] = b.n; j->device_bytes += (int64_t) b.n * 4;
    }
    GK_CUDA(cudaMemcpyAsync(j->seg[g], b.p, b.n * sizeof(int32_t), cudaMemcpyHostToDevice, j->copy_stream));
    GK_CUDA(cudaEventRecord(j->seg_ev[g], j->copy_stream));
    GK_CUDA(cudaStreamWaitEvent(j->stream, j->seg_ev[g], 0));
    return GK_OK;
}

int64_t plan_bytes(const gk_phi_job *j) {
    int64_t b = 0;
    for (const PlanBlob &x : j->blob) if (x.packed) b += (int64_t) x.n * 4;
    return b;
}

// Blocking: pack every segment of a COMPLETE plan (in parallel) and copy them all.
int upload_plan(gk_phi_job *j) {
    const int G = j->planp->G;
    std::atomic<bool> bad(false);
    gk::parallel_for(0, G, std::min(planner_threads(), 24), [&](int g) { if (!pack_cut(j, g)) bad = true; });
    if (bad) return fail(GK_ERR_STATE, j->pack_error);
    int rc = begin_copies(j);
    for (int g = 0; g <= G && !rc; g++) rc = feed_segment(j, g);
    return rc;
}

}  // namespace

extern "C" {

int gk_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
                const gk_opts *opts, gk_phi_job **job) {
    if (!job) return fail(GK_ERR_ARG, "job pointer is null");
    *job = nullptr;
    gk_phi_job *j = new_job(n, sire, dam, m, pro, opts);
    if (!j) return GK_ERR_ARG;
    const int threads = planner_threads();
    bool planned;
    { gk::PhiPlanner pl(n, sire, dam, m, pro, j->opts.compress, j->opts.order, j->world, *j->planp, threads); planned = pl.begin(); if (planned) { pl.start(); planned = pl.finish(); } }
    if (!planned && j->planp->error_code == 3) {
        // the sibship compression does not pay for this pedigree: the same job, one row per individual (same results)
        j->opts.compress = 0;
        gk::PhiPlanner pl(n, sire, dam, m, pro, 0, j->opts.order, j->world, *j->planp, threads);
        planned = pl.begin();
        if (planned) { pl.start(); planned = pl.finish(); }
    }
    if (!planned) {
        const int code = plan_error(*j->planp);
        delete j;
        return code;
    }
    configure_job(j);
    *job = j;
    return GK_OK;
}

int gk_phi_upload(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (j->uploaded) return GK_OK;
    int rc = prepare_device(j);
    if (rc) return rc;
    configure_job(j);                 // again: the ring depths depend on the device (resident CTAs of the step kernel)
    rc = alloc_buffers(j);
    if (rc) return rc;
    rc = upload_plan(j);
    if (rc) return rc;
    GK_CUDA(cudaStreamSynchronize(j->copy_stream));
    j->uploaded = true;
    return GK_OK;
}

// ---- multi-GPU plumbing: every rank's ping-pong buffers become addressable by every rank -------
int gk_phi_ipc_export(gk_phi_job *j, void *handles /* 2 * 64 bytes */) {
    if (!j || !j->uploaded || !handles) return fail(GK_ERR_STATE, "job not uploaded");
    GK_CUDA(cudaSetDevice(j->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int p = 0; p < 2; p++)
        GK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handles) + p, j->buf[p]));
    return GK_OK;
}
int gk_phi_ipc_open(gk_phi_job *j, int32_t q, const void *handles) {
    if (!j || !j->uploaded || !handles) return fail(GK_ERR_STATE, "job not uploaded");
    if (q < 0 || q >= j->world || q == j->rank) return fail(GK_ERR_ARG, "bad peer rank");
    GK_CUDA(cudaSetDevice(j->device));
    for (int p = 0; p < 2; p++) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, reinterpret_cast<const cudaIpcMemHandle_t *>(handles) + p, sizeof h);
        void *ptr = nullptr;
        GK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        j->peer[q][p] = static_cast<double *>(ptr);
        j->peer_ipc[q][p] = true;
    }
    return GK_OK;
}
int gk_phi_set_stream(gk_phi_job *j, void *stream) {
    if (!j || j->uploaded) return fail(GK_ERR_STATE, "gk_phi_set_stream after gk_phi_upload");
    j->opts.stream = stream;
    return GK_OK;
}
int gk_phi_stream(gk_phi_job *j, void **stream) {
    if (!j || !j->uploaded || !stream) return fail(GK_ERR_STATE, "job not uploaded");
    *stream = (void *) j->stream;
    return GK_OK;
}
int gk_phi_local_buffers(gk_phi_job *j, void **buf0, void **buf1) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    if (buf0) *buf0 = j->buf[0];
    if (buf1) *buf1 = j->buf[1];
    return GK_OK;
}
int gk_phi_set_peer(gk_phi_job *j, int32_t q, void *buf0, void *buf1) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    if (q < 0 || q >= j->world || q == j->rank) return fail(GK_ERR_ARG, "bad peer rank");
    j->peer[q][0] = static_cast<double *>(buf0); j->peer[q][1] = static_cast<double *>(buf1);
    return GK_OK;
}
int gk_phi_peers_ready(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    for (int q = 0; q < j->world; q++)
        for (int p = 0; p < 2; p++)
            if (!j->peer[q][p]) return fail(GK_ERR_STATE, "a peer buffer has not been set");
    j->peers_set = true;
    return GK_OK;
}
int gk_phi_dvec(gk_phi_job *j, int32_t step, void **dev, int64_t *chunk_doubles, int64_t *total_doubles) {
    if (!j || !j->uploaded || step < 0 || step >= j->planp->G) return fail(GK_ERR_ARG, "bad step");
    if (dev) *dev = j->dvec[step & 1];
    if (chunk_doubles) *chunk_doubles = (int64_t) j->off[step].cpp * gk::kPanel;
    if (total_doubles) *total_doubles = (int64_t) j->off[step].cpp * gk::kPanel * j->world;
    return GK_OK;
}

// Same-process emulation of the all-gather (several ranks' jobs on ONE GPU, one stream): copy the
// chunk rank `src` owns of the self-kinship vector of cut `step` into this job's replica.
int gk_phi_dvec_copy_from(gk_phi_job *j, gk_phi_job *src, int32_t step) {
    if (!j || !src || !j->uploaded || !src->uploaded || step < 0 || step >= j->planp->G) return fail(GK_ERR_ARG, "bad argument");
    if (j == src) return GK_OK;
    const int64_t chunk = (int64_t) j->off[step].cpp * gk::kPanel;
    GK_CUDA(cudaMemcpyAsync(j->dvec[step & 1] + src->rank * chunk, src->dvec[step & 1] + src->rank * chunk,
                            (size_t) chunk * sizeof(double), cudaMemcpyDeviceToDevice, j->stream));
    return GK_OK;
}

// One cut: step 0 initialises the founders' matrix, step g >= 1 runs the recurrence kernel and the
// diagonal kernel.  With world > 1 the driver all-gathers the self-kinship vector (gk_phi_dvec) after
// every step -- that collective is also the barrier that makes the rows written here readable by
// the peers' next step.
int gk_phi_run_step(gk_phi_job *j, int32_t g) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (!j->uploaded) return fail(GK_ERR_STATE, "gk_phi_run_step before gk_phi_upload");
    if (!j->peers_set) return fail(GK_ERR_STATE, "peer buffers are not set (gk_phi_ipc_open / gk_phi_set_peer, gk_phi_peers_ready)");
    gk::PhiPlan &P = *j->planp;
    if (g < 0 || g >= P.G) return fail(GK_ERR_ARG, "bad step");
    GK_CUDA(cudaSetDevice(j->device));
    cudaStream_t st = j->stream;
    const gk::StepPlan &S = P.steps[g];
    const StepOffsets &o = j->off[g];
    if (g == 0) {
        GK_CUDA(cudaMemsetAsync(j->counters, 0, j->counters_count * sizeof(int32_t), st));
        if (j->prof) GK_CUDA(cudaMemsetAsync(j->prof, 0, 16 * sizeof(unsigned long long), st));
        j->launches = 0;
        if (P.single_cut) { j->steps_done = 1; return GK_OK; }
        // own rows of the founders' matrix
        const int64_t rows = (int64_t) (o.pe - o.pb) * gk::kPanel;
        if (j->world == 1) {
            GK_CUDA(gk::launch_init(j->buf[0], j->dvec[0], j->seg[0] + o.flags, S.K, S.Kpad, st));
        } else {
            // (rare) uncompressed multi-GPU start: every rank zeroes its rows, the diagonal comes from the flags
            GK_CUDA(cudaMemsetAsync(j->buf[0], 0, (size_t) std::max<int64_t>(rows * S.Kpad, 1) * sizeof(double), st));
            GK_CUDA(gk::launch_init_rows(j->buf[0], j->dvec[0], j->seg[0] + o.flags, S.K, S.Kpad,
                                         (int64_t) o.pb * gk::kPanel, rows, j->dvec_len, st));
        }
        j->launches++;
        j->steps_done = 1;
        return GK_OK;
    }
    if (j->opts.verbose && j->rank == 0) std::printf("Cut %d out of %d\n", g, P.G);
    const int32_t *sg = j->seg[g];
    gk::StepDev d{};
    d.prev = row_table(j, g - 1);
    d.outt = row_table(j, g);
    d.out_base = j->buf[g & 1]; d.out_row0 = (int64_t) o.pb * gk::kPanel;
    d.ld_prev = S.Kpad_prev; d.ld_new = S.Kpad;
    d.Kpad_prev = (int32_t) S.Kpad_prev; d.np = S.np;
    d.pF = sg + o.pF; d.pM = sg + o.pM;
    d.prow = sg + o.prow; d.pcnt = sg + o.pcnt; d.slots = sg + o.slots;
    d.first_use = sg + o.first_use;
    d.prow8 = sg + o.prow8; d.pcnt8 = sg + o.pcnt8; d.slots8 = sg + o.slots8;
    d.zero_row = j->zero_row;
    d.ring = j->ring; d.ring_stride = S.Kpad_prev * gk::kPanel; d.ring_n = o.ring_n;
    d.done1 = j->counters + o.done; d.done2 = d.done1 + S.np;
    d.grp_off = reinterpret_cast<const int64_t *>(sg + o.grp_off); d.grp_info = sg + o.grp_info;
    d.total_items = o.total_items; d.ngroups = o.ngroups;
    d.ncb = (int32_t) ((S.Kpad_prev + gk::kCB - 1) / gk::kCB);
    d.panel_begin = o.pb; d.panel_end = o.pe;
    d.full_rows = j->full_rows;
    { const char *e = std::getenv("GK_OUT_POLICY"); d.out_policy = e ? std::atoi(e) : 0; }
    { const char *e = std::getenv("GK_DEBUG_NOWAIT"); d.debug_nowait = e ? std::atoi(e) : 0; }
    GK_CUDA(cudaEventRecord(j->ev[2 * g], st));
    d.grp_off2 = reinterpret_cast<const int64_t *>(sg + o.grp_off2);
    d.total1 = o.total1; d.total2 = o.total2; d.n1pp = o.n1pp;
    d.ncb256 = (int32_t) ((S.Kpad_prev + gk::kP1Cols - 1) / gk::kP1Cols);
    if (o.use_g4) {
        d.prow4 = sg + o.prow4; d.pcnt4 = sg + o.pcnt4; d.slots4 = sg + o.slots4;
        d.item_off = reinterpret_cast<const int64_t *>(sg + o.item_off);
        d.total_g4 = o.total_g4;
        d.prof = j->prof;
        { const char *e = std::getenv("GK_CHUNK2"); d.chunk2 = e ? std::atoi(e) : 0; }
        d.queue2 = reinterpret_cast<unsigned long long *>(j->counters + ((o.done + 2 * (size_t) S.np + 1) / 2 * 2));
        GK_CUDA(gk::launch_step_g4(d, st));
    } else if (j->use_tma) GK_CUDA(gk::launch_step_tma(d, st));
    else GK_CUDA(gk::launch_step(d, st));
    gk::DiagDev q{};
    q.prev = d.prev; q.out = row_table(j, g); q.ld_prev = d.ld_prev; q.ld_new = d.ld_new;
    q.dprev = j->dvec[(g - 1) & 1]; q.dnew = j->dvec[g & 1];
    q.n_peer = j->n_dvec_peer;
    for (int r = 0; r < j->n_dvec_peer; r++) q.dnew_peer[r] = j->dvec_peer[r][g & 1];
    q.pF = d.pF; q.pM = d.pM; q.flags = sg + o.flags;
    q.c_begin = (int32_t) std::min<int64_t>(S.K, (int64_t) o.pb * gk::kPanel);
    q.c_end = (int32_t) std::min<int64_t>(S.K, (int64_t) o.pe * gk::kPanel);
    q.pg_i = sg + o.pg_i; q.pg_j = sg + o.pg_j; q.pg_off = sg + o.pg_off;
    q.pt_cls = sg + o.pt_cls; q.pt_mult = sg + o.pt_mult; q.ngroups = o.npatch;
    GK_CUDA(gk::launch_diag(q, st));
    GK_CUDA(cudaEventRecord(j->ev[2 * g + 1], st));
    j->launches += 2;
    j->steps_done = g + 1;
    return GK_OK;
}

// Ranked (deep) sharded jobs only, after the barrier that follows gk_phi_run_step(job, g) and before a second barrier: pull
// the mirror tiles of this rank's rows out of the rows of the higher ranks (see pull_mirror_kernel).  A no-op for every
// other job.
int gk_phi_run_pull(gk_phi_job *j, int32_t g) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    gk::PhiPlan &P = *j->planp;
    if (g < 0 || g >= P.G) return fail(GK_ERR_ARG, "bad step");
    if (j->world < 2 || !P.ranked || g == 0 || P.single_cut) return GK_OK;
    GK_CUDA(cudaSetDevice(j->device));
    const StepOffsets &o = j->off[g];
    const gk::StepPlan &S = P.steps[g];
    GK_CUDA(gk::launch_pull_mirror(row_table(j, g), S.Kpad, o.pb, o.pe, S.np, j->buf[g & 1], (int64_t) o.pb * gk::kPanel, j->stream));
    j->launches++;
    return GK_OK;
}

// Final expansion to the caller's proband order + the fused phiMean partial sums of the own rows.
int gk_phi_finish(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    gk::PhiPlan &P = *j->planp;
    const int G = P.G;
    if (j->steps_done < G) return fail(GK_ERR_STATE, "gk_phi_finish before the last gk_phi_run_step");
    GK_CUDA(cudaSetDevice(j->device));
    cudaStream_t st = j->stream;
    const int last = G - 1;
    GK_CUDA(cudaEventRecord(j->ev[0], st));
    if (j->no_expand) {
        GK_CUDA(cudaEventRecord(j->ev[1], st));
        j->ran = true;
        return GK_OK;
    }
    int64_t nblk = j->nblocks_expand;
    if (P.single_cut) {
        GK_CUDA(gk::launch_eye(j->out, P.m, j->row0, j->row1, j->partials, st));
    } else {
        const int32_t *sf = j->seg[G];
        gk::ExpandDev e{};
        e.d = j->dvec[last & 1]; e.ld = P.steps[last].Kpad;
        e.cls = sf + j->off_final_cls; e.ind = sf + j->off_final_ind;
        e.cact = sf + j->off_cact; e.coff = sf + j->off_coff; e.cmem = sf + j->off_cmem;
        e.out = j->out; e.m = P.m; e.row0 = j->row0; e.row1 = j->row1;
        e.nact = j->nact; e.nchunks = j->nchunks; e.partials = j->partials;
        if (j->world == 1) { e.G = j->buf[last & 1]; e.compact = 0; }
        else {
            // the class rows of the own probands live on their owners: fetch them (peer reads over NVLink)
            const int64_t ldt = gk::round_up(std::max(j->nact, 1), 32);
            const int64_t gl = ldt * P.steps[G - 1].Kpad;
            if (j->gloc_doubles < gl) {
                pool_free(j->gloc); j->gloc = nullptr; j->gloc_doubles = 0;
                GK_CUDA(pool_alloc_t(&j->gloc, (size_t) gl * sizeof(double), j->device));
                j->gloc_doubles = gl; j->device_bytes += gl * 8;
            }
            const char *envx = std::getenv("GK_EXPAND");
            // even m: the rows are fetched TRANSPOSED and expanded by the kernel that reads 64-byte segments (the gather
            // kernel is bound by its 8-byte gathers: 26 ms for HALF of c4's rows at two GPUs)
            const bool symw = (P.m & 1) == 0 && !(envx && envx[0] == 'o') && gk::expand_sym_blocks(j->nact, P.m) <= j->partials_blocks;
            e.G = j->gloc; e.compact = 1;
            if (symw) { e.ld = ldt; nblk = gk::expand_sym_blocks(j->nact, P.m); }
            // Fetch and expand in chunks of the active classes (blocks of 32): the fetch of chunk k + 1 (copy stream, NVLink
            // reads) runs while chunk k is expanded (HBM writes).  The chunks start with the rows of the NEXT rank: every rank
            // walks the owners in its own order instead of all of them queueing on rank 0's links first, then on rank 1's, ...
            const int64_t nb32 = (j->nact + 31) / 32;
            const int nchunk = (int) std::max<int64_t>(1, std::min<int64_t>(8, nb32 / 16));
            if ((int) j->fetch_ev.size() < nchunk + 1) {
                const size_t have = j->fetch_ev.size();
                j->fetch_ev.resize((size_t) nchunk + 1, nullptr);
                for (size_t k = have; k < j->fetch_ev.size(); k++) GK_CUDA(cudaEventCreateWithFlags(&j->fetch_ev[k], cudaEventDisableTiming));
            }
            GK_CUDA(cudaEventRecord(j->fetch_ev[nchunk], st));                       // the last cut is complete (on every rank: the driver's collective precedes)
            GK_CUDA(cudaStreamWaitEvent(j->copy_stream, j->fetch_ev[nchunk], 0));
            const gk::RowTable rt = row_table(j, last);
            const int64_t Kpad_last = P.steps[last].Kpad;
            for (int k = 0; k < nchunk && j->nact > 0; k++) {
                const int64_t p0 = nb32 * k / nchunk, p1 = nb32 * (k + 1) / nchunk;
                // blocks [p0, p1) of the ROTATED list = at most two runs of the list
                const int64_t s0 = (j->fetch_first / 32 + p0) % nb32, len = p1 - p0, run1 = std::min<int64_t>(len, nb32 - s0);
                const int64_t runs[2][2] = { { s0, run1 }, { 0, len - run1 } };
                for (int r = 0; r < 2; r++) {
                    const int32_t a_begin = (int32_t) (runs[r][0] * 32);
                    const int32_t a_count = (int32_t) std::min<int64_t>(runs[r][1] * 32, j->nact - a_begin);
                    if (runs[r][1] <= 0 || a_count <= 0) continue;
                    if (symw) GK_CUDA(gk::launch_fetch_rows_t(rt, Kpad_last, e.cact, a_begin, a_count, j->nact, (int32_t) (Kpad_last / 32), j->gloc, ldt, j->copy_stream));
                    else GK_CUDA(gk::launch_fetch_rows(rt, Kpad_last, e.cact, a_begin, a_count, j->gloc, j->copy_stream));
                }
                GK_CUDA(cudaEventRecord(j->fetch_ev[k], j->copy_stream));
                GK_CUDA(cudaStreamWaitEvent(st, j->fetch_ev[k], 0));
                for (int r = 0; r < 2; r++) {
                    const int32_t a_begin = (int32_t) (runs[r][0] * 32);
                    const int32_t a_count = (int32_t) std::min<int64_t>(runs[r][1] * 32, j->nact - a_begin);
                    if (runs[r][1] <= 0 || a_count <= 0) continue;
                    if (symw) GK_CUDA(gk::launch_expand_sym(e, a_begin, a_count, st));
                    else GK_CUDA(gk::launch_expand(e, a_begin, a_count, st));
                }
                j->launches += 2;
            }
        }
        // one GPU and every class of the last cut has a caller entry (always, unless the driver shards the result rows):
        // the expansion that reads 64-byte segments of the symmetric matrix (GK_EXPAND=old: the gather kernel)
        const char *env = std::getenv("GK_EXPAND");
        const bool sym = j->world == 1 && j->nact == (int32_t) P.steps[last].K && !(env && env[0] == 'o') && (P.m & 1) == 0 &&
                         gk::expand_sym_blocks(j->nact, P.m) <= j->partials_blocks;
        if (sym) { nblk = gk::expand_sym_blocks(j->nact, P.m); GK_CUDA(gk::launch_expand_sym(e, 0, j->nact, st)); }
        else if (j->world == 1) GK_CUDA(gk::launch_expand(e, 0, j->nact, st));
    }
    GK_CUDA(gk::launch_mean_finish(j->partials, nblk, j->scratch, j->sums_dev, st));
    GK_CUDA(cudaEventRecord(j->ev[1], st));
    j->launches += 3;
    GK_CUDA(cudaMemcpyAsync(j->sums_host, j->sums_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    j->ran = true;
    return GK_OK;
}

int gk_phi_run(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (!j->uploaded) return fail(GK_ERR_STATE, "gk_phi_run before gk_phi_upload");
    if (j->world > 1) return fail(GK_ERR_STATE, "a sharded job is driven step by step (gk_phi_run_step + the all-gather of gk_phi_dvec)");
    gk::PhiPlan &P = *j->planp;
    if (j->opts.verbose) {
        // same text as lib/compute.cpp:662-676
        for (int g = 0; g < P.G; g++) std::printf("Cut size %d/%d: %d\n", g + 1, P.G, (int) P.steps[g].cut_size);
        std::printf("Computing the kinship matrix:\n");
    }
    for (int g = 0; g < P.G; g++) { int rc = gk_phi_run_step(j, g); if (rc) return rc; }
    return gk_phi_finish(j);
}

int gk_phi_sync(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    return GK_OK;
}

int gk_phi_mean(gk_phi_job *j, double *sum_all, double *sum_diag, double *mean) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_mean before gk_phi_run");
    if (j->no_expand) return fail(GK_ERR_STATE, "the job was planned with GK_FLAG_DIAGONAL_ONLY: it has no m x m result");
    int rc = gk_phi_sync(j);
    if (rc) return rc;
    const double total = j->sums_host[0], diag = j->sums_host[1];
    if (sum_all) *sum_all = total;
    if (sum_diag) *sum_diag = diag;
    const double n = (double) j->planp->m;
    // geneakit/compute.py:131-132 (with world > 1 these are this shard's partial sums: the driver
    // all-reduces sum_all / sum_diag across ranks and applies the same formula)
    if (mean) *mean = (total - diag) / (n * n - n);
    return GK_OK;
}

int gk_phi_copy_out(gk_phi_job *j, double *host_out) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_copy_out before gk_phi_run");
    if (j->no_expand) return fail(GK_ERR_STATE, "the job was planned with GK_FLAG_DIAGONAL_ONLY: it has no m x m result");
    if (!host_out) return fail(GK_ERR_ARG, "null output");
    GK_CUDA(cudaSetDevice(j->device));
    const size_t bytes = (size_t) (j->row1 - j->row0) * j->planp->m * sizeof(double);
    GK_CUDA(cudaMemcpyAsync(host_out, j->out, bytes, cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    return GK_OK;
}

int gk_phi_copy_rows(gk_phi_job *j, int64_t row_begin, int64_t nrows, double *host_out) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_copy_rows before gk_phi_run");
    if (!host_out || row_begin < j->row0 || nrows < 0 || row_begin + nrows > j->row1) return fail(GK_ERR_ARG, "rows outside this job's block");
    if (j->no_expand) return fail(GK_ERR_STATE, "the job was planned with GK_FLAG_DIAGONAL_ONLY: it has no m x m result");
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaMemcpyAsync(host_out, j->out + (row_begin - j->row0) * (int64_t) j->planp->m, (size_t) nrows * j->planp->m * sizeof(double),
                            cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    return GK_OK;
}

int gk_phi_device_result(gk_phi_job *j, double **dev, int64_t *ld) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    if (dev) *dev = j->out;
    if (ld) *ld = j->planp->m;
    return GK_OK;
}

int gk_phi_get_stats(gk_phi_job *j, gk_phi_stats *s) {
    if (!j || !s) return fail(GK_ERR_ARG, "null argument");
    const gk::PhiPlan &P = *j->planp;
    std::memset(s, 0, sizeof *s);
    s->n_cuts = P.G; s->n_steps = P.G - 1; s->compressed = P.compressed; s->ranked = P.ranked;
    s->tile = gk::kPanel; s->m = P.m;
    for (int g = 0; g < P.G; g++) {
        const gk::StepPlan &S = P.steps[g];
        s->max_cut = std::max(s->max_cut, S.cut_size);
        s->max_classes = std::max(s->max_classes, S.K);
        if (g >= 1) {
            const gk::StepPlan &Sp = P.steps[g - 1];
            s->cells_reference += S.cut_size * S.cut_size;
            s->cells_device += S.K * S.K;
            s->algorithmic_bytes_reference += 8 * (S.P_ref * Sp.cut_size + S.cut_size * S.cut_size - S.n_kept * S.n_kept);
            s->algorithmic_bytes_device += 8 * (S.P_cls * Sp.K + S.K * S.K - S.kept_cls * S.kept_cls);
            s->panel_row_bytes += 8 * S.panel_rows * Sp.Kpad;
            s->corrections += (int64_t) S.pg_i.size();
        }
    }
    const int64_t mm = (int64_t) P.m * P.m, Kl = P.steps[P.G - 1].K;
    s->cells_device += mm;
    s->algorithmic_bytes_device += 8 * (mm + std::min(mm, Kl * Kl));
    s->device_bytes = j->device_bytes;
    s->kernel_launches = j->launches ? j->launches : (P.single_cut ? 3 : 2 * (P.G - 1) + 4);
    s->plan_seconds = P.plan_seconds;
    s->plan_bytes = plan_bytes(j);
    for (int64_t x : j->cut_row_bytes) s->rank_row_bytes += x;
    for (int64_t x : j->cut_remote_bytes) s->rank_remote_row_bytes += x;
    return GK_OK;
}

int gk_phi_cut_sizes(gk_phi_job *j, int32_t *sizes, int32_t cap) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    for (int g = 0; g < j->planp->G && g < cap; g++) sizes[g] = (int32_t) j->planp->steps[g].cut_size;
    return GK_OK;
}
int gk_phi_class_sizes(gk_phi_job *j, int32_t *sizes, int32_t cap) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    for (int g = 0; g < j->planp->G && g < cap; g++) sizes[g] = (int32_t) j->planp->steps[g].K;
    return GK_OK;
}

int gk_phi_step_times(gk_phi_job *j, float *ms, int32_t cap) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "no run recorded");
    int rc = gk_phi_sync(j);
    if (rc) return rc;
    const int G = j->planp->G;
    if (j->planp->single_cut) return GK_OK;
    for (int g = 1; g < G && g - 1 < cap; g++) GK_CUDA(cudaEventElapsedTime(&ms[g - 1], j->ev[2 * g], j->ev[2 * g + 1]));
    if (G - 1 < cap) GK_CUDA(cudaEventElapsedTime(&ms[G - 1], j->ev[0], j->ev[1]));   // expansion + mean
    return GK_OK;
}

int gk_phi_step_bytes(gk_phi_job *j, int64_t *bytes, int32_t cap) {
    if (!j || !bytes) return fail(GK_ERR_ARG, "null argument");
    const gk::PhiPlan &P = *j->planp;
    for (int g = 1; g < P.G && g - 1 < cap; g++) {
        const gk::StepPlan &S = P.steps[g], &Sp = P.steps[g - 1];
        bytes[g - 1] = 8 * (S.P_cls * Sp.K + S.K * S.K - S.kept_cls * S.kept_cls);
    }
    const int64_t mm = (int64_t) P.m * P.m, Kl = P.steps[P.G - 1].K;
    if (P.G - 1 < cap) bytes[P.G - 1] = 8 * (mm + std::min(mm, Kl * Kl));
    return GK_OK;
}

int gk_phi_step_view(gk_phi_job *j, int32_t step, gk_step_view *v) {
    if (!j || !v || step < 0 || step >= j->planp->G) return fail(GK_ERR_ARG, "bad step");
    const gk::PhiPlan &P = *j->planp;
    const gk::StepPlan &S = P.steps[step];
    std::memset(v, 0, sizeof *v);
    v->K = S.K; v->Kpad = S.Kpad; v->K_prev = S.K_prev; v->Kpad_prev = S.Kpad_prev;
    v->np = S.np; v->panel = gk::kPanel; v->ranked = P.ranked ? 1 : 0; v->ngroups = (int32_t) S.pg_i.size();
    v->panel_begin = j->off[step].pb; v->panel_end = j->off[step].pe;
    v->pF = S.pF.data(); v->pM = S.pM.data(); v->flags = S.flags.data();
    v->cls_ind = S.cls_ind.data(); v->cls_size = S.cls_size.data();
    v->pg_i = S.pg_i.data(); v->pg_j = S.pg_j.data(); v->pg_off = S.pg_off.data();
    v->pt_cls = S.pt_cls.data(); v->pt_mult = S.pt_mult.data();
    v->panel_rows = S.panel_rows;
    return GK_OK;
}
int gk_phi_final_view(gk_phi_job *j, const int32_t **cls, const int32_t **ind, int64_t *K_last) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (cls) *cls = j->planp->final_cls.data();
    if (ind) *ind = j->planp->final_ind.data();
    if (K_last) *K_last = j->planp->steps[j->planp->G - 1].K;
    return GK_OK;
}

int gk_phi_debug_prof(gk_phi_job *j, uint64_t *out16) {
    if (!j || !j->uploaded || !out16) return fail(GK_ERR_ARG, "bad argument");
    std::memset(out16, 0, 16 * sizeof(uint64_t));
    if (!j->prof) return GK_OK;
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    GK_CUDA(cudaMemcpy(out16, j->prof, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return GK_OK;
}

void gk_phi_free(gk_phi_job *j) { destroy(j); }

// Self kinship of every caller entry (the diagonal of the result, lib/compute.cpp:553-568) from the device's d vector.
int gk_phi_diag(gk_phi_job *j, double *host_out) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_diag before gk_phi_run");
    if (!host_out) return fail(GK_ERR_ARG, "null output");
    const gk::PhiPlan &P = *j->planp;
    GK_CUDA(cudaSetDevice(j->device));
    if (P.single_cut) { for (int32_t i = 0; i < P.m; i++) host_out[i] = 0.5; return GK_OK; }
    const int last = P.G - 1;
    std::vector<double> d((size_t) P.steps[last].Kpad);
    GK_CUDA(cudaMemcpyAsync(d.data(), j->dvec[last & 1], d.size() * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    for (int32_t i = 0; i < P.m; i++) host_out[i] = d[P.final_cls[i]];
    return GK_OK;
}

// Inbreeding coefficients F = 2 phi(i, i) - 1 = phi(father, mother) (gen.f, geneakit/compute.py:255-282; the reference
// computes them with a separate algorithm, lib/compute.cpp:725-814).
int gk_phi_f(gk_phi_job *j, double *host_out) {
    int rc = gk_phi_diag(j, host_out);
    if (rc) return rc;
    for (int32_t i = 0; i < j->planp->m; i++) host_out[i] = 2.0 * host_out[i] - 1.0;
    return GK_OK;
}

namespace {
// count + fill of the selected pairs of this job's result rows; the arrays stay on the device until fetched
struct Selection {
    int64_t *cnt = nullptr, *off = nullptr;
    int32_t *oi = nullptr, *oj = nullptr;
    double *ov = nullptr;
    float *ovf = nullptr;
    int64_t total = 0;
    int mode = 0;
    std::vector<int64_t> host_off;
    void release() { pool_free(cnt); pool_free(off); pool_free(oi); pool_free(oj); pool_free(ov); pool_free(ovf); *this = Selection(); }
};
thread_local Selection g_sel;

int run_selection(gk_phi_job *j, int mode, double thr, int64_t *total) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "no result: run the job first");
    if (j->no_expand) return fail(GK_ERR_STATE, "the job was planned with GK_FLAG_DIAGONAL_ONLY: it has no m x m result");
    GK_CUDA(cudaSetDevice(j->device));
    g_sel.release();
    g_sel.mode = mode;
    const int64_t rows = j->row1 - j->row0;
    gk::SelectDev q{};
    q.M = j->out; q.m = j->planp->m; q.row0 = j->row0; q.row1 = j->row1; q.mode = mode; q.thr = thr;
    GK_CUDA(pool_alloc_t(&g_sel.cnt, (size_t) std::max<int64_t>(rows, 1) * 8, j->device));
    GK_CUDA(pool_alloc_t(&g_sel.off, (size_t) std::max<int64_t>(rows, 1) * 8, j->device));
    q.row_count = g_sel.cnt; q.row_off = g_sel.off;
    GK_CUDA(gk::launch_select(q, 0, j->stream));
    std::vector<int64_t> cnt((size_t) rows);
    GK_CUDA(cudaMemcpyAsync(cnt.data(), g_sel.cnt, (size_t) rows * 8, cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    g_sel.host_off.assign((size_t) rows + 1, 0);
    for (int64_t r = 0; r < rows; r++) g_sel.host_off[r + 1] = g_sel.host_off[r] + cnt[r];
    g_sel.total = g_sel.host_off[rows];
    GK_CUDA(cudaMemcpyAsync(g_sel.off, g_sel.host_off.data(), (size_t) rows * 8, cudaMemcpyHostToDevice, j->stream));
    const size_t n = (size_t) std::max<int64_t>(g_sel.total, 1);
    GK_CUDA(pool_alloc_t(&g_sel.oj, n * 4, j->device));
    if (mode == 0) { GK_CUDA(pool_alloc_t(&g_sel.oi, n * 4, j->device)); GK_CUDA(pool_alloc_t(&g_sel.ov, n * 8, j->device)); }
    else GK_CUDA(pool_alloc_t(&g_sel.ovf, n * 4, j->device));
    q.oi = g_sel.oi; q.oj = g_sel.oj; q.ov = g_sel.ov; q.ovf = g_sel.ovf;
    GK_CUDA(gk::launch_select(q, 1, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    if (total) *total = g_sel.total;
    return GK_OK;
}
}  // namespace

// gen.phiOver on the device (geneakit/compute.py:135-190, dense branch): pairs (i > j) with kinship > threshold, in the
// reference's order (row by row).  gk_phi_over counts and keeps the pairs in HBM; gk_phi_over_fetch copies the `count`
// triples (row index, column index in CALLER order, value) to the host and frees them.
int gk_phi_over(gk_phi_job *j, double threshold, int64_t *count) { return run_selection(j, 0, threshold, count); }
int gk_phi_over_fetch(gk_phi_job *j, int32_t *pro1_idx, int32_t *pro2_idx, double *kinship) {
    if (!j || g_sel.mode != 0 || !g_sel.oj) return fail(GK_ERR_STATE, "gk_phi_over_fetch without gk_phi_over");
    GK_CUDA(cudaSetDevice(j->device));
    if (g_sel.total > 0) {
        GK_CUDA(cudaMemcpy(pro1_idx, g_sel.oi, (size_t) g_sel.total * 4, cudaMemcpyDeviceToHost));
        GK_CUDA(cudaMemcpy(pro2_idx, g_sel.oj, (size_t) g_sel.total * 4, cudaMemcpyDeviceToHost));
        GK_CUDA(cudaMemcpy(kinship, g_sel.ov, (size_t) g_sel.total * 8, cudaMemcpyDeviceToHost));
    }
    g_sel.release();
    return GK_OK;
}

// The CSR that gen.phi(sparse=True) returns (compute_kinships_sparse, lib/compute.cpp:272-487; binding
// lib/cgeneakit.cpp:433-486): lower triangle incl. the diagonal, float32 values (float) (2 phi) / 2, entries with
// |2 phi| > 1e-9, rows and columns in caller order -- here compacted on the device from the dense result.
int gk_phi_sparse(gk_phi_job *j, int64_t *nnz) { return run_selection(j, 1, 1e-9, nnz); }
int gk_phi_sparse_fetch(gk_phi_job *j, int64_t *indptr, int32_t *indices, float *data) {
    if (!j || g_sel.mode != 1 || !g_sel.oj) return fail(GK_ERR_STATE, "gk_phi_sparse_fetch without gk_phi_sparse");
    GK_CUDA(cudaSetDevice(j->device));
    std::memcpy(indptr, g_sel.host_off.data(), g_sel.host_off.size() * 8);
    if (g_sel.total > 0) {
        GK_CUDA(cudaMemcpy(indices, g_sel.oj, (size_t) g_sel.total * 4, cudaMemcpyDeviceToHost));
        GK_CUDA(cudaMemcpy(data, g_sel.ovf, (size_t) g_sel.total * 4, cudaMemcpyDeviceToHost));
    }
    g_sel.release();
    return GK_OK;
}

// gen.phiCI's bootstrap statistic on the device (geneakit/compute.py:193-252): for each of the b resamples idx[r][0..n)
// (indices into the result, with replacement) the mean of the entries < 0.5 of phi[idx][:, idx]; NaN if there is none.
int gk_phi_ci_stats(gk_phi_job *j, const int32_t *idx, int32_t b, int32_t n, double *stats) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "no result: run the job first");
    if (j->no_expand || j->world != 1) return fail(GK_ERR_STATE, "gk_phi_ci_stats needs the whole m x m result on one GPU");
    if (!idx || !stats || b < 0 || n < 0) return fail(GK_ERR_ARG, "bad argument");
    for (int64_t k = 0; k < (int64_t) b * n; k++) if (idx[k] < 0 || idx[k] >= j->planp->m) return fail(GK_ERR_INDEX, "resample index out of range");
    GK_CUDA(cudaSetDevice(j->device));
    int32_t *didx = nullptr;
    double *dstat = nullptr;
    GK_CUDA(pool_alloc_t(&didx, (size_t) std::max<int64_t>((int64_t) b * n, 1) * 4, j->device));
    cudaError_t e = pool_alloc_t(&dstat, (size_t) std::max(b, 1) * 8, j->device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(didx, idx, (size_t) b * n * 4, cudaMemcpyHostToDevice, j->stream);
    if (e == cudaSuccess) e = gk::launch_ci(j->out, j->planp->m, didx, b, n, dstat, j->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(stats, dstat, (size_t) b * 8, cudaMemcpyDeviceToHost, j->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(j->stream);
    pool_free(didx); pool_free(dstat);
    if (e != cudaSuccess) return cuda_fail(e, "gk_phi_ci_stats");
    return GK_OK;
}

// Plan again from the inputs the job was created with and copy the plan arrays to the device once more.  Blocking; with
// gk_phi_run it is one "flattened arrays on the host -> matrix resident in HBM" pass (SURVEY 8d's t_phi) on buffers that
// are already allocated -- gk_phi_pass does the same with the planning overlapped with the device's work.
int gk_phi_replan(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "gk_phi_replan before gk_phi_upload");
    if (!j->inputs) return fail(GK_ERR_STATE, "the job does not hold its inputs");
    auto fresh = std::make_shared<gk::PhiPlan>();
    const gk_plan_inputs &in = *j->inputs;
    {
        gk::PhiPlanner pl(in.n, in.sire.data(), in.dam.data(), in.m, in.pro.data(), j->opts.compress, j->opts.order, j->world, *fresh, planner_threads());
        bool ok = pl.begin();
        if (ok) { pl.start(); ok = pl.finish(); }
        if (!ok) return plan_error(*fresh);
    }
    j->planp = fresh;
    return gk_phi_reupload(j);
}

// (multi-GPU jobs of one process share the plan: rank 0 replans, the others take its plan and upload their own view)
int gk_phi_reupload(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "gk_phi_reupload before gk_phi_upload");
    GK_CUDA(cudaSetDevice(j->device));
    configure_job(j);
    int rc = alloc_buffers(j);
    if (!rc) rc = upload_plan(j);
    if (rc) return rc;
    GK_CUDA(cudaStreamSynchronize(j->copy_stream));
    j->ran = false; j->steps_done = 0;
    return GK_OK;
}

// Another rank of the same job in THIS process: shares the plan (planned once for `world` ranks) and the inputs.
int gk_phi_clone_rank(gk_phi_job *src, int32_t rank, int32_t device, gk_phi_job **out) {
    if (!src || !out) return fail(GK_ERR_ARG, "null argument");
    if (rank < 0 || rank >= src->world) return fail(GK_ERR_ARG, "rank outside [0, world)");
    gk_phi_job *j = new gk_phi_job();
    j->planp = src->planp; j->inputs = src->inputs; j->opts = src->opts;
    j->opts.rank = rank; j->opts.device = device; j->opts.stream = nullptr;
    j->rank = rank; j->world = src->world;
    configure_job(j);
    *out = j;
    return GK_OK;
}

}  // extern "C"

// ---- a PASS: plan + copy + run with the three overlapped ------------------------------------------------------------
// The planner finds the cuts and their classes (all sizes), then plans the cuts in order on its thread pool; a worker
// packs each finished cut into the pinned segment of every job that shares the plan; the driver (this thread) copies
// segment g and launches step g as soon as cut g is ready, while the later cuts are still being planned.
namespace {

// Start planning `jobs` (the ranks of one process that share a plan; jobs[0] holds the inputs) again from their inputs.
// After it returns the sizes are known, every job is configured and has its buffers; the workers are running.
int pass_begin(const std::vector<gk_phi_job *> &jobs, int compress) {
    gk_phi_job *j0 = jobs[0];
    if (!j0->inputs) return fail(GK_ERR_ARG, "gk_phi: empty pedigree or proband list (the binding substitutes get_proband_ids for an empty list)");
    const gk_plan_inputs &in = *j0->inputs;
    j0->planner.reset();
    j0->pass_t0 = now_s();
    j0->pass_plan = std::make_shared<gk::PhiPlan>();
    // (a rank that plans for every rank of a process-per-GPU job has the host to itself: the others do not plan)
    const int threads = j0->shadows.empty() ? planner_threads() : (int) std::max(1u, std::thread::hardware_concurrency());
    j0->planner.reset(new gk::PhiPlanner(in.n, in.sire.data(), in.dam.data(), in.m, in.pro.data(), compress, j0->opts.order, j0->world,
                                         *j0->pass_plan, threads));
    if (!j0->planner->begin()) { const int rc = plan_error(*j0->pass_plan); j0->planner.reset(); return rc; }
    for (gk_phi_job *j : jobs) {
        if (j->shadow) { j->planp = j0->pass_plan; configure_job(j); j->pack_error.clear(); continue; }
        int rc = prepare_device(j);
        if (rc) { j0->planner.reset(); return rc; }
        j->planp = j0->pass_plan;
        configure_job(j);
        rc = alloc_buffers(j);
        if (!rc) rc = begin_copies(j);
        if (rc) { j0->planner.reset(); return rc; }
        j->pack_error.clear();
        j->uploaded = true; j->ran = false; j->steps_done = 0;
    }
    j0->pass_prefix_s = now_s() - j0->pass_t0;
    std::vector<gk_phi_job *> js = jobs;
    j0->planner->start([js](int g) { for (gk_phi_job *j : js) pack_cut(j, g); });
    return GK_OK;
}

// Cut g is planned: copy its segment (and, with the last cut, the final maps) to every job's device.
int pass_feed(const std::vector<gk_phi_job *> &jobs, int g) {
    gk_phi_job *j0 = jobs[0];
    if (!j0->planner) return fail(GK_ERR_STATE, "no pass in flight");
    const int G = j0->planp->G;
    if (g < 0 || g >= G) return fail(GK_ERR_ARG, "bad step");
    if (!j0->planner->wait_cut(g)) {
        j0->planner->finish();
        const int rc = j0->pass_plan->error_code == 3 ? -3 : plan_error(*j0->pass_plan);     // -3: plan again uncompressed
        return rc;
    }
    // the other ranks' views, straight into their devices (before this rank's own segment: its event covers them all)
    for (gk_phi_job *sh : j0->shadows)
        for (int sg = g; sg <= (g == G - 1 ? G : g); sg++) {
            const PlanBlob &b = sh->blob[sg];
            if (!b.packed) return fail(GK_ERR_STATE, sh->pack_error.empty() ? "internal: plan segment not packed" : sh->pack_error);
            if ((int) j0->peer_seg.size() <= sh->rank || (int) j0->peer_seg[sh->rank].size() <= sg || !j0->peer_seg[sh->rank][sg] ||
                j0->peer_seg_cap[sh->rank][sg] < b.n)
                return fail(GK_ERR_STATE, "a peer's plan segment is not mapped or too small (gk_phi_seg_ipc_open)");
            GK_CUDA(cudaSetDevice(j0->device));
            GK_CUDA(cudaMemcpyAsync(j0->peer_seg[sh->rank][sg], b.p, b.n * sizeof(int32_t), cudaMemcpyDefault, j0->copy_stream));
        }
    for (gk_phi_job *j : jobs) {
        if (j->shadow) continue;
        int rc = feed_segment(j, g);
        if (!rc && g == G - 1) rc = feed_segment(j, G);
        if (rc) return rc;
    }
    return GK_OK;
}

int pass_end(gk_phi_job *j0) {
    if (!j0->planner) return GK_OK;
    const bool ok = j0->planner->finish();
    j0->planner.reset();
    return ok ? GK_OK : plan_error(*j0->pass_plan);
}

void drain(const std::vector<gk_phi_job *> &jobs) {
    for (gk_phi_job *j : jobs) {
        if (!j->device_ready) continue;
        cudaSetDevice(j->device);
        if (j->stream && cudaStreamSynchronize(j->stream) != cudaSuccess) cudaGetLastError();
        if (j->copy_stream && cudaStreamSynchronize(j->copy_stream) != cudaSuccess) cudaGetLastError();
    }
}

}  // namespace

extern "C" {

// One pass of a ONE-GPU job from its inputs: plan, copy, all cuts, expansion -- asynchronous like gk_phi_run (the planning
// itself has finished when it returns; the kernels of the last cuts may still be running).
int gk_phi_pass(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (j->world > 1) return fail(GK_ERR_STATE, "a sharded job is driven step by step (gk_phi_pass_begin / gk_phi_pass_step + gk_phi_run_step)");
    std::vector<gk_phi_job *> jobs(1, j);
    int rc = pass_begin(jobs, j->opts.compress);
    if (rc) return rc;
    const int G = j->planp->G;
    for (int g = 0; g < G && !rc; g++) {
        rc = pass_feed(jobs, g);
        if (!rc) rc = gk_phi_run_step(j, g);
    }
    if (!rc) rc = gk_phi_finish(j);
    if (rc) { drain(jobs); j->planner.reset(); return rc < 0 ? fail(GK_ERR_STATE, j->pass_plan->error) : rc; }
    return pass_end(j);
}

// The same for the rank of a sharded job that its driver runs cut by cut: begin, then per cut gk_phi_pass_step (waits for
// the cut's plan, enqueues its copy) before gk_phi_run_step, and gk_phi_pass_end after gk_phi_finish.
int gk_phi_pass_begin(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (!j->uploaded) return fail(GK_ERR_STATE, "gk_phi_pass_begin before gk_phi_upload (peers are wired to the uploaded buffers)");
    double *b0 = j->buf[0], *b1 = j->buf[1];
    std::vector<gk_phi_job *> jobs(1, j);
    int rc = pass_begin(jobs, j->opts.compress);
    if (!rc && (j->buf[0] != b0 || j->buf[1] != b1)) { drain(jobs); j->planner.reset(); return fail(GK_ERR_STATE, "internal: the buffers moved between two passes of the same inputs"); }
    return rc;
}
int gk_phi_pass_step(gk_phi_job *j, int32_t g) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    std::vector<gk_phi_job *> jobs(1, j);
    jobs.insert(jobs.end(), j->shadows.begin(), j->shadows.end());
    const int rc = pass_feed(jobs, g);
    if (rc) { drain(jobs); j->planner.reset(); return rc < 0 ? fail(GK_ERR_STATE, j->pass_plan->error) : rc; }
    return GK_OK;
}
int gk_phi_pass_end(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    return pass_end(j);
}

// One process per GPU, ONE planner: rank 0 plans every pass for all the ranks (the box's host threads plan once, not
// `world` times) and copies each rank's view of a cut straight into that rank's device segments (CUDA-IPC mappings of the
// segments the ranks allocated at upload).  The other ranks only enqueue their kernels; what orders a copy before the
// kernel that reads it is the per-cut collective (see geneakit_b200/distributed.py).
int gk_phi_seg_count(gk_phi_job *j, int32_t *n) {
    if (!j || !j->uploaded || !n) return fail(GK_ERR_STATE, "job not uploaded");
    *n = (int32_t) j->seg.size();
    return GK_OK;
}
int gk_phi_seg_ipc_export(gk_phi_job *j, int32_t g, void *handle, int64_t *cap_ints) {
    if (!j || !j->uploaded || !handle || g < 0 || g >= (int) j->seg.size() || !j->seg[g]) return fail(GK_ERR_STATE, "no such plan segment");
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle), j->seg[g]));
    if (cap_ints) *cap_ints = (int64_t) j->seg_cap[g];
    return GK_OK;
}
int gk_phi_seg_ipc_open(gk_phi_job *j, int32_t q, int32_t g, const void *handle, int64_t cap_ints) {
    if (!j || !j->uploaded || !handle) return fail(GK_ERR_STATE, "job not uploaded");
    if (q < 0 || q >= j->world || q == j->rank || g < 0 || g >= (int) j->seg.size()) return fail(GK_ERR_ARG, "bad peer rank or segment");
    GK_CUDA(cudaSetDevice(j->device));
    if (j->peer_seg.empty()) {
        j->peer_seg.assign((size_t) j->world, std::vector<int32_t *>(j->seg.size(), nullptr));
        j->peer_seg_cap.assign((size_t) j->world, std::vector<size_t>(j->seg.size(), 0));
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void *ptr = nullptr;
    GK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    j->peer_seg[q][g] = static_cast<int32_t *>(ptr);
    j->peer_seg_cap[q][g] = (size_t) cap_ints;
    return GK_OK;
}
int gk_phi_pass_begin_all(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (!j->uploaded || j->world < 2) return fail(GK_ERR_STATE, "gk_phi_pass_begin_all: an uploaded rank of a sharded job");
    if (j->shadows.empty())
        for (int q = 0; q < j->world; q++) {
            if (q == j->rank) continue;
            gk_phi_job *sh = new gk_phi_job();
            sh->inputs = j->inputs; sh->opts = j->opts; sh->opts.rank = q; sh->opts.stream = nullptr;
            sh->rank = q; sh->world = j->world; sh->shadow = true;
            j->shadows.push_back(sh);
        }
    double *b0 = j->buf[0], *b1 = j->buf[1];
    std::vector<gk_phi_job *> jobs(1, j);
    jobs.insert(jobs.end(), j->shadows.begin(), j->shadows.end());
    int rc = pass_begin(jobs, j->opts.compress);
    if (!rc && (j->buf[0] != b0 || j->buf[1] != b1)) { drain(jobs); j->planner.reset(); return fail(GK_ERR_STATE, "internal: the buffers moved between two passes of the same inputs"); }
    return rc;
}

}  // extern "C"

// ---- one process, several GPUs (SURVEY 8b "Threading", 5 "distributed backend"): what gen.phi() itself uses ----------
// The ranks of the sharded algorithm (DESIGN.md "Multi-GPU") as jobs of THIS process, one per device, one plan shared by
// all of them.  Peers are plain pointers (cudaDeviceEnablePeerAccess), the self-kinship vector is written into every
// rank's replica by the diagonal kernel, and the barrier between cuts is an event every stream waits on -- no host
// thread per GPU, no NCCL.  The result rows go from every GPU straight into the caller's one m x m host matrix, in parallel.
namespace {

struct MultiPhi {
    std::vector<gk_phi_job *> jobs;
    std::vector<cudaEvent_t> cut_done;       // one per rank
    bool virtual_devices = false;             // tests on a one-GPU box: every rank on the current device, ONE stream
    ~MultiPhi() {
        if (!jobs.empty()) jobs[0]->planner.reset();         // the planner threads pack into every rank's blobs
        for (size_t q = 0; q < cut_done.size(); q++) if (cut_done[q]) { cudaSetDevice(jobs[q]->device); cudaEventDestroy(cut_done[q]); }
        for (size_t q = jobs.size(); q-- > 0;) destroy(jobs[q]);      // rank 0 last: with virtual devices the others enqueue on its stream
    }
};

int multi_devices(const gk_opts *o, int *ndev, bool *virt) {
    int want = (o && o->world > 1 && o->rank < 0) ? o->world : 0;
    if (!want) if (const char *e = std::getenv("GENEAKIT_GPUS")) want = std::atoi(e);
    *virt = false;
    if (const char *e = std::getenv("GENEAKIT_VIRTUAL_DEVICES")) *virt = e[0] == '1';
    if (want <= 1) { *ndev = 1; return GK_OK; }
    if (want > gk::kMaxShards) return fail(GK_ERR_ARG, "at most 8 GPUs");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return fail(GK_ERR_CUDA, "no usable CUDA device; geneakit_b200 has no CPU fallback"); }
    if (!*virt && want > count) want = count;         // use the GPUs the box has
    *ndev = want;
    return GK_OK;
}

// D2H of a one-GPU result.  The matrix is symmetric: only the block rows of the lower triangle cross PCIe (a little more
// than half of the bytes), and host threads mirror each block into the upper triangle while the next ones are in flight.
int copy_out_symmetric(gk_phi_job *j, double *out);

// rc = -3: the sibship compression cannot be used for this pedigree (the caller plans again uncompressed)
int phi_pipelined(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro, double *out,
                  double *mean_or_null, const gk_opts *opts, int ndev, bool virt, int compress) {
    MultiPhi M;
    M.virtual_devices = virt;
    gk_opts o{};
    if (opts) std::memcpy(&o, opts, std::min<size_t>(sizeof(gk_opts), (size_t) std::max(0, opts->struct_size)));
    else { o.device = -1; o.compress = -1; o.order = -1; }
    int base = 0;
    if (virt || ndev == 1) { if (o.device >= 0) base = o.device; else if (cudaGetDevice(&base) != cudaSuccess) { cudaGetLastError(); base = 0; } }
    o.struct_size = sizeof(gk_opts); o.world = ndev; o.rank = 0; o.compress = compress;
    o.device = (virt || ndev == 1) ? base : 0;
    if (ndev > 1) o.stream = nullptr;
    const double t0 = now_s();
    gk_phi_job *j0 = new_job(n, sire, dam, m, pro, &o);
    if (!j0) return GK_ERR_ARG;
    M.jobs.push_back(j0);
    for (int q = 1; q < ndev; q++) {
        gk_phi_job *jq = new gk_phi_job();
        jq->opts = j0->opts; jq->opts.rank = q; jq->opts.device = virt ? base : q; jq->opts.stream = nullptr;
        jq->rank = q; jq->world = ndev; jq->inputs = j0->inputs;
        M.jobs.push_back(jq);
    }
    if (virt && ndev > 1) {
        // one stream for every rank: persistent kernels of different ranks must never be co-scheduled on one GPU
        int rc = prepare_device(j0);
        if (rc) return rc;
        for (int q = 1; q < ndev; q++) M.jobs[q]->opts.stream = j0->stream;
    }
    int rc = pass_begin(M.jobs, compress);           // cuts and classes found, buffers allocated on every device, workers running
    if (rc) return rc;
    const double t1 = now_s();
    if (ndev > 1) {
        if (!virt)
            for (int q = 0; q < ndev; q++) {
                GK_CUDA(cudaSetDevice(M.jobs[q]->device));
                for (int p = 0; p < ndev; p++) {
                    if (p == q) continue;
                    cudaError_t e = cudaDeviceEnablePeerAccess(M.jobs[p]->device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess (the GPUs of a multi-GPU job must be NVLink / P2P peers)");
                }
            }
        for (int q = 0; q < ndev; q++) {
            gk_phi_job *j = M.jobs[q];
            int k = 0;
            for (int p = 0; p < ndev; p++) {
                if (p == q) continue;
                j->peer[p][0] = M.jobs[p]->buf[0]; j->peer[p][1] = M.jobs[p]->buf[1];
                j->dvec_peer[k][0] = M.jobs[p]->dvec[0]; j->dvec_peer[k][1] = M.jobs[p]->dvec[1];
                k++;
            }
            j->n_dvec_peer = k;
            j->peers_set = true;
        }
        M.cut_done.assign(ndev, nullptr);
        for (int q = 0; q < ndev; q++) {
            GK_CUDA(cudaSetDevice(M.jobs[q]->device));
            GK_CUDA(cudaEventCreateWithFlags(&M.cut_done[q], cudaEventDisableTiming));
        }
    }
    const int G = j0->planp->G;
    if (o.verbose) {
        for (int g = 0; g < G; g++) std::printf("Cut size %d/%d: %d\n", g + 1, G, (int) j0->planp->steps[g].cut_size);
        std::printf("Computing the kinship matrix:\n");
    }
    for (int g = 0; g < G; g++) {
        rc = pass_feed(M.jobs, g);
        if (rc) { drain(M.jobs); return rc; }
        for (int q = 0; q < ndev; q++) { rc = gk_phi_run_step(M.jobs[q], g); if (rc) { drain(M.jobs); return rc; } }
        // the barrier between cuts: every rank's step g (rows + self kinships written to every replica) before any step g + 1;
        // ranked jobs then pull their mirror tiles out of the higher ranks' rows, and a second barrier completes the cut
        for (int round = 0; round < (ndev > 1 && j0->planp->ranked && g > 0 ? 2 : 1); round++) {
            if (round == 1) for (int q = 0; q < ndev; q++) { rc = gk_phi_run_pull(M.jobs[q], g); if (rc) { drain(M.jobs); return rc; } }
            if (ndev > 1 && !virt) {
                for (int q = 0; q < ndev; q++) { GK_CUDA(cudaSetDevice(M.jobs[q]->device)); GK_CUDA(cudaEventRecord(M.cut_done[q], M.jobs[q]->stream)); }
                for (int q = 0; q < ndev; q++) {
                    GK_CUDA(cudaSetDevice(M.jobs[q]->device));
                    for (int p = 0; p < ndev; p++) if (p != q) GK_CUDA(cudaStreamWaitEvent(M.jobs[q]->stream, M.cut_done[p], 0));
                }
            }
        }
    }
    rc = pass_end(j0);
    if (rc) { drain(M.jobs); return rc; }
    const double t2 = now_s();                       // planning is over; the device still works on the last cuts
    for (int q = 0; q < ndev; q++) { rc = gk_phi_finish(M.jobs[q]); if (rc) { drain(M.jobs); return rc; } }
    double t3 = t2;
    if (ndev == 1) {
        rc = copy_out_symmetric(j0, out);
        if (rc) { drain(M.jobs); return rc; }
        t3 = g_phi_timing[2];                        // set by copy_out_symmetric: when the kernels were done
    } else {
        for (int q = 0; q < ndev; q++) { GK_CUDA(cudaSetDevice(M.jobs[q]->device)); GK_CUDA(cudaStreamSynchronize(M.jobs[q]->stream)); }
        t3 = now_s();
        // result rows: every GPU copies its block into the caller's matrix (its own PCIe link), all at once
        for (int q = 0; q < ndev; q++) {
            gk_phi_job *j = M.jobs[q];
            GK_CUDA(cudaSetDevice(j->device));
            const size_t bytes = (size_t) (j->row1 - j->row0) * m * sizeof(double);
            if (bytes) GK_CUDA(cudaMemcpyAsync(out + (size_t) j->row0 * m, j->out, bytes, cudaMemcpyDeviceToHost, j->stream));
        }
    }
    double total = 0.0, diag = 0.0;
    for (int q = 0; q < ndev; q++) {
        gk_phi_job *j = M.jobs[q];
        GK_CUDA(cudaSetDevice(j->device));
        GK_CUDA(cudaStreamSynchronize(j->stream));        // (a peer may still be reading this rank's rows in its expansion: every stream is drained before anything is freed)
        total += j->sums_host[0]; diag += j->sums_host[1];
    }
    if (mean_or_null) *mean_or_null = (total - diag) / ((double) m * (double) m - (double) m);
    if (virt) GK_CUDA(cudaSetDevice(base));
    const double t4 = now_s();
    // plan prefix (cuts, classes) + allocation; cuts planned / copied / launched; until the kernels were done; D2H; total
    g_phi_timing[0] = j0->pass_prefix_s; g_phi_timing[1] = (t1 - t0) - j0->pass_prefix_s; g_phi_timing[2] = t3 - t1; g_phi_timing[3] = t4 - t3;
    g_phi_timing[4] = t4 - t0; g_phi_timing[5] = ndev;
    return GK_OK;
}

int copy_out_symmetric(gk_phi_job *j, double *out) {
    const int64_t m = j->planp->m;
    GK_CUDA(cudaSetDevice(j->device));
    const char *env = std::getenv("GK_D2H_SYMMETRIC");
    // default: matrices of 4096 probands and more; GK_D2H_SYMMETRIC=0 never, =1 always (tests)
    const bool sym = !(env && env[0] == '0') && (m >= 4096 || (env && env[0] == '1' && m >= 2)) && !j->planp->single_cut;
    if (!sym) {
        GK_CUDA(cudaStreamSynchronize(j->stream));
        g_phi_timing[2] = now_s();
        GK_CUDA(cudaMemcpyAsync(out, j->out, (size_t) m * m * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
        GK_CUDA(cudaStreamSynchronize(j->stream));
        return GK_OK;
    }
    // Rows [0, split) arrive as block rows of the lower triangle (rows [r0, r1), columns [0, r1)) and get their upper part from
    // the host mirror; rows [split, m) arrive whole (GK_D2H_MIRROR_ROWS = split / m).  Measured on two 16-core hosts (c4, 80 GB):
    // whole matrix over the link 1.40 - 1.56 s; split 0.5 .. 0.8: 1.07 - 1.22 s; everything mirrored: 0.95 - 1.05 s (the link
    // and the mirror threads share the host's memory system, a split does not help) -> 1.0.  Blocks of equal copied AREA.
    double frac = 1.0;
    if (const char *e = std::getenv("GK_D2H_MIRROR_ROWS")) frac = std::min(1.0, std::max(0.0, std::atof(e)));
    const int64_t split = std::min<int64_t>(m, (int64_t) (frac * (double) m) / 64 * 64);
    const int64_t nb_want = std::min<int64_t>(64, (m + 255) / 256);
    std::vector<int64_t> rb(1, 0);
    {
        const double area = 0.5 * (double) split * (double) split + (double) (m - split) * (double) m, per = area / (double) nb_want;
        const int64_t nlow = split > 0 ? std::max<int64_t>(1, (int64_t) (0.5 * (double) split * (double) split / per + 0.5)) : 0;
        for (int64_t k = 1; k < nlow; k++) {
            const int64_t r = (int64_t) ((double) split * std::sqrt((double) k / (double) nlow)) / 64 * 64;
            if (r > rb.back()) rb.push_back(r);
        }
        if (split > rb.back()) rb.push_back(split);
        const int64_t nhigh = m > split ? std::max<int64_t>(1, nb_want - nlow) : 0;
        for (int64_t k = 1; k <= nhigh; k++) {
            const int64_t r = k == nhigh ? m : split + (m - split) * k / nhigh / 64 * 64;
            if (r > rb.back()) rb.push_back(r);
        }
    }
    const int64_t nb = (int64_t) rb.size() - 1;
    std::vector<cudaEvent_t> ev((size_t) nb, nullptr);
    cudaEvent_t kernels_done = nullptr;
    GK_CUDA(cudaEventCreateWithFlags(&kernels_done, cudaEventDisableTiming));
    GK_CUDA(cudaEventRecord(kernels_done, j->stream));
    int rc = GK_OK;
    for (int64_t k = 0; k < nb && !rc; k++) {
        const int64_t r0 = rb[k], r1 = rb[k + 1], width = r0 >= split ? m : r1;
        cudaError_t e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(out + r0 * m, (size_t) m * sizeof(double), j->out + r0 * m, (size_t) m * sizeof(double),
                                  (size_t) width * sizeof(double), (size_t) (r1 - r0), cudaMemcpyDeviceToHost, j->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[k], j->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "D2H of the result");
    }
    if (!rc) {
        // mirror: out[c][a] = out[a][c] for the rows a of a landed block and c < a.  Worker threads take tiles of 64 x 64: the
        // tile is transposed into a thread-local buffer (contiguous 512-byte reads) and leaves as 512-byte runs of streaming
        // stores (no read-for-ownership of the destination lines).  The threads live for the whole copy; block k's tiles
        // are released when its copy has completed.
        const int nt = std::max(1, std::min(planner_threads(), 32));
        const bool skip = std::getenv("GK_D2H_DEBUG_NOMIRROR") != nullptr;       // timing experiments only (wrong upper triangle)
        std::unique_ptr<std::atomic<int64_t>[]> next(new std::atomic<int64_t>[(size_t) nb]);
        std::unique_ptr<std::atomic<int>[]> landed(new std::atomic<int>[(size_t) nb]);
        for (int64_t k = 0; k < nb; k++) { next[k].store(0); landed[k].store(0); }
        // tile = 16 landed rows x 256 columns: 2 KB sequential reads, 128-byte runs of streaming stores (measured on the host:
        // 3 - 6 x the rate of 64 x 64 tiles, whose transposed writes into the local buffer collide on a few L1 sets)
        constexpr int64_t TA = 16, TC = 256;
        auto work = [&]() {
            alignas(64) double loc[TA * TC];
            for (int64_t k = 0; k < nb; k++) {
                int spins = 0, st;
                while ((st = landed[k].load(std::memory_order_acquire)) == 0) { if (++spins < 256) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(50)); }
                if (st < 0) return;
                const int64_t r0 = rb[k], r1 = rb[k + 1];
                const int64_t cmax = std::min(r1, split);                 // destination rows: above the split the rows arrive whole
                if (r1 <= r0 || cmax <= 0 || skip) continue;
                const int64_t ntr = (r1 - r0 + TA - 1) / TA, ntc = (cmax + TC - 1) / TC, ntiles = ntr * ntc;
                for (int64_t t = next[k].fetch_add(1); t < ntiles; t = next[k].fetch_add(1)) {
                    const int64_t tr = t % ntr, tc = t / ntr;             // consecutive tiles: the same 256 destination rows
                    const int64_t a0 = r0 + tr * TA, a1 = std::min(a0 + TA, r1), c0 = tc * TC, c1 = std::min(c0 + TC, cmax);
                    if (c0 >= a1) continue;                               // entirely on the other side of the diagonal
                    const int64_t na = a1 - a0, nc = c1 - c0;
                    for (int64_t a = 0; a < na; a++) {
                        const double *src = out + (a0 + a) * m + c0;
                        for (int64_t c = 0; c < nc; c++) loc[c * TA + a] = src[c];
                    }
                    if (c1 <= a0 && na == TA && (((uintptr_t) (out + c0 * m + a0)) & 15) == 0 && (m & 1) == 0) {
                        for (int64_t c = 0; c < nc; c++) {
                            double *dst = out + (c0 + c) * m + a0;
                            const double *l = loc + c * TA;
                            for (int64_t a = 0; a < TA; a += 2) _mm_stream_pd(dst + a, _mm_load_pd(l + a));
                        }
                    } else {
                        for (int64_t c = 0; c < nc; c++) {
                            double *dst = out + (c0 + c) * m;
                            const int64_t lo = std::max(a0, c0 + c + 1);      // strictly above the diagonal: column a > row c
                            for (int64_t a = lo; a < a1; a++) dst[a] = loc[c * TA + (a - a0)];
                        }
                    }
                }
                _mm_sfence();
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(work);
        for (int64_t k = 0; k < nb; k++) {
            if (k == 0) { if (cudaEventSynchronize(kernels_done) == cudaSuccess) g_phi_timing[2] = now_s(); }
            const cudaError_t e = cudaEventSynchronize(ev[k]);
            if (e != cudaSuccess) { rc = cuda_fail(e, "D2H of the result"); for (int64_t q = k; q < nb; q++) landed[q].store(-1, std::memory_order_release); break; }
            landed[k].store(1, std::memory_order_release);
        }
        for (auto &th : pool) th.join();
    }
    cudaStreamSynchronize(j->stream);
    for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    cudaEventDestroy(kernels_done);
    return rc;
}

}  // namespace

extern "C" {

int gk_phi_f64(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
               double *out, double *mean_or_null, const gk_opts *opts) {
    if (!out) return fail(GK_ERR_ARG, "null output");
    std::lock_guard<std::mutex> call_lock(g_call_mu);         // one call at a time
    int ndev = 1;
    bool virt = false;
    int rc = multi_devices(opts, &ndev, &virt);
    if (rc) return rc;
    const int compress = opts && opts->struct_size >= 16 ? opts->compress : -1;
    rc = phi_pipelined(n, sire, dam, m, pro, out, mean_or_null, opts, ndev, virt, compress);
    if (rc == -3) rc = phi_pipelined(n, sire, dam, m, pro, out, mean_or_null, opts, ndev, virt, 0);   // the same job, one row per individual
    return rc;
}

void gk_phi_last_timing(double *out6) { if (out6) std::memcpy(out6, g_phi_timing, sizeof g_phi_timing); }

// Replaces compute_inbreedings(Pedigree<>&, proband_ids) (include/compute.hpp, lib/compute.cpp:725-814) behind the
// binding of gen.f: out[i] = F of proband i (caller order).  Runs the kinship recurrence without the final expansion.
int gk_f_f64(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro, double *out, const gk_opts *opts) {
    if (!out) return fail(GK_ERR_ARG, "null output");
    std::lock_guard<std::mutex> call_lock(g_call_mu);
    gk_opts o{};
    if (opts) std::memcpy(&o, opts, std::min<size_t>(sizeof(gk_opts), (size_t) std::max(0, opts->struct_size)));
    else { o.device = -1; o.compress = -1; o.order = -1; }
    o.struct_size = sizeof(gk_opts); o.rank = 0; o.world = 1; o.flags |= GK_FLAG_DIAGONAL_ONLY;
    gk_phi_job *j = nullptr;
    int rc = gk_phi_plan(n, sire, dam, m, pro, &o, &j);
    if (rc) return rc;
    rc = gk_phi_upload(j);
    if (!rc) rc = gk_phi_run(j);
    if (!rc) rc = gk_phi_f(j, out);
    std::string keep = g_err;
    destroy(j);
    if (rc) g_err = keep;
    return rc;
}

int gk_gc_f64(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
              int32_t a, const int32_t *anc, double *out, const gk_opts *opts) {
    if (!out) return fail(GK_ERR_ARG, "null output");
    std::lock_guard<std::mutex> call_lock(g_call_mu);         // one call at a time
    const auto t_begin = std::chrono::steady_clock::now();
    gk::GcPlan P;
    if (!gk::build_gc_plan(n, sire, dam, m, pro, a, anc, P))
        return fail(P.error_code == 2 ? GK_ERR_INDEX : GK_ERR_ARG, P.error);
    int dev = 0;
    int rc = select_device(opts, &dev);
    if (rc) return rc;
    cudaStream_t st = (opts && opts->stream) ? (cudaStream_t) opts->stream : nullptr;
    bool own = false;
    if (!st) { GK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); own = true; }
    double *V[2] = { nullptr, nullptr }, *dout = nullptr;
    int32_t *ints = nullptr;
    cudaEvent_t ev[2] = { nullptr, nullptr };
    std::vector<int32_t> arena;
    struct Off { size_t p0, p1, so, sc; };
    std::vector<Off> off(P.G);
    auto push = [&](const std::vector<int32_t> &v) { size_t o = arena.size(); arena.insert(arena.end(), v.begin(), v.end()); return o; };
    int64_t member_rows = 0;
    for (int g = 0; g < P.G; g++) {
        off[g].p0 = push(P.steps[g].p0); off[g].p1 = push(P.steps[g].p1);
        off[g].so = push(P.steps[g].seed_off); off[g].sc = push(P.steps[g].seed_col);
        member_rows += P.steps[g].rows;
    }
    const size_t orow = push(P.row_of);
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        pool_free(V[0]); pool_free(V[1]); pool_free(dout); pool_free(ints);
        if (ev[0]) cudaEventDestroy(ev[0]);
        if (ev[1]) cudaEventDestroy(ev[1]);
        if (own) cudaStreamDestroy(st);
    };
#define GK_CUDA_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = cuda_fail(e_, #call); cleanup(); return c_; } } while (0)
    // The ancestor columns are independent, so a wide ancestor list is propagated block of columns by block
    // of columns: the two live matrices (largest cut x block) plus the m x block result must fit the HBM that is free.
    size_t free_b = 0, total_b = 0;
    GK_CUDA_C(cudaMemGetInfo(&free_b, &total_b));
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (const PoolBlock &b : g_pool_free) if (b.dev == dev) free_b += b.bytes;      // cached blocks are dropped on demand
    }
    const int64_t rows_live = 2 * std::max<int64_t>(P.max_rows, 1) + m;
    int64_t budget = (int64_t) (0.85 * (double) free_b) - (int64_t) arena.size() * 4;
    if (const char *e = std::getenv("GK_GC_BUDGET_MB")) budget = (int64_t) std::atoll(e) << 20;   // tests: force several blocks
    int64_t blk = budget / (rows_live * 8) / 16 * 16;
    if (blk < 16) { cleanup(); return fail(GK_ERR_OOM, "out of device memory: gen.gc needs 16 ancestor columns x (2 x largest cut + probands) doubles of HBM"); }
    blk = std::min<int64_t>(blk, P.ld);
    const size_t vbytes = (size_t) std::max<int64_t>(P.max_rows, 1) * blk * sizeof(double);
    GK_CUDA_C(pool_alloc_t(&V[0], vbytes, dev));
    GK_CUDA_C(pool_alloc_t(&V[1], vbytes, dev));
    GK_CUDA_C(pool_alloc_t(&dout, (size_t) m * blk * sizeof(double), dev));
    GK_CUDA_C(pool_alloc_t(&ints, std::max<size_t>(arena.size(), 1) * sizeof(int32_t), dev));
    GK_CUDA_C(cudaEventCreate(&ev[0]));
    GK_CUDA_C(cudaEventCreate(&ev[1]));
    GK_CUDA_C(cudaMemcpyAsync(ints, arena.data(), arena.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    double kernel_ms = 0.0;
    int64_t nblocks = 0, launches = 0;
    for (int64_t c0 = 0; c0 < a; c0 += blk, nblocks++) {
        const int64_t w = std::min<int64_t>(blk, a - c0);            // ancestor columns [c0, c0 + w) of this block
        const int64_t ldb = (w + 15) / 16 * 16;
        GK_CUDA_C(cudaEventRecord(ev[0], st));
        for (int g = 0; g < P.G; g++) {
            gk::GcStepDev d{};
            d.Vprev = V[(g + 1) & 1]; d.Vnew = V[g & 1]; d.ld = ldb; d.rows = P.steps[g].rows;
            d.p0 = ints + off[g].p0; d.p1 = ints + off[g].p1; d.seed_off = ints + off[g].so; d.seed_col = ints + off[g].sc;
            d.col0 = (int32_t) c0; d.ncols = (int32_t) w;
            GK_CUDA_C(gk::launch_gc_step(d, st));
        }
        GK_CUDA_C(gk::launch_gc_gather(V[(P.G - 1) & 1], ldb, ints + orow, dout, m, w, st));
        GK_CUDA_C(cudaEventRecord(ev[1], st));
        launches += P.G + 1;
        GK_CUDA_C(cudaMemcpy2DAsync(out + c0, (size_t) a * sizeof(double), dout, (size_t) w * sizeof(double), (size_t) w * sizeof(double),
                                    (size_t) m, cudaMemcpyDeviceToHost, st));
        GK_CUDA_C(cudaStreamSynchronize(st));
        float ms = 0.f;
        GK_CUDA_C(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        kernel_ms += ms;
    }
    g_gc_stats.kernel_ms = kernel_ms;
    g_gc_stats.algorithmic_bytes = 24 * member_rows * (int64_t) a;          // 2 parent rows read + 1 row written per member, per ancestor
    g_gc_stats.member_rows = member_rows;
    g_gc_stats.column_blocks = nblocks;
    g_gc_stats.block_columns = blk;
    g_gc_stats.kernel_launches = launches;
    g_gc_stats.n_cuts = P.G;
    cleanup();
#undef GK_CUDA_C
    g_gc_stats.total_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    return GK_OK;
}

int gk_gc_last_stats(gk_gc_stats *s) {
    if (!s) return fail(GK_ERR_ARG, "null argument");
    *s = g_gc_stats;
    return GK_OK;
}

}  // extern "C"
