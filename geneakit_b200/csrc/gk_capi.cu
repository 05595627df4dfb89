// C ABI of libgeneakit_b200.so (see include/geneakit_b200.h).  Host orchestration only:
// plan on the host (gk_plan.hpp), keep matrices in HBM, enqueue the sm_100a kernels
// (gk_kernels.cu) on one stream.  There is no CPU compute path in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/geneakit_b200.h"
#include "gk_gcplan.hpp"
#include "gk_kernels.cuh"
#include "gk_plan.hpp"

namespace {

thread_local std::string g_err;

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

struct StepOffsets {
    size_t prow = 0, pcnt = 0, slots = 0, prow8 = 0, pcnt8 = 0, slots8 = 0, first_use = 0;
    size_t pF = 0, pM = 0, flags = 0, pg_i = 0, pg_j = 0, pg_off = 0, pt_cls = 0, pt_mult = 0, grp_off = 0, grp_info = 0;
    size_t done = 0;              // offset into the counter arena: done1 at +0, done2 at +np
    int32_t ngroups = 0, npatch = 0;
    int64_t total_items = 0;
    int32_t pb = 0, pe = 0, cpp = 0;      // own panels [pb, pe), panels per rank
    size_t grp_off2 = 0; int64_t total1 = 0, total2 = 0; int32_t n1pp = 0, ring_tma = 4;   // multi-GPU kernel: two item streams
    int32_t lag = 1, ring_n = 3;          // phase 2 trails phase 1 by `lag` panels; panel buffers in the ring
};

}  // namespace

struct gk_phi_job {
    gk::PhiPlan plan;
    gk_opts opts{};
    int device = 0;
    int rank = 0, world = 1;
    bool uploaded = false, ran = false;
    int steps_done = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double *buf[2] = { nullptr, nullptr };            // own rows of the class matrices, ping-pong by step parity
    double *peer[gk::kMaxShards][2] = {};             // every rank's buffers (own entry = buf); IPC mappings for the others
    bool peer_ipc[gk::kMaxShards][2] = {};
    bool peers_set = false;
    double *dvec[2] = { nullptr, nullptr };           // self kinship per class (full length, replicated)
    int64_t dvec_len = 0;
    double *ring = nullptr;
    double *zero_row = nullptr;
    int use_tma = 0;                  // step kernel: 0 = warp-autonomous LDGSTS (one GPU), 1 = TMA phase 1 (multi-GPU; GK_STEP_KERNEL=tma|ldgsts)
    int full_rows = 0;                // phase 2 evaluates every tile J and stores only the own rows (multi-GPU; GK_FULL_ROWS=1 forces it)
    int lag_env = 0, ring_env = 0, lag_scale_pct = 100;   // tuning overrides (GK_LAG, GK_RING, GK_LAG_PCT)
    double *out = nullptr;            // result rows; single GPU: aliases the buffer the last step does not use
    double *out_own = nullptr;        // separately allocated result (world > 1)
    double *gloc = nullptr;           // world > 1: class rows needed by the expansion, fetched from their owners
    double *partials = nullptr;
    double *sums_dev = nullptr;
    double *sums_host = nullptr;      // pinned
    int32_t *ints = nullptr;          // all plan arrays, one allocation
    int32_t *counters = nullptr;
    size_t counters_count = 0;
    std::vector<StepOffsets> off;
    size_t off_final_cls = 0, off_final_ind = 0, off_cact = 0, off_coff = 0, off_cmem = 0, off_flags0 = 0, ints_count = 0;
    int64_t row0 = 0, row1 = 0;       // result rows owned by this shard (caller order)
    int32_t nact = 0, nchunks = 0;
    int64_t nblocks_expand = 0;
    double *scratch = nullptr;
    std::vector<cudaEvent_t> ev;      // start / end events of every step (2g, 2g+1); 0, 1: the final expansion
    int64_t device_bytes = 0;
    int64_t launches = 0;
    int64_t row_bytes = 0, remote_row_bytes = 0;   // bytes of parent rows this rank's phase 1 reads / reads from peers
};

namespace {

void destroy(gk_phi_job *j) {
    if (!j) return;
    cudaSetDevice(j->device);
    for (int q = 0; q < gk::kMaxShards; q++)
        for (int p = 0; p < 2; p++)
            if (j->peer_ipc[q][p] && j->peer[q][p]) cudaIpcCloseMemHandle(j->peer[q][p]);
    for (int p = 0; p < 2; p++) { if (j->buf[p]) cudaFree(j->buf[p]); if (j->dvec[p]) cudaFree(j->dvec[p]); }
    if (j->ring) cudaFree(j->ring);
    if (j->zero_row) cudaFree(j->zero_row);
    if (j->out_own) cudaFree(j->out_own);
    if (j->gloc) cudaFree(j->gloc);
    if (j->partials) cudaFree(j->partials);
    if (j->scratch) cudaFree(j->scratch);
    if (j->sums_dev) cudaFree(j->sums_dev);
    if (j->sums_host) cudaFreeHost(j->sums_host);
    if (j->ints) cudaFree(j->ints);
    if (j->counters) cudaFree(j->counters);
    for (cudaEvent_t e : j->ev) cudaEventDestroy(e);
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

std::once_flag g_cfg_once;
cudaError_t g_cfg_err = cudaSuccess;

// panels of step g owned by each rank: equal blocks of cpp = ceil(np / world) panels
void panel_range(int np, int rank, int world, int32_t *pb, int32_t *pe, int32_t *cpp) {
    const int c = (np + world - 1) / world;
    *cpp = c;
    *pb = std::min(np, rank * c);
    *pe = std::min(np, (rank + 1) * c);
}

gk::RowTable row_table(const gk_phi_job *j, int g) {
    gk::RowTable t{};
    const gk::StepPlan &S = j->plan.steps[g];
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

}  // namespace

extern "C" {

const char *gk_last_error(void) { return g_err.c_str(); }
int32_t gk_abi_version(void) { return GK_ABI_VERSION; }

void *gk_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        g_err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
void gk_host_free(void *p) { if (p) cudaFreeHost(p); }

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
int gk_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
                const gk_opts *opts, gk_phi_job **job) {
    if (!job) return fail(GK_ERR_ARG, "job pointer is null");
    *job = nullptr;
    gk_phi_job *j = new gk_phi_job();
    if (opts) std::memcpy(&j->opts, opts, std::min<size_t>(sizeof(gk_opts), (size_t) std::max(0, opts->struct_size)));
    else { j->opts.device = -1; j->opts.compress = -1; j->opts.order = -1; }
    if (j->opts.world < 1) j->opts.world = 1;
    if (j->opts.world > gk::kMaxShards) { delete j; return fail(GK_ERR_ARG, "world > 8 shards"); }
    if (j->opts.rank < 0 || j->opts.rank >= j->opts.world) { delete j; return fail(GK_ERR_ARG, "rank outside [0, world)"); }
    j->rank = j->opts.rank; j->world = j->opts.world;
    if (!gk::build_phi_plan(n, sire, dam, m, pro, j->opts.compress, j->opts.order, j->world, j->plan)) {
        int code = j->plan.error_code == 2 ? GK_ERR_INDEX : GK_ERR_ARG;
        std::string msg = j->plan.error;
        delete j;
        return fail(code, msg);
    }
    // a rank of a multi-GPU job evaluates every tile J and stores only its own rows; RANKED jobs (rounding order
    // fixed by rank) must keep J <= I and store the mirror tile into the rows of J's owner instead
    j->full_rows = (j->world > 1 && !j->plan.ranked) ? 1 : 0;
    j->use_tma = j->world > 1 ? 1 : 0;
    if (const char *e = std::getenv("GK_STEP_KERNEL")) j->use_tma = (e[0] == 't') ? 1 : 0;
    if (j->world > 1 && j->plan.ranked) j->use_tma = 1;      // only that kernel stores mirror tiles into a peer's rows
    if (const char *e = std::getenv("GK_FULL_ROWS")) if (e[0] == '1' && !j->plan.ranked) j->full_rows = 1;
    if (const char *e = std::getenv("GK_LAG")) j->lag_env = std::atoi(e);
    if (const char *e = std::getenv("GK_RING")) j->ring_env = std::atoi(e);
    if (const char *e = std::getenv("GK_LAG_PCT")) j->lag_scale_pct = std::max(1, std::atoi(e));
    const int G = j->plan.G;
    j->off.assign(G, StepOffsets{});
    for (int g = 0; g < G; g++)
        panel_range(j->plan.steps[g].np, j->rank, j->world, &j->off[g].pb, &j->off[g].pe, &j->off[g].cpp);
    gk_shard_rows(j->plan.m, j->rank, j->world, &j->row0, &j->row1);
    // rows of the previous matrix this rank's panels read, and how many of them live on a peer (NVLink)
    for (int g = 1; g < G; g++) {
        const gk::StepPlan &S = j->plan.steps[g];
        const StepOffsets &o = j->off[g];
        const int64_t rps = (int64_t) j->off[g - 1].cpp * gk::kPanel;
        std::vector<int32_t> stamp((size_t) std::max<int64_t>(S.K_prev, 1), -1);
        const int unit = j->use_tma ? gk::kSub : gk::kPanel;          // classes whose parent rows are staged together
        for (int p = o.pb * gk::kPanel / unit; p < o.pe * gk::kPanel / unit; p++)
            for (int i = p * unit; i < (p + 1) * unit; i++)
                for (int q = 0; q < 2; q++) {
                    const int32_t r = q ? S.pM[i] : S.pF[i];
                    if (r < 0 || stamp[r] == p) continue;
                    stamp[r] = p;
                    j->row_bytes += 8 * S.Kpad_prev;
                    if (r / rps != j->rank) j->remote_row_bytes += 8 * S.Kpad_prev;
                }
    }
    *job = j;
    return GK_OK;
}

int gk_phi_upload(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (j->uploaded) return GK_OK;
    int rc = select_device(&j->opts, &j->device);
    if (rc) return rc;
    std::call_once(g_cfg_once, [] { g_cfg_err = gk::configure_kernels(); });
    if (g_cfg_err != cudaSuccess) return cuda_fail(g_cfg_err, "kernel configuration (is this an sm_100 device?)");
    gk::PhiPlan &P = j->plan;
    const int G = P.G;
    if (j->opts.stream) j->stream = (cudaStream_t) j->opts.stream;
    else { GK_CUDA(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking)); j->own_stream = true; }

    // ---- plan arrays: one int32 arena (every array 16-byte aligned) -------------------------
    std::vector<int32_t> arena;
    auto push = [&](const std::vector<int32_t> &v) {
        size_t o = arena.size();
        arena.insert(arena.end(), v.begin(), v.end());
        while (arena.size() % 4) arena.push_back(0);
        return o;
    };
    size_t ncount = 0;
    j->off_flags0 = push(P.steps[0].flags);
    for (int g = 1; g < G; g++) {
        gk::StepPlan &S = P.steps[g];
        StepOffsets &o = j->off[g];
        o.pF = push(S.pF); o.pM = push(S.pM); o.flags = push(S.flags);
        o.prow = push(S.prow); o.pcnt = push(S.pcnt); o.slots = push(S.slots);
        o.prow8 = push(S.prow8); o.pcnt8 = push(S.pcnt8); o.slots8 = push(S.slots8);
        o.first_use = push(S.first_use);
        o.pg_i = push(S.pg_i); o.pg_j = push(S.pg_j); o.pg_off = push(S.pg_off);
        o.pt_cls = push(S.pt_cls); o.pt_mult = push(S.pt_mult);
        o.npatch = (int32_t) S.pg_i.size();
        // work queue of the own panels: phase 2 trails phase 1 by `lag` panels: P1(s), P2(s - lag), ...
        const int ncb = (int) ((S.Kpad_prev + gk::kCB - 1) / gk::kCB);
        std::vector<int64_t> goff(1, 0);
        std::vector<int32_t> ginfo;
        auto add = [&](int phase, int panel) {
            const int64_t items = phase == 1 ? ncb : (j->full_rows ? S.np : panel + 1);
            ginfo.push_back(panel | ((phase - 1) << 30));
            goff.push_back(goff.back() + items);
        };
        // The resident warps work on a window of ~nwarps consecutive queue items: phase 2 of a panel
        // should start only after that many items have been dealt since its phase 1 ended.
        {
            const int64_t nwarps = (int64_t) gk::step_grid() * gk::kStepWarps;
            int lag = (int) ((nwarps * j->lag_scale_pct / 100 + ncb - 1) / ncb);
            lag = std::max(1, std::min(lag, (int) gk::kRingMax - 2));
            if (j->lag_env > 0) lag = std::min(j->lag_env, (int) gk::kRingMax - 1);
            o.lag = lag;
            o.ring_n = j->ring_env > 0 ? std::max(lag + 1, std::min(j->ring_env, (int) gk::kRingMax)) : lag + 2;
        }
        for (int s = o.pb; s < o.pe + o.lag; s++) {
            if (s < o.pe) add(1, s);
            if (s - o.lag >= o.pb) add(2, s - o.lag);
        }
        o.ngroups = (int32_t) ginfo.size();
        o.total_items = goff.back();
        std::vector<int32_t> g32(goff.size() * 2);
        std::memcpy(g32.data(), goff.data(), goff.size() * 8);
        o.grp_off = push(g32); o.grp_info = push(ginfo);
        {   // multi-GPU kernel (phi_step_tma_kernel): phase-1 items = 8 classes x 256 columns, phase-2 items = one tile
            const int ncb256 = (int) ((S.Kpad_prev + gk::kP1Cols - 1) / gk::kP1Cols);
            o.n1pp = (gk::kPanel / gk::kP1Classes) * ncb256;
            o.total1 = (int64_t) (o.pe - o.pb) * o.n1pp;
            std::vector<int64_t> goff2(1, 0);
            for (int s = o.pb; s < o.pe; s++) goff2.push_back(goff2.back() + (j->full_rows ? S.np : s + 1));
            o.total2 = goff2.back();
            std::vector<int32_t> g32b(goff2.size() * 2);
            std::memcpy(g32b.data(), goff2.data(), goff2.size() * 8);
            o.grp_off2 = push(g32b);
            o.ring_tma = j->ring_env > 0 ? std::max(2, std::min(j->ring_env, (int) gk::kRingMax)) : 4;
            if (j->use_tma) o.ring_n = o.ring_tma;
        }
        o.done = ncount; ncount += 2 * (size_t) S.np;
    }
    j->off_final_cls = push(P.final_cls);
    j->off_final_ind = push(P.final_ind);
    // result rows of this shard and, for the expansion kernel, its caller entries grouped by class
    {
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
        j->nchunks = (int32_t) ((P.m + gk::kExpandCols - 1) / gk::kExpandCols);
        j->off_cact = push(cact); j->off_coff = push(coff); j->off_cmem = push(cmem);
        j->nblocks_expand = P.single_cut ? 1 : (int64_t) j->nact * j->nchunks;
    }
    j->ints_count = arena.size();
    GK_CUDA(cudaMalloc(&j->ints, std::max<size_t>(arena.size(), 4) * sizeof(int32_t)));
    GK_CUDA(cudaMemcpyAsync(j->ints, arena.data(), arena.size() * sizeof(int32_t), cudaMemcpyHostToDevice, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    j->device_bytes += (int64_t) arena.size() * 4;
    j->counters_count = std::max<size_t>(ncount, 1);
    GK_CUDA(cudaMalloc(&j->counters, j->counters_count * sizeof(int32_t)));

    // ---- matrices: two ping-pong buffers holding the own rows of each cut's class matrix --------
    int64_t need[2] = { 0, 0 };
    int64_t max_prev = 1, dlen = 1;
    for (int g = 0; g < G; g++) {
        const gk::StepPlan &S = P.steps[g];
        const int64_t rows = (int64_t) (j->off[g].pe - j->off[g].pb) * gk::kPanel;
        need[g & 1] = std::max(need[g & 1], rows * S.Kpad);
        if (g + 1 < G) max_prev = std::max(max_prev, S.Kpad);
        dlen = std::max(dlen, (int64_t) j->off[g].cpp * gk::kPanel * j->world);
    }
    const int64_t mm = (j->row1 - j->row0) * (int64_t) P.m;
    const int out_buf = P.single_cut ? 0 : 1 - ((G - 1) & 1);
    if (j->world == 1) need[out_buf] = std::max(need[out_buf], mm);
    for (int p = 0; p < 2; p++) {
        need[p] = std::max<int64_t>(need[p], 32);
        cudaError_t e = cudaMalloc(&j->buf[p], (size_t) need[p] * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            char msg[256];
            std::snprintf(msg, sizeof msg, "out of device memory: kinship buffers need %.2f GB + %.2f GB of HBM",
                          need[0] * 8 / 1e9, need[1] * 8 / 1e9);
            return fail(GK_ERR_OOM, msg);
        }
        j->device_bytes += need[p] * 8;
        j->peer[j->rank][p] = j->buf[p];
    }
    if (j->world == 1) { j->out = j->buf[out_buf]; j->peers_set = true; }
    else {
        GK_CUDA(cudaMalloc(&j->out_own, (size_t) std::max<int64_t>(mm, 1) * sizeof(double)));
        j->out = j->out_own;
        const int64_t gl = (int64_t) std::max(j->nact, 1) * P.steps[G - 1].Kpad;
        GK_CUDA(cudaMalloc(&j->gloc, (size_t) gl * sizeof(double)));
        j->device_bytes += (mm + gl) * 8;
    }
    j->dvec_len = dlen;
    for (int p = 0; p < 2; p++) { GK_CUDA(cudaMalloc(&j->dvec[p], (size_t) dlen * sizeof(double))); j->device_bytes += dlen * 8; }
    int64_t ring_doubles = 32;
    for (int g = 1; g < G; g++) ring_doubles = std::max(ring_doubles, (int64_t) j->off[g].ring_n * P.steps[g].Kpad_prev * gk::kPanel);
    GK_CUDA(cudaMalloc(&j->ring, (size_t) ring_doubles * sizeof(double)));
    GK_CUDA(cudaMalloc(&j->zero_row, gk::kP1Cols * sizeof(double)));
    GK_CUDA(cudaMemsetAsync(j->zero_row, 0, gk::kP1Cols * sizeof(double), j->stream));
    j->device_bytes += ring_doubles * 8;
    GK_CUDA(cudaMalloc(&j->partials, (size_t) std::max<int64_t>(j->nblocks_expand, 1) * 2 * sizeof(double)));
    GK_CUDA(cudaMalloc(&j->scratch, 2 * 256 * sizeof(double)));
    j->device_bytes += j->nblocks_expand * 16;
    GK_CUDA(cudaMalloc(&j->sums_dev, 2 * sizeof(double)));
    GK_CUDA(cudaHostAlloc(&j->sums_host, 2 * sizeof(double), cudaHostAllocDefault));
    j->ev.resize(2 * (G + 1));
    for (auto &e : j->ev) GK_CUDA(cudaEventCreate(&e));
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
    if (!j || !j->uploaded || step < 0 || step >= j->plan.G) return fail(GK_ERR_ARG, "bad step");
    if (dev) *dev = j->dvec[step & 1];
    if (chunk_doubles) *chunk_doubles = (int64_t) j->off[step].cpp * gk::kPanel;
    if (total_doubles) *total_doubles = (int64_t) j->off[step].cpp * gk::kPanel * j->world;
    return GK_OK;
}

// Same-process emulation of the all-gather (several ranks' jobs on ONE GPU, one stream): copy the
// chunk rank `src` owns of the self-kinship vector of cut `step` into this job's replica.
int gk_phi_dvec_copy_from(gk_phi_job *j, gk_phi_job *src, int32_t step) {
    if (!j || !src || !j->uploaded || !src->uploaded || step < 0 || step >= j->plan.G) return fail(GK_ERR_ARG, "bad argument");
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
    gk::PhiPlan &P = j->plan;
    if (g < 0 || g >= P.G) return fail(GK_ERR_ARG, "bad step");
    GK_CUDA(cudaSetDevice(j->device));
    cudaStream_t st = j->stream;
    const gk::StepPlan &S = P.steps[g];
    const StepOffsets &o = j->off[g];
    if (g == 0) {
        GK_CUDA(cudaMemsetAsync(j->counters, 0, j->counters_count * sizeof(int32_t), st));
        j->launches = 0;
        if (P.single_cut) { j->steps_done = 1; return GK_OK; }
        // own rows of the founders' matrix
        const int64_t rows = (int64_t) (o.pe - o.pb) * gk::kPanel;
        if (j->world == 1) {
            GK_CUDA(gk::launch_init(j->buf[0], j->dvec[0], j->ints + j->off_flags0, S.K, S.Kpad, st));
        } else {
            // (rare) uncompressed multi-GPU start: every rank zeroes its rows, the diagonal comes from the flags
            GK_CUDA(cudaMemsetAsync(j->buf[0], 0, (size_t) std::max<int64_t>(rows * S.Kpad, 1) * sizeof(double), st));
            GK_CUDA(gk::launch_init_rows(j->buf[0], j->dvec[0], j->ints + j->off_flags0, S.K, S.Kpad,
                                         (int64_t) o.pb * gk::kPanel, rows, j->dvec_len, st));
        }
        j->launches++;
        j->steps_done = 1;
        return GK_OK;
    }
    if (j->opts.verbose && j->rank == 0) std::printf("Cut %d out of %d\n", g, P.G);
    gk::StepDev d{};
    d.prev = row_table(j, g - 1);
    d.outt = row_table(j, g);
    d.out_base = j->buf[g & 1]; d.out_row0 = (int64_t) o.pb * gk::kPanel;
    d.ld_prev = S.Kpad_prev; d.ld_new = S.Kpad;
    d.Kpad_prev = (int32_t) S.Kpad_prev; d.np = S.np;
    d.pF = j->ints + o.pF; d.pM = j->ints + o.pM;
    d.prow = j->ints + o.prow; d.pcnt = j->ints + o.pcnt; d.slots = j->ints + o.slots;
    d.first_use = j->ints + o.first_use;
    d.prow8 = j->ints + o.prow8; d.pcnt8 = j->ints + o.pcnt8; d.slots8 = j->ints + o.slots8;
    d.zero_row = j->zero_row;
    d.ring = j->ring; d.ring_stride = S.Kpad_prev * gk::kPanel; d.ring_n = o.ring_n;
    d.done1 = j->counters + o.done; d.done2 = d.done1 + S.np;
    d.grp_off = reinterpret_cast<const int64_t *>(j->ints + o.grp_off); d.grp_info = j->ints + o.grp_info;
    d.total_items = o.total_items; d.ngroups = o.ngroups;
    d.ncb = (int32_t) ((S.Kpad_prev + gk::kCB - 1) / gk::kCB);
    d.panel_begin = o.pb; d.panel_end = o.pe;
    d.full_rows = j->full_rows;
    { const char *e = std::getenv("GK_OUT_POLICY"); d.out_policy = e ? std::atoi(e) : 0; }
    { const char *e = std::getenv("GK_DEBUG_NOWAIT"); d.debug_nowait = (e && e[0] == '1') ? 1 : 0; }
    GK_CUDA(cudaEventRecord(j->ev[2 * g], st));
    d.grp_off2 = reinterpret_cast<const int64_t *>(j->ints + o.grp_off2);
    d.total1 = o.total1; d.total2 = o.total2; d.n1pp = o.n1pp;
    d.ncb256 = (int32_t) ((S.Kpad_prev + gk::kP1Cols - 1) / gk::kP1Cols);
    if (j->use_tma) GK_CUDA(gk::launch_step_tma(d, st));
    else GK_CUDA(gk::launch_step(d, st));
    gk::DiagDev q{};
    q.prev = d.prev; q.out = row_table(j, g); q.ld_prev = d.ld_prev; q.ld_new = d.ld_new;
    q.dprev = j->dvec[(g - 1) & 1]; q.dnew = j->dvec[g & 1];
    q.pF = d.pF; q.pM = d.pM; q.flags = j->ints + o.flags;
    q.c_begin = (int32_t) std::min<int64_t>(S.K, (int64_t) o.pb * gk::kPanel);
    q.c_end = (int32_t) std::min<int64_t>(S.K, (int64_t) o.pe * gk::kPanel);
    q.pg_i = j->ints + o.pg_i; q.pg_j = j->ints + o.pg_j; q.pg_off = j->ints + o.pg_off;
    q.pt_cls = j->ints + o.pt_cls; q.pt_mult = j->ints + o.pt_mult; q.ngroups = o.npatch;
    GK_CUDA(gk::launch_diag(q, st));
    GK_CUDA(cudaEventRecord(j->ev[2 * g + 1], st));
    j->launches += 2;
    j->steps_done = g + 1;
    return GK_OK;
}

// Final expansion to the caller's proband order + the fused phiMean partial sums of the own rows.
int gk_phi_finish(gk_phi_job *j) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    gk::PhiPlan &P = j->plan;
    const int G = P.G;
    if (j->steps_done < G) return fail(GK_ERR_STATE, "gk_phi_finish before the last gk_phi_run_step");
    GK_CUDA(cudaSetDevice(j->device));
    cudaStream_t st = j->stream;
    const int last = G - 1;
    GK_CUDA(cudaEventRecord(j->ev[0], st));
    if (P.single_cut) {
        GK_CUDA(gk::launch_eye(j->out, P.m, j->row0, j->row1, j->partials, st));
    } else {
        gk::ExpandDev e{};
        e.d = j->dvec[last & 1]; e.ld = P.steps[last].Kpad;
        e.cls = j->ints + j->off_final_cls; e.ind = j->ints + j->off_final_ind;
        e.cact = j->ints + j->off_cact; e.coff = j->ints + j->off_coff; e.cmem = j->ints + j->off_cmem;
        e.out = j->out; e.m = P.m; e.row0 = j->row0; e.row1 = j->row1;
        e.nact = j->nact; e.nchunks = j->nchunks; e.partials = j->partials;
        if (j->world == 1) { e.G = j->buf[last & 1]; e.compact = 0; }
        else {
            // the class rows of the own probands live on their owners: fetch them (coalesced peer reads)
            GK_CUDA(gk::launch_fetch_rows(row_table(j, last), e.ld, e.cact, j->nact, j->gloc, st));
            j->launches++;
            e.G = j->gloc; e.compact = 1;
        }
        GK_CUDA(gk::launch_expand(e, st));
    }
    GK_CUDA(gk::launch_mean_finish(j->partials, j->nblocks_expand, j->scratch, j->sums_dev, st));
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
    gk::PhiPlan &P = j->plan;
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
    int rc = gk_phi_sync(j);
    if (rc) return rc;
    const double total = j->sums_host[0], diag = j->sums_host[1];
    if (sum_all) *sum_all = total;
    if (sum_diag) *sum_diag = diag;
    const double n = (double) j->plan.m;
    // geneakit/compute.py:131-132 (with world > 1 these are this shard's partial sums: the driver
    // all-reduces sum_all / sum_diag across ranks and applies the same formula)
    if (mean) *mean = (total - diag) / (n * n - n);
    return GK_OK;
}

int gk_phi_copy_out(gk_phi_job *j, double *host_out) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_copy_out before gk_phi_run");
    if (!host_out) return fail(GK_ERR_ARG, "null output");
    GK_CUDA(cudaSetDevice(j->device));
    const size_t bytes = (size_t) (j->row1 - j->row0) * j->plan.m * sizeof(double);
    GK_CUDA(cudaMemcpyAsync(host_out, j->out, bytes, cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    return GK_OK;
}

int gk_phi_copy_rows(gk_phi_job *j, int64_t row_begin, int64_t nrows, double *host_out) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "gk_phi_copy_rows before gk_phi_run");
    if (!host_out || row_begin < j->row0 || nrows < 0 || row_begin + nrows > j->row1) return fail(GK_ERR_ARG, "rows outside this job's block");
    GK_CUDA(cudaSetDevice(j->device));
    GK_CUDA(cudaMemcpyAsync(host_out, j->out + (row_begin - j->row0) * (int64_t) j->plan.m, (size_t) nrows * j->plan.m * sizeof(double),
                            cudaMemcpyDeviceToHost, j->stream));
    GK_CUDA(cudaStreamSynchronize(j->stream));
    return GK_OK;
}

int gk_phi_device_result(gk_phi_job *j, double **dev, int64_t *ld) {
    if (!j || !j->uploaded) return fail(GK_ERR_STATE, "job not uploaded");
    if (dev) *dev = j->out;
    if (ld) *ld = j->plan.m;
    return GK_OK;
}

int gk_phi_get_stats(gk_phi_job *j, gk_phi_stats *s) {
    if (!j || !s) return fail(GK_ERR_ARG, "null argument");
    const gk::PhiPlan &P = j->plan;
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
    s->plan_bytes = (int64_t) j->ints_count * 4;
    s->rank_row_bytes = j->row_bytes; s->rank_remote_row_bytes = j->remote_row_bytes;
    return GK_OK;
}

int gk_phi_cut_sizes(gk_phi_job *j, int32_t *sizes, int32_t cap) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    for (int g = 0; g < j->plan.G && g < cap; g++) sizes[g] = (int32_t) j->plan.steps[g].cut_size;
    return GK_OK;
}
int gk_phi_class_sizes(gk_phi_job *j, int32_t *sizes, int32_t cap) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    for (int g = 0; g < j->plan.G && g < cap; g++) sizes[g] = (int32_t) j->plan.steps[g].K;
    return GK_OK;
}

int gk_phi_step_times(gk_phi_job *j, float *ms, int32_t cap) {
    if (!j || !j->ran) return fail(GK_ERR_STATE, "no run recorded");
    int rc = gk_phi_sync(j);
    if (rc) return rc;
    const int G = j->plan.G;
    if (j->plan.single_cut) return GK_OK;
    for (int g = 1; g < G && g - 1 < cap; g++) GK_CUDA(cudaEventElapsedTime(&ms[g - 1], j->ev[2 * g], j->ev[2 * g + 1]));
    if (G - 1 < cap) GK_CUDA(cudaEventElapsedTime(&ms[G - 1], j->ev[0], j->ev[1]));   // expansion + mean
    return GK_OK;
}

int gk_phi_step_bytes(gk_phi_job *j, int64_t *bytes, int32_t cap) {
    if (!j || !bytes) return fail(GK_ERR_ARG, "null argument");
    const gk::PhiPlan &P = j->plan;
    for (int g = 1; g < P.G && g - 1 < cap; g++) {
        const gk::StepPlan &S = P.steps[g], &Sp = P.steps[g - 1];
        bytes[g - 1] = 8 * (S.P_cls * Sp.K + S.K * S.K - S.kept_cls * S.kept_cls);
    }
    const int64_t mm = (int64_t) P.m * P.m, Kl = P.steps[P.G - 1].K;
    if (P.G - 1 < cap) bytes[P.G - 1] = 8 * (mm + std::min(mm, Kl * Kl));
    return GK_OK;
}

int gk_phi_step_view(gk_phi_job *j, int32_t step, gk_step_view *v) {
    if (!j || !v || step < 0 || step >= j->plan.G) return fail(GK_ERR_ARG, "bad step");
    const gk::PhiPlan &P = j->plan;
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
    if (cls) *cls = j->plan.final_cls.data();
    if (ind) *ind = j->plan.final_ind.data();
    if (K_last) *K_last = j->plan.steps[j->plan.G - 1].K;
    return GK_OK;
}

void gk_phi_free(gk_phi_job *j) { destroy(j); }

int gk_phi_f64(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
               double *out, double *mean_or_null, const gk_opts *opts) {
    if (!out) return fail(GK_ERR_ARG, "null output");
    gk_phi_job *j = nullptr;
    int rc = gk_phi_plan(n, sire, dam, m, pro, opts, &j);
    if (rc) return rc;
    rc = gk_phi_upload(j);
    if (!rc) rc = gk_phi_run(j);
    if (!rc) rc = gk_phi_copy_out(j, out);
    if (!rc && mean_or_null) rc = gk_phi_mean(j, nullptr, nullptr, mean_or_null);
    std::string keep = g_err;
    destroy(j);
    if (rc) g_err = keep;
    return rc;
}

int gk_gc_f64(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
              int32_t a, const int32_t *anc, double *out, const gk_opts *opts) {
    if (!out) return fail(GK_ERR_ARG, "null output");
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
    std::vector<int32_t> arena;
    struct Off { size_t p0, p1, so, sc; };
    std::vector<Off> off(P.G);
    auto push = [&](const std::vector<int32_t> &v) { size_t o = arena.size(); arena.insert(arena.end(), v.begin(), v.end()); return o; };
    for (int g = 0; g < P.G; g++) {
        off[g].p0 = push(P.steps[g].p0); off[g].p1 = push(P.steps[g].p1);
        off[g].so = push(P.steps[g].seed_off); off[g].sc = push(P.steps[g].seed_col);
    }
    const size_t orow = push(P.row_of);
    auto cleanup = [&]() {
        if (V[0]) cudaFree(V[0]); if (V[1]) cudaFree(V[1]); if (dout) cudaFree(dout); if (ints) cudaFree(ints);
        if (own) cudaStreamDestroy(st);
    };
#define GK_CUDA_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int c_ = cuda_fail(e_, #call); cleanup(); return c_; } } while (0)
    const size_t vbytes = (size_t) std::max<int64_t>(P.max_rows, 1) * P.ld * sizeof(double);
    GK_CUDA_C(cudaMalloc(&V[0], vbytes));
    GK_CUDA_C(cudaMalloc(&V[1], vbytes));
    GK_CUDA_C(cudaMalloc(&dout, (size_t) m * a * sizeof(double)));
    GK_CUDA_C(cudaMalloc(&ints, std::max<size_t>(arena.size(), 1) * sizeof(int32_t)));
    GK_CUDA_C(cudaMemcpyAsync(ints, arena.data(), arena.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    for (int g = 0; g < P.G; g++) {
        gk::GcStepDev d{};
        d.Vprev = V[(g + 1) & 1]; d.Vnew = V[g & 1]; d.ld = P.ld; d.rows = P.steps[g].rows;
        d.p0 = ints + off[g].p0; d.p1 = ints + off[g].p1; d.seed_off = ints + off[g].so; d.seed_col = ints + off[g].sc;
        GK_CUDA_C(gk::launch_gc_step(d, st));
    }
    GK_CUDA_C(gk::launch_gc_gather(V[(P.G - 1) & 1], P.ld, ints + orow, dout, m, a, st));
    GK_CUDA_C(cudaMemcpyAsync(out, dout, (size_t) m * a * sizeof(double), cudaMemcpyDeviceToHost, st));
    GK_CUDA_C(cudaStreamSynchronize(st));
    cleanup();
#undef GK_CUDA_C
    return GK_OK;
}

}  // extern "C"
