// C ABI of libgeneakit_b200.so (see include/geneakit_b200.h).  Host orchestration only:
// plan on the host (gk_plan.hpp), keep matrices in HBM, enqueue the sm_100a kernels
// (gk_kernels.cu) on one stream.  There is no CPU compute path in this file.
#include <cstdio>
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

constexpr int64_t kPadDoubles = 256;   // slack after every matrix: window reads may run past the last row

struct StepOffsets { size_t desc; };

}  // namespace

struct gk_phi_job {
    gk::PhiPlan plan;
    gk_opts opts{};
    int device = 0;
    bool uploaded = false, ran = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double *buf[2] = { nullptr, nullptr };
    int64_t buf_doubles[2] = { 0, 0 };
    double *dvec[2] = { nullptr, nullptr };
    double *out = nullptr;            // aliases buf[1 - parity(last)] (or buf[0] when single cut)
    double *partials = nullptr;
    double *sums_dev = nullptr;
    double *sums_host = nullptr;      // pinned
    int32_t *ints = nullptr;          // all plan arrays, one allocation
    std::vector<StepOffsets> off;
    size_t off_final_cls = 0, off_final_ind = 0, off_cact = 0, off_coff = 0, off_cmem = 0, ints_count = 0;
    int64_t row0 = 0, row1 = 0;       // result rows owned by this shard (caller order)
    int32_t nact = 0, nchunks = 0;
    int64_t nblocks_expand = 0;
    double *scratch = nullptr;
    std::vector<cudaEvent_t> ev;      // G + 2 events around the step launches
    int64_t device_bytes = 0;
    int64_t launches = 0;
};

namespace {

void destroy(gk_phi_job *j) {
    if (!j) return;
    cudaSetDevice(j->device);
    for (int p = 0; p < 2; p++) { if (j->buf[p]) cudaFree(j->buf[p]); if (j->dvec[p]) cudaFree(j->dvec[p]); }
    if (j->partials) cudaFree(j->partials);
    if (j->scratch) cudaFree(j->scratch);
    if (j->sums_dev) cudaFree(j->sums_dev);
    if (j->sums_host) cudaFreeHost(j->sums_host);
    if (j->ints) cudaFree(j->ints);
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
    else { j->opts.device = -1; j->opts.compress = -1; }
    if (j->opts.world < 1) j->opts.world = 1;
    if (j->opts.rank < 0 || j->opts.rank >= j->opts.world) { delete j; return fail(GK_ERR_ARG, "rank outside [0, world)"); }
    int tile = j->opts.tile;
    if (!gk::build_phi_plan(n, sire, dam, m, pro, j->opts.compress, tile, j->plan)) {
        int code = j->plan.error_code == 2 ? GK_ERR_INDEX : GK_ERR_ARG;
        std::string msg = j->plan.error;
        delete j;
        return fail(code, msg);
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
    if (g_cfg_err != cudaSuccess) return cuda_fail(g_cfg_err, "cudaFuncSetAttribute (is this an sm_100 device?)");
    gk::PhiPlan &P = j->plan;
    const int G = P.G;
    if (j->opts.stream) j->stream = (cudaStream_t) j->opts.stream;
    else { GK_CUDA(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking)); j->own_stream = true; }

    // ---- plan arrays: one int32 arena -------------------------------------------------------
    std::vector<int32_t> arena;
    auto push = [&](const std::vector<int32_t> &v) {
        size_t o = arena.size();
        arena.insert(arena.end(), v.begin(), v.end());
        while (arena.size() % 4) arena.push_back(0);     // keep every array 16-byte aligned
        return o;
    };
    j->off.assign(G, StepOffsets{});
    for (int g = 1; g < G; g++) {
        gk::StepPlan &S = P.steps[g];
        StepOffsets &o = j->off[g];
        o.desc = push(S.desc);
    }
    j->off_final_cls = push(P.final_cls);
    j->off_final_ind = push(P.final_ind);
    // result rows of this shard and, for the expansion kernel, its caller entries grouped by class
    gk_shard_rows(P.m, j->opts.rank, j->opts.world, &j->row0, &j->row1);
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

    // ---- matrices: two ping-pong buffers; the m x m result aliases the one the last step read --
    int64_t need[2] = { 0, 0 };
    int64_t maxK = 1;
    for (int g = 0; g < G; g++) {
        const gk::StepPlan &S = P.steps[g];
        need[g & 1] = std::max(need[g & 1], S.K * S.ld + kPadDoubles);
        maxK = std::max(maxK, S.K);
    }
    const int64_t mm = (j->row1 - j->row0) * (int64_t) P.m;
    const int out_buf = P.single_cut ? 0 : 1 - ((G - 1) & 1);
    need[out_buf] = std::max(need[out_buf], mm + kPadDoubles);
    for (int p = 0; p < 2; p++) {
        if (need[p] == 0) continue;
        cudaError_t e = cudaMalloc(&j->buf[p], (size_t) need[p] * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            char msg[256];
            std::snprintf(msg, sizeof msg, "out of device memory: kinship buffers need %.2f GB + %.2f GB of HBM",
                          need[0] * 8 / 1e9, need[1] * 8 / 1e9);
            return fail(GK_ERR_OOM, msg);
        }
        j->buf_doubles[p] = need[p];
        j->device_bytes += need[p] * 8;
    }
    j->out = j->buf[out_buf];
    for (int p = 0; p < 2; p++) { GK_CUDA(cudaMalloc(&j->dvec[p], (size_t) maxK * sizeof(double))); j->device_bytes += maxK * 8; }
    GK_CUDA(cudaMalloc(&j->partials, (size_t) std::max<int64_t>(j->nblocks_expand, 1) * 2 * sizeof(double)));
    GK_CUDA(cudaMalloc(&j->scratch, 2 * 256 * sizeof(double)));
    j->device_bytes += j->nblocks_expand * 16;
    GK_CUDA(cudaMalloc(&j->sums_dev, 2 * sizeof(double)));
    GK_CUDA(cudaHostAlloc(&j->sums_host, 2 * sizeof(double), cudaHostAllocDefault));
    j->ev.resize(G + 2);
    for (auto &e : j->ev) GK_CUDA(cudaEventCreate(&e));
    j->uploaded = true;
    return GK_OK;
}

int gk_phi_run(gk_phi_job *j) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (!j->uploaded) return fail(GK_ERR_STATE, "gk_phi_run before gk_phi_upload");
    GK_CUDA(cudaSetDevice(j->device));
    gk::PhiPlan &P = j->plan;
    const int G = P.G;
    cudaStream_t st = j->stream;
    int64_t launches = 0;
    if (j->opts.verbose) {
        // same text as lib/compute.cpp:662-676
        for (int g = 0; g < G; g++) std::printf("Cut size %d/%d: %d\n", g + 1, G, (int) P.steps[g].cut_size);
        std::printf("Computing the kinship matrix:\n");
    }
    if (!P.single_cut) {
        const gk::StepPlan &S0 = P.steps[0];
        GK_CUDA(gk::launch_init(j->buf[0], j->dvec[0], S0.K, S0.ld, st));
        launches++;
        for (int g = 1; g < G; g++) {
            if (j->opts.verbose) std::printf("Cut %d out of %d\n", g, G);
            const gk::StepPlan &S = P.steps[g];
            const StepOffsets &o = j->off[g];
            gk::StepDev d{};
            d.Gprev = j->buf[(g - 1) & 1]; d.dprev = j->dvec[(g - 1) & 1];
            d.Gnew = j->buf[g & 1]; d.dnew = j->dvec[g & 1];
            d.ld_prev = P.steps[g - 1].ld; d.ld_new = S.ld;
            d.K_prev = S.K_prev; d.K = S.K;
            d.desc = j->ints + o.desc;
            d.nt = S.nt; d.npairs = (int64_t) S.nt * (S.nt + 1) / 2; d.prefetch = gk::step_prefetch_default();
            GK_CUDA(cudaEventRecord(j->ev[g], st));
            GK_CUDA(gk::launch_step(d, P.ranked, st));
            launches++;
        }
        GK_CUDA(cudaEventRecord(j->ev[G], st));
    }
    const int last = G - 1;
    if (P.single_cut) {
        GK_CUDA(gk::launch_eye(j->out, P.m, j->row0, j->row1, j->partials, st));
    } else {
        gk::ExpandDev e{};
        e.G = j->buf[last & 1]; e.d = j->dvec[last & 1]; e.ld = P.steps[last].ld;
        e.cls = j->ints + j->off_final_cls; e.ind = j->ints + j->off_final_ind;
        e.cact = j->ints + j->off_cact; e.coff = j->ints + j->off_coff; e.cmem = j->ints + j->off_cmem;
        e.out = j->out; e.m = P.m; e.row0 = j->row0; e.row1 = j->row1;
        e.nact = j->nact; e.nchunks = j->nchunks; e.partials = j->partials;
        GK_CUDA(gk::launch_expand(e, st));
    }
    GK_CUDA(gk::launch_mean_finish(j->partials, j->nblocks_expand, j->scratch, j->sums_dev, st));
    GK_CUDA(cudaEventRecord(j->ev[G + 1], st));
    launches += 3;
    GK_CUDA(cudaMemcpyAsync(j->sums_host, j->sums_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    j->launches = launches;
    j->ran = true;
    return GK_OK;
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
    s->tile = P.tile; s->m = P.m;
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
        }
    }
    const int64_t mm = (int64_t) P.m * P.m, Kl = P.steps[P.G - 1].K;
    s->cells_device += mm;
    s->algorithmic_bytes_device += 8 * (mm + std::min(mm, Kl * Kl));
    s->device_bytes = j->device_bytes;
    s->kernel_launches = j->launches ? j->launches : (P.single_cut ? 3 : P.G + 3);
    s->plan_seconds = P.plan_seconds;
    int64_t ints = (int64_t) P.final_cls.size() + (int64_t) P.final_ind.size();
    for (int g = 1; g < P.G; g++) {
        const gk::StepPlan &S = P.steps[g];
        ints += (int64_t) S.desc.size();
    }
    s->plan_bytes = ints * 4;
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
    for (int g = 1; g < G && g - 1 < cap; g++) GK_CUDA(cudaEventElapsedTime(&ms[g - 1], j->ev[g], j->ev[g + 1]));
    if (G - 1 < cap) GK_CUDA(cudaEventElapsedTime(&ms[G - 1], j->ev[G], j->ev[G + 1]));   // expansion + mean
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
    if (!j || !v || step < 1 || step >= j->plan.G) return fail(GK_ERR_ARG, "bad step");
    const gk::PhiPlan &P = j->plan;
    const gk::StepPlan &S = P.steps[step];
    v->K = S.K; v->ld = S.ld; v->K_prev = S.K_prev; v->nt = S.nt; v->tile = P.tile; v->win = P.win;
    v->lmax = P.lmax; v->nr_max = S.nr_max;
    v->desc = S.desc.data(); v->desc_words = gk::kDescWords; v->conf_max = gk::kConfMax;
    v->off_list = gk::kDescList; v->off_conf = gk::kDescConf; v->off_la = gk::kDescLa;
    v->off_pind = gk::kDescPind; v->off_rank = gk::kDescRank; v->ranked = P.ranked ? 1 : 0;
    return GK_OK;
}
int gk_phi_final_view(gk_phi_job *j, const int32_t **cls, const int32_t **ind, int64_t *K_last) {
    if (!j) return fail(GK_ERR_ARG, "null job");
    if (cls) *cls = j->plan.final_cls.data();
    if (ind) *ind = j->plan.final_ind.data();
    if (K_last) *K_last = j->plan.steps[j->plan.G - 1].K;
    return GK_OK;
}
int gk_phi_initial_classes(gk_phi_job *j, int64_t *K0) {
    if (!j || !K0) return fail(GK_ERR_ARG, "null argument");
    *K0 = j->plan.steps[0].K;
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
