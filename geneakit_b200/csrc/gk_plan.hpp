// Host planner of the B200 kinship path (pure C++17, no CUDA).
//
// Turns a rank-space pedigree + proband list into the per-generation STEP PLAN the sm_100a
// kernels consume.  It replaces, on the host, the reference's cut_vertices()
// (lib/compute.cpp:119-135 and helpers :28-114) and the slot bookkeeping of compute_kinships()
// (:678-689), but not with their std::set algebra:
//
//   * cuts:  cut(L) = { x : minpath(x) <= L <= maxpath(x) } where min/maxpath are the
//     shortest / longest ancestral path lengths from a proband; one reverse-topological sweep,
//     O(N+E).  Cut i of the reference is cut(L = G-1-i).
//   * classes: the members of a cut that entered it together and have the SAME two parents have
//     identical kinship rows (except with themselves), so one matrix row per sibship ("class")
//     is stored; individuals kept from the previous cut are singleton classes whose two
//     "parents" are their own old row.  compress=0 gives one class per individual.
//     Convention for the class matrix A of a cut: A[c][c'] (c != c') = kinship between any member
//     of c and any member of c';  A[c][c] = kinship between two DISTINCT members of c when the
//     class has several members, and the self kinship when it is a singleton.  d[c] = self kinship.
//   * one recurrence step is B = P A P^T (+ sparse corrections), P = the K x K_prev matrix with
//     1/2 at (i, F_i) and (i, M_i) (F_i / M_i = class rows of the parents of class i).  The kernel
//     factorises it:  N_I = P_I A  for a PANEL I of 32 classes (whole rows of A: contiguous),
//     kept transposed in an L2-resident ring, then  B[I][J] = N_I P_J^T  for every tile J.
//     Nothing constrains the ORDER of the classes any more, so they are ordered for locality:
//     a greedy walk over the "shares a parent class" graph makes consecutive classes share parent
//     rows (fewer distinct rows of A per panel = fewer HBM reads).
//   * corrections ("same individual -> self kinship", lib/compute.cpp:522-528): wherever the two
//     classes i, j have a parent INDIVIDUAL p in common, the term A[c_p][c_p] must read d[c_p]:
//     B[i][j] += 1/4 * mult * (d[c_p] - A[c_p][c_p]).  They form a short sparse list (one entry per
//     parent individual in monogamous pedigrees) applied by the diagonal kernel.
//
// Rounding: for G-1 <= 26 cuts below the founders every value is a dyadic rational that FP64
// holds exactly, so any evaluation order is bit-identical to the reference.  Deeper pedigrees
// switch to RANKED mode: no compression (no corrections either: every class is a singleton whose
// diagonal already is the self kinship) and classes in DESCENDING rank order, so that the tile
// pairs J <= I the kernel evaluates are exactly "N of the lower-ranked individual, parents of the
// higher-ranked one" -- the grouping lib/compute.cpp:529-548 rounds with.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace gk {

constexpr int kPanel = 32;          // classes per panel / tile
constexpr int kFlagSingleton = 1;   // class has exactly one member: its diagonal holds the self kinship
constexpr int kFlagKept = 2;        // member kept from the previous cut (both "parents" = its own old row)
constexpr int kFlagBoth = 4;        // new class with both parents known
constexpr int kZeroSlot = 2 * kPanel;   // staged row 64 = zeros
constexpr int kSub = 8;                 // classes per sub-panel
constexpr int kZeroSlot8 = 2 * kSub;    // staged row 16 = zeros
constexpr int kUnitStageBytes = 20 * 514 * 8;   // a stage of the default step kernel: 20 rows x 512 columns (pitch 514 doubles), or 39 x 256 (pitch 258)
constexpr int kUnitRowsMax = 40;        // staged positions of a unit: <= 32 distinct parent rows + the zero row, spread over 8 bank colours

struct StepPlan {
    int64_t K = 0, Kpad = 0, K_prev = 0, Kpad_prev = 0;
    int np = 0;                // panels = Kpad / 32
    int64_t cut_size = 0;      // |cut_g| as the reference counts it (last cut: m, duplicates included)
    int64_t n_kept = 0;        // members also in the previous cut
    int64_t P_ref = 0;         // distinct individuals of cut g-1 that are a parent of a NEW member
    int64_t P_cls = 0;         // distinct rows of matrix g-1 read by this step
    int64_t kept_cls = 0;      // singleton classes of kept individuals
    int64_t panel_rows = 0;    // sum over panels of the distinct parent rows the panel reads
    std::vector<int32_t> pF, pM;       // Kpad: parent class rows in matrix g-1, -1 = none (padding: -1, -1)
    // de-duplicated parent rows per panel: prow[64 p + s] = s-th distinct parent row of panel p (-1 past the
    // end), pcnt[p] = how many; slots[i] = slot of class i's father row | slot of its mother row << 8
    // (kZeroSlot = unknown parent: an all-zero staged row)
    std::vector<int32_t> first_use;    // Kpad_prev: lowest tile (panel index) of this cut holding a child of previous class c; np if none
    std::vector<int32_t> prow, pcnt, slots;
    // the same per SUB-PANEL of 8 classes (the TMA phase-1 items of the multi-GPU kernel): 16 slots each,
    // kZeroSlot8 = unknown parent
    std::vector<int32_t> prow8, pcnt8, slots8;
    int64_t sub_rows = 0;      // sum over sub-panels of their distinct parent rows
    // per UNIT of 16 classes (the phase-1 items of the default step kernel): urow[kUnitRowsMax u + p] = parent row staged at
    // position p (-1: unused position; Kpad_prev: the all-zero row of an unknown parent), ucnt[u] = positions | rows << 8 |
    // (items are 512 columns wide) << 16, uslots[i] = staged position of class i's father | of its mother << 8
    std::vector<int32_t> urow, ucnt, uslots;
    bool has_panel = false, has_sub = false, has_units = false;   // which of the three row lists this cut's kernel needs (KernelChoice)
    std::vector<int32_t> flags;        // Kpad
    std::vector<int32_t> cls_ind;      // K: lowest-ranked member (= the individual, uncompressed)
    std::vector<int32_t> cls_size;     // K: members
    // corrections, CSR over (i, j) groups with i >= j
    std::vector<int32_t> pg_i, pg_j, pg_off, pt_cls, pt_mult;
};

struct PhiPlan {
    int32_t n = 0, m = 0, G = 0;
    bool compressed = true, ranked = false, single_cut = false;
    std::vector<StepPlan> steps;            // steps[g], g = 0 .. G-1 (steps[0]: founders' cut, no parents)
    std::vector<int32_t> final_cls, final_ind;   // m entries, caller order
    double plan_seconds = 0.0;
    std::string error;
    int error_code = 0;
};

inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Cut sizes only (for gk_cut_sizes / gk_phi_required_gb).  Returns G or <0.
inline int cut_sizes_only(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m,
                          const int32_t *pro, std::vector<int64_t> &sizes, std::string &err) {
    sizes.clear();
    if (n < 0 || m < 0 || (n > 0 && (!sire || !dam)) || (m > 0 && !pro)) { err = "null or negative argument"; return -1; }
    for (int32_t i = 0; i < m; i++)
        if (pro[i] < 0 || pro[i] >= n) { err = "proband rank out of range"; return -2; }
    if (m == 0) return 0;
    std::vector<int32_t> minp(n, -1), maxp(n, -1);
    for (int32_t i = 0; i < m; i++) { minp[pro[i]] = 0; maxp[pro[i]] = std::max(maxp[pro[i]], 0); }
    int32_t top = 0;
    for (int32_t x = n - 1; x >= 0; x--) {
        if (maxp[x] < 0) continue;
        top = std::max(top, maxp[x]);
        const int32_t par[2] = { sire[x], dam[x] };
        for (int q = 0; q < 2; q++) {
            int32_t p = par[q];
            if (p < 0) continue;
            if (p >= x) { err = "parent does not precede child in rank order"; return -1; }
            minp[p] = (minp[p] < 0) ? minp[x] + 1 : std::min(minp[p], minp[x] + 1);
            maxp[p] = std::max(maxp[p], maxp[x] + 1);
        }
    }
    int G = top + 1;
    sizes.assign(G, 0);
    for (int32_t x = 0; x < n; x++) {
        if (maxp[x] < 0) continue;
        for (int L = minp[x]; L <= maxp[x]; L++) sizes[G - 1 - L]++;
    }
    sizes[G - 1] = m;   // lib/compute.cpp:132 : last cut := caller's vector, duplicates kept
    return G;
}

struct ClassRec {
    int32_t pi0, pi1, pc0, pc1;   // parent individuals (-1 missing; kept: both = the individual), their natural class ids in the previous cut
    int32_t first_member;         // lowest individual rank
    int32_t size;
    uint8_t kept;
};

// Which step kernel a job runs per cut decides which plan arrays that cut needs (gk_capi.cu reads the same choice).
struct KernelChoice {
    int use_g4 = 1;        // default step kernel for large cuts (TMA bulk-copy phase 1, L2-direct phase 2)
    int use_tma = 0;       // sharded jobs: the kernel with TMA phase 1 on 8-class items (small cuts)
    int g4_min_k = 8192;   // smallest previous-cut matrix the default kernel is used for
    static KernelChoice resolve(int world, bool ranked) {
        KernelChoice k;
        k.use_tma = world > 1 ? 1 : 0;
        if (const char *e = std::getenv("GK_STEP_KERNEL")) { k.use_tma = (e[0] == 't') ? 1 : 0; k.use_g4 = (e[0] == 'g') ? 1 : 0; }
        if (const char *e = std::getenv("GK_G4_MIN_K")) { const int v = std::atoi(e); if (v > 0) k.g4_min_k = v; }
        return k;
    }
    bool cut_uses_g4(int64_t Kpad_prev) const { return use_g4 && Kpad_prev >= g4_min_k; }
};

// Greedy walk over (a subset of) the classes of the new cut: move from a class to an unvisited class
// that shares one of its parent rows (preferring the row we did NOT arrive through), so that
// consecutive classes -- hence panels of 32 -- read few distinct rows of the previous matrix.
// `subset` lists the classes to order (indices into cls); their order is appended to `order`.
inline void locality_order(const std::vector<ClassRec> &cls, const std::vector<int32_t> &subset, int64_t K_prev,
                           std::vector<int32_t> &order) {
    const int32_t K = (int32_t) subset.size();
    if (K == 0) return;
    if (K_prev <= 0) { for (int32_t i : subset) order.push_back(i); return; }
    std::vector<int32_t> off((size_t) K_prev + 1, 0);
    auto rows_of = [&](const ClassRec &c, int32_t r[2]) {
        int k = 0;
        if (c.pc0 >= 0) r[k++] = c.pc0;
        if (c.pc1 >= 0 && c.pc1 != c.pc0) r[k++] = c.pc1;
        return k;
    };
    // local ids 0..K-1 <-> subset[local]
    for (int32_t i = 0; i < K; i++) { int32_t r[2]; int k = rows_of(cls[subset[i]], r); for (int q = 0; q < k; q++) off[r[q] + 1]++; }
    for (int64_t c = 0; c < K_prev; c++) off[c + 1] += off[c];
    std::vector<int32_t> adj(off[K_prev]), cur(off.begin(), off.end() - 1);
    for (int32_t i = 0; i < K; i++) { int32_t r[2]; int k = rows_of(cls[subset[i]], r); for (int q = 0; q < k; q++) adj[cur[r[q]]++] = i; }
    std::copy(off.begin(), off.end() - 1, cur.begin());      // cursor: first possibly-unvisited entry of each row's list
    std::vector<uint8_t> visited(K, 0);
    auto next_unvisited = [&](int32_t row) -> int32_t {
        int32_t &p = cur[row];
        while (p < off[row + 1] && visited[adj[p]]) p++;
        return p < off[row + 1] ? adj[p] : -1;
    };
    // walk starts: classes whose parent rows have the fewest users first (path ends before path middles); a stable
    // counting sort by that degree
    std::vector<int32_t> starts(K), deg(K);
    int32_t maxdeg = 0;
    for (int32_t i = 0; i < K; i++) {
        int32_t r[2]; int k = rows_of(cls[subset[i]], r);
        int32_t d = 0;
        for (int q = 0; q < k; q++) d += off[r[q] + 1] - off[r[q]];
        deg[i] = d; maxdeg = std::max(maxdeg, d);
    }
    {
        std::vector<int32_t> cnt((size_t) maxdeg + 2, 0);
        for (int32_t i = 0; i < K; i++) cnt[deg[i] + 1]++;
        for (int32_t d = 0; d <= maxdeg; d++) cnt[d + 1] += cnt[d];
        for (int32_t i = 0; i < K; i++) starts[cnt[deg[i]]++] = i;
    }
    size_t sp = 0;
    int32_t done = 0;
    std::vector<int32_t> trail;                                    // the current walk, for backtracking
    while (done < K) {
        while (visited[starts[sp]]) sp++;
        int32_t x = starts[sp], came = -1;
        trail.clear();
        for (;;) {
            visited[x] = 1; order.push_back(subset[x]); done++; trail.push_back(x);
            int32_t r[2]; int k = rows_of(cls[subset[x]], r);
            if (k == 2 && came == r[0]) std::swap(r[0], r[1]);     // try the row we did not arrive through first
            int32_t nxt = -1;
            for (int q = 0; q < k && nxt < 0; q++) { nxt = next_unvisited(r[q]); if (nxt >= 0) came = r[q]; }
            // dead end: back up along the walk (a few panels at most) to a class that still has an unvisited
            // neighbour, so that the next classes stay close to rows already staged (c4: 40.1 -> 37.5 distinct
            // parent rows per 32 classes)
            for (int depth = 0; nxt < 0 && depth < 64 && !trail.empty(); depth++) {
                const int32_t b = trail.back(); trail.pop_back();
                int32_t rb[2]; const int kb = rows_of(cls[subset[b]], rb);
                for (int q = 0; q < kb && nxt < 0; q++) { nxt = next_unvisited(rb[q]); if (nxt >= 0) came = rb[q]; }
            }
            if (nxt < 0) break;
            x = nxt;
        }
    }
}

// Run fn(i) for i in [i0, i1) on up to `threads` host threads.
template <class F>
inline void parallel_for(int i0, int i1, int threads, F fn) {
    if (threads <= 1 || i1 - i0 <= 1) { for (int i = i0; i < i1; i++) fn(i); return; }
    std::atomic<int> next(i0);
    std::vector<std::thread> pool;
    const int nt = std::min(threads, i1 - i0);
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&]() { for (int i = next.fetch_add(1); i < i1; i = next.fetch_add(1)) fn(i); });
    for (auto &th : pool) th.join();
}

// Per-cut scratch of the planner.  Stage A fills the classes (in NATURAL order: new sibships sorted by (father, mother),
// then the members kept from the previous cut by rank), stage B the order of the classes, stage C the arrays the kernels read.
struct CutWork {
    const int32_t *members = nullptr;              // individuals of the cut, ascending
    int64_t nm = 0;
    std::vector<int32_t> nat_of;                   // natural class id of members[k]
    std::vector<ClassRec> cls;                     // natural order; pc0 / pc1 = NATURAL ids in the previous cut (stage B)
    std::vector<int32_t> pos;                      // natural id -> position in the class matrix
    int64_t n_kept = 0, P_ref = 0;
    int32_t base = 0;
    std::vector<int32_t> table;                    // table[ind - base] = index into members (built when the cut's rank range is compact)
    void build_table() {
        if (nm == 0) return;
        base = members[0];
        const int64_t range = (int64_t) members[nm - 1] - base + 1;
        if (range > 8 * nm + 1024) return;         // sparse cut: keep the binary search
        table.assign((size_t) range, -1);
        for (int64_t k = 0; k < nm; k++) table[members[k] - base] = (int32_t) k;
    }
    int32_t midx(int32_t ind) const {              // index of an individual of this cut in members
        if (!table.empty()) return table[ind - base];
        return (int32_t) (std::lower_bound(members, members + nm, ind) - members);
    }
    int32_t nat(int32_t ind) const { return nat_of[midx(ind)]; }   // its natural class id
};

// The planner of one job.  begin() validates the inputs, finds the cuts and the classes of every cut (so every matrix
// size is known: the caller can allocate), start() plans the cuts IN ORDER on a pool of host threads, and wait_cut(g)
// returns as soon as the arrays of cut g are complete -- the driver copies them to the device and launches step g while
// the later cuts are still being planned.  finish() joins the pool.
class PhiPlanner {
public:
    PhiPlanner(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
               int compress_opt, int order_opt, int world, PhiPlan &plan, int max_threads = 0)
        : n_(n), m_(m), sire_(sire), dam_(dam), pro_(pro), compress_opt_(compress_opt), order_opt_(order_opt),
          world_(world < 1 ? 1 : world), P(plan) {
        t0_ = std::chrono::steady_clock::now();
        threads_ = (int) std::thread::hardware_concurrency();
        threads_ = std::max(1, std::min(threads_, 24));
        if (max_threads > 0) threads_ = std::min(threads_, max_threads);
        if (const char *e = std::getenv("GK_PLAN_THREADS")) threads_ = std::max(1, std::atoi(e));
        timing_ = std::getenv("GK_PLAN_TIMING") != nullptr;
        build_all_ = std::getenv("GK_PLAN_ALL") != nullptr;
    }
    ~PhiPlanner() { join(); }
    PhiPlanner(const PhiPlanner &) = delete;
    PhiPlanner &operator=(const PhiPlanner &) = delete;

    bool begin();
    // after_cut(g) runs on the worker that finished cut g, before wait_cut(g) returns (the driver packs its device view there)
    void start(std::function<void(int)> after_cut = {});
    bool wait_cut(int g);
    bool finish();
    int threads() const { return threads_; }
    KernelChoice choice;                 // valid after begin()
    bool build_all_ = false;             // tests / tools: build the row lists of every kernel for every cut

private:
    void mark(const char *what) {
        if (timing_) std::fprintf(stderr, "[plan] %-28s %.1f ms\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count());
    }
    int enter(int32_t x) const { return G_ - 1 - pl_[2 * (size_t) x + 1]; }
    void stage_a(int g);
    void stage_b(int g);
    void stage_c(int g);
    void join() { for (auto &t : pool_) if (t.joinable()) t.join(); pool_.clear(); }
    void fail_all() { failed_any_ = true; for (int g = 0; g < G_; g++) state_[g].store(-1, std::memory_order_release); }

    int32_t n_, m_;
    const int32_t *sire_, *dam_, *pro_;
    int compress_opt_, order_opt_, world_;
    PhiPlan &P;
    int threads_ = 1, G_ = 0;
    bool timing_ = false;
    std::chrono::steady_clock::time_point t0_;
    std::vector<int32_t> pl_;            // (min, max) ancestral path length from a proband per individual, interleaved; -1 = not an ancestor
    std::vector<int32_t> members_;       // members of every cut, concatenated
    std::vector<int64_t> moff_;          // G + 1
    std::vector<uint8_t> owner_;         // multi-GPU: lineage owner of every individual that takes part
    std::vector<CutWork> W;
    std::unique_ptr<std::atomic<int>[]> state_;    // per cut: 0 = not ordered, 1 = ordered (stage B), 2 = complete, -1 = failed
    std::atomic<int> next_{0};
    std::atomic<bool> failed_{false}, over_budget_{false}, failed_any_{false};
    std::vector<std::thread> pool_;
    std::function<void(int)> after_cut_;
};

inline bool PhiPlanner::begin() {
    const int32_t n = n_, m = m_;
    const int32_t *sire = sire_, *dam = dam_, *pro = pro_;
    P = PhiPlan();
    P.n = n; P.m = m;
    if (n <= 0 || m <= 0 || !sire || !dam || !pro) {
        P.error = "gk_phi: empty pedigree or proband list (the binding substitutes get_proband_ids for an empty list)";
        P.error_code = 1; return false;
    }
    for (int32_t i = 0; i < m; i++)
        if (pro[i] < 0 || pro[i] >= n) { P.error = "proband rank out of range"; P.error_code = 2; return false; }

    // ---- min / max ancestral path length from a proband (one reverse-topological sweep) ----
    pl_.assign(2 * (size_t) n, -1);
    int32_t *pl = pl_.data();
    mark("  (path-length array)");
    for (int32_t i = 0; i < m; i++) { pl[2 * (size_t) pro[i]] = 0; pl[2 * (size_t) pro[i] + 1] = std::max(pl[2 * (size_t) pro[i] + 1], 0); }
    int32_t top = 0;
    for (int32_t x = n - 1; x >= 0; x--) {
        if (x >= 16) {                               // the parents' entries are random accesses: fetch them ahead
            const int32_t fs = sire[x - 16], fd = dam[x - 16];
            if (fs >= 0) __builtin_prefetch(pl + 2 * (size_t) fs, 1);
            if (fd >= 0) __builtin_prefetch(pl + 2 * (size_t) fd, 1);
        }
        const int32_t mx = pl[2 * (size_t) x + 1];
        if (mx < 0) continue;
        const int32_t mn = pl[2 * (size_t) x];
        top = std::max(top, mx);
        const int32_t par[2] = { sire[x], dam[x] };
        for (int q = 0; q < 2; q++) {
            const int32_t p = par[q];
            if (p < 0) continue;
            if (p >= x) { P.error = "parent does not precede child in rank order"; P.error_code = 1; return false; }
            int32_t &pmn = pl[2 * (size_t) p], &pmx = pl[2 * (size_t) p + 1];
            pmn = (pmn < 0) ? mn + 1 : std::min(pmn, mn + 1);
            pmx = std::max(pmx, mx + 1);
        }
    }
    const int G = top + 1;
    G_ = G;
    P.G = G;
    P.single_cut = (G == 1);
    P.ranked = (G - 1 > 26);
    P.compressed = P.ranked ? false : (compress_opt_ != 0);
    P.steps.assign(G, StepPlan());
    choice = KernelChoice::resolve(world_, P.ranked);
    mark("path lengths");

    // ---- members of every cut (x belongs to cuts enter(x) = G-1-maxp .. leave(x) = G-1-minp), ascending: counted and
    //      filled by ranges of individuals in parallel ----
    {
        const int T = std::max(1, std::min(threads_, n / 65536 + 1));
        std::vector<std::vector<int64_t>> cnt(T, std::vector<int64_t>((size_t) G, 0));
        auto range = [&](int t, int32_t &x0, int32_t &x1) { x0 = (int32_t) ((int64_t) n * t / T); x1 = (int32_t) ((int64_t) n * (t + 1) / T); };
        parallel_for(0, T, T, [&](int t) {
            int32_t x0, x1; range(t, x0, x1);
            int64_t *c = cnt[t].data();
            for (int32_t x = x0; x < x1; x++) {
                const int32_t mx = pl[2 * (size_t) x + 1];
                if (mx < 0) continue;
                for (int g = G - 1 - mx; g <= G - 1 - pl[2 * (size_t) x]; g++) c[g]++;
            }
        });
        moff_.assign((size_t) G + 1, 0);
        for (int g = 0; g < G; g++) {
            int64_t tot = 0;
            for (int t = 0; t < T; t++) { const int64_t c = cnt[t][g]; cnt[t][g] = tot; tot += c; }     // offset of thread t inside cut g
            moff_[g + 1] = moff_[g] + tot;
        }
        members_.resize((size_t) moff_[G]);
        parallel_for(0, T, T, [&](int t) {
            int32_t x0, x1; range(t, x0, x1);
            int64_t *c = cnt[t].data();
            for (int32_t x = x0; x < x1; x++) {
                const int32_t mx = pl[2 * (size_t) x + 1];
                if (mx < 0) continue;
                for (int g = G - 1 - mx; g <= G - 1 - pl[2 * (size_t) x]; g++) members_[(size_t) (moff_[g] + c[g]++)] = x;
            }
        });
    }
    W.assign(G, CutWork());
    for (int g = 0; g < G; g++) { W[g].members = members_.data() + moff_[g]; W[g].nm = moff_[g + 1] - moff_[g]; }
    mark("members");
    // multi-GPU: lineage owner of every individual that takes part (one top-down sweep)
    if (world_ > 1) {
        owner_.assign((size_t) n, 0);
        int32_t rr = 0;
        for (int32_t x = 0; x < n; x++) {
            if (pl[2 * (size_t) x + 1] < 0) continue;
            const int32_t f = sire[x], d = dam[x];
            if (f >= 0 && pl[2 * (size_t) f + 1] >= 0) owner_[x] = owner_[f];
            else if (d >= 0 && pl[2 * (size_t) d + 1] >= 0) owner_[x] = owner_[d];
            else owner_[x] = (uint8_t) (rr++ % world_);
        }
    }
    // ---- stage A (parallel): index of every cut, then its classes in natural order --------------------------------
    parallel_for(0, G, threads_, [&](int g) { W[g].build_table(); });
    parallel_for(0, G, threads_, [&](int g) { stage_a(g); });
    for (int g = 0; g < G; g++) {
        StepPlan &S = P.steps[g];
        const int64_t K = (int64_t) W[g].cls.size();
        S.K = K;
        S.Kpad = round_up(std::max<int64_t>(K, 1), kPanel);
        S.np = (int) (S.Kpad / kPanel);
        S.K_prev = g > 0 ? P.steps[g - 1].K : 0;
        S.Kpad_prev = g > 0 ? P.steps[g - 1].Kpad : 0;
        S.n_kept = W[g].n_kept; S.P_ref = W[g].P_ref;
        S.cut_size = (g == G - 1) ? (int64_t) m : W[g].nm;
    }
    mark("stage A (classes)");
    return true;
}

// Classes of cut g.  New members with equal (father, mother) share a class; members kept from the previous cut and the
// members of uncompressed jobs are singletons.  Linear: the sibships are chained on their father's (else their mother's)
// position in the previous cut -- every parent of a new member belongs to the previous cut.
inline void PhiPlanner::stage_a(int g) {
    CutWork &C = W[g];
    const CutWork *Cp = g > 0 ? &W[g - 1] : nullptr;
    const int64_t nm = C.nm;
    C.nat_of.assign((size_t) nm, -1);
    if (!P.compressed) {
        C.cls.resize((size_t) nm);
        std::vector<uint8_t> seen(Cp ? (size_t) Cp->nm : 0, 0);
        for (int64_t k = 0; k < nm; k++) {
            const int32_t x = C.members[k];
            ClassRec &c = C.cls[k];
            c.kept = enter(x) < g;
            c.pi0 = c.kept ? x : sire_[x]; c.pi1 = c.kept ? x : dam_[x];
            c.pc0 = c.pc1 = -1; c.first_member = x; c.size = 1;
            C.nat_of[k] = (int32_t) k;
            if (c.kept) C.n_kept++;
            else if (Cp) for (int q = 0; q < 2; q++) {
                const int32_t p = q ? c.pi1 : c.pi0;
                if (p < 0) continue;
                const int32_t a = Cp->midx(p);
                if (!seen[a]) { seen[a] = 1; C.P_ref++; }
            }
        }
        return;
    }
    struct Tmp { int32_t pi0, pi1, first, size, next; };
    std::vector<Tmp> tmp;
    const int64_t np = Cp ? Cp->nm : 0;
    std::vector<int32_t> headS((size_t) np, -1), headD((size_t) np, -1), tmp_of((size_t) nm, -1);
    int32_t orphan = -1;                          // the class of the new members without any known parent
    for (int64_t k = 0; k < nm; k++) {
        const int32_t x = C.members[k];
        if (enter(x) < g) { C.n_kept++; continue; }
        const int32_t s = sire_[x], d = dam_[x];
        int32_t *head = s >= 0 ? &headS[Cp->midx(s)] : (d >= 0 ? &headD[Cp->midx(d)] : &orphan);
        int32_t t = *head;
        while (t >= 0 && !(tmp[t].pi0 == s && tmp[t].pi1 == d)) t = tmp[t].next;
        if (t < 0) { t = (int32_t) tmp.size(); tmp.push_back({ s, d, x, 0, *head }); *head = t; }
        tmp[t].size++;
        tmp_of[k] = t;
    }
    // natural order: ascending (father + 1, mother + 1), an unknown parent first
    std::vector<int32_t> nat_of_tmp(tmp.size(), -1);
    C.cls.reserve(tmp.size() + (size_t) C.n_kept);
    auto emit = [&](int32_t t) {
        nat_of_tmp[t] = (int32_t) C.cls.size();
        ClassRec c;
        c.kept = 0; c.pi0 = tmp[t].pi0; c.pi1 = tmp[t].pi1; c.pc0 = c.pc1 = -1; c.first_member = tmp[t].first; c.size = tmp[t].size;
        C.cls.push_back(c);
    };
    if (orphan >= 0) emit(orphan);
    for (int64_t a = 0; a < np; a++) if (headD[a] >= 0) emit(headD[a]);          // (unknown, mother): at most one class per mother
    std::vector<int32_t> chain;
    for (int64_t a = 0; a < np; a++) {
        if (headS[a] < 0) continue;
        if (tmp[headS[a]].next < 0) { emit(headS[a]); continue; }
        chain.clear();
        for (int32_t t = headS[a]; t >= 0; t = tmp[t].next) chain.push_back(t);
        std::sort(chain.begin(), chain.end(), [&](int32_t u, int32_t v) { return tmp[u].pi1 < tmp[v].pi1; });
        for (int32_t t : chain) emit(t);
    }
    for (int64_t k = 0; k < nm; k++) {
        if (tmp_of[k] >= 0) { C.nat_of[k] = nat_of_tmp[tmp_of[k]]; continue; }
        const int32_t x = C.members[k];
        C.nat_of[k] = (int32_t) C.cls.size();
        ClassRec c;
        c.kept = 1; c.pi0 = c.pi1 = x; c.pc0 = c.pc1 = -1; c.first_member = x; c.size = 1;
        C.cls.push_back(c);
    }
    // distinct parents of the new classes
    std::vector<uint8_t> seen((size_t) np, 0);
    for (const Tmp &t : tmp)
        for (int q = 0; q < 2; q++) {
            const int32_t p = q ? t.pi1 : t.pi0;
            if (p < 0) continue;
            const int32_t a = Cp->midx(p);
            if (!seen[a]) { seen[a] = 1; C.P_ref++; }
        }
}

// Parent classes (natural ids in the previous cut) and the order of the classes.  The order only needs to know WHICH
// classes share a parent row, so the cuts are ordered independently of one another.
inline void PhiPlanner::stage_b(int g) {
    CutWork &C = W[g];
    const int64_t K = (int64_t) C.cls.size();
    const int64_t K_prev = g > 0 ? (int64_t) W[g - 1].cls.size() : 0;
    if (g > 0)
        for (ClassRec &c : C.cls) {
            c.pc0 = c.pi0 >= 0 ? W[g - 1].nat(c.pi0) : -1;
            c.pc1 = c.pi1 >= 0 ? W[g - 1].nat(c.pi1) : -1;
        }
    std::vector<int32_t> order;
    order.reserve(K);
    if (P.ranked) {
        // descending rank: the natural order of an uncompressed cut is ascending rank
        order.resize(K);
        for (int64_t i = 0; i < K; i++) order[i] = (int32_t) (K - 1 - i);
    } else if (order_opt_ == 0 || g == 0) {
        order.resize(K);
        for (int64_t i = 0; i < K; i++) order[i] = (int32_t) i;
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            const ClassRec &x = C.cls[a], &y = C.cls[b];
            if (x.pc0 != y.pc0) return x.pc0 < y.pc0;
            if (x.pc1 != y.pc1) return x.pc1 < y.pc1;
            return x.first_member < y.first_member;
        });
    } else {
        std::vector<int32_t> subset;
        if (world_ <= 1) {
            subset.resize(K);
            for (int64_t i = 0; i < K; i++) subset[i] = (int32_t) i;
            locality_order(C.cls, subset, K_prev, order);
        } else {
            // Multi-GPU: the rows of every matrix are sharded in equal blocks of panels.  Group the classes by the
            // LINEAGE OWNER of their first known parent (owner[x] = owner of x's father, else of its mother; founders
            // dealt round-robin), then order each group for locality.  A class and its first parent's class fall into
            // the same group of their respective cuts, so -- up to the spill at the block boundaries -- that parent's
            // row is a LOCAL read for the rank that owns the class.  The owner of an individual does not depend on the
            // order of any cut, so the cuts are still ordered independently (in parallel).
            std::vector<std::vector<int32_t>> groups(world_);
            for (int64_t i = 0; i < K; i++) {
                const int32_t pi = C.cls[i].pi0 >= 0 ? C.cls[i].pi0 : C.cls[i].pi1;
                const int q = pi >= 0 ? owner_[pi] : owner_[C.cls[i].first_member];
                groups[q].push_back((int32_t) i);
            }
            for (int q = 0; q < world_; q++) locality_order(C.cls, groups[q], K_prev, order);
        }
    }
    C.pos.assign(K, 0);
    for (int64_t p = 0; p < K; p++) C.pos[order[p]] = (int32_t) p;
}

// The arrays the kernels read for cut g (needs the order of cut g - 1).
inline void PhiPlanner::stage_c(int g) {
    const int G = G_;
    StepPlan &S = P.steps[g];
    const CutWork &C = W[g];
    const CutWork *Cp = g > 0 ? &W[g - 1] : nullptr;
    const int64_t K = S.K;
    S.pF.assign((size_t) S.Kpad, -1); S.pM.assign((size_t) S.Kpad, -1); S.flags.assign((size_t) S.Kpad, 0);
    S.cls_ind.resize(K); S.cls_size.resize(K);
    std::vector<int32_t> nat_at(K);                    // position -> natural id
    for (int64_t i = 0; i < K; i++) nat_at[C.pos[i]] = (int32_t) i;
    for (int64_t p = 0; p < K; p++) {
        const ClassRec &c = C.cls[nat_at[p]];
        if ((c.pi0 >= 0 && g > 0 && c.pc0 < 0) || (c.pi1 >= 0 && g > 0 && c.pc1 < 0)) failed_ = true;
        S.pF[p] = c.pc0 >= 0 ? Cp->pos[c.pc0] : -1;
        S.pM[p] = c.pc1 >= 0 ? Cp->pos[c.pc1] : -1;
        S.flags[p] = (c.size == 1 ? kFlagSingleton : 0) | (c.kept ? kFlagKept : 0) |
                     ((!c.kept && c.pc0 >= 0 && c.pc1 >= 0) ? kFlagBoth : 0);
        S.cls_ind[p] = c.first_member; S.cls_size[p] = c.size;
        if (c.kept) S.kept_cls++;
    }
    if (g == G - 1) {
        // ---- final expansion map (caller order, duplicates expand to the same row) --------------
        P.final_cls.resize(m_); P.final_ind.resize(m_);
        for (int32_t i = 0; i < m_; i++) { P.final_ind[i] = pro_[i]; P.final_cls[i] = C.pos[C.nat(pro_[i])]; }
    }
    if (g == 0) return;
    const bool g4 = choice.cut_uses_g4(S.Kpad_prev);
    S.has_units = g4 || build_all_;
    S.has_panel = !g4 || build_all_;                                   // both other kernels stage whole panels in phase 2
    S.has_sub = (!g4 && choice.use_tma) || build_all_;

    // ---- distinct parent rows: overall, per panel (64 slots), per sub-panel of 8 classes (16 slots) ----
    std::vector<int32_t> stamp((size_t) S.K_prev, -1), unit_stamp((size_t) S.K_prev, -1), slot_of((size_t) S.K_prev, 0);
    if (S.has_panel) {
        S.prow.assign((size_t) S.np * 2 * kPanel, -1);
        S.slots.assign((size_t) S.Kpad, kZeroSlot | (kZeroSlot << 8));
    }
    S.pcnt.assign((size_t) S.np, 0);
    int64_t pcls = 0, prow = 0;
    for (int64_t pos = 0; pos < K; pos++) {
        const int32_t panel = (int32_t) (pos / kPanel);
        int32_t sl[2] = { kZeroSlot, kZeroSlot };
        for (int q = 0; q < 2; q++) {
            const int32_t pc = q ? S.pM[pos] : S.pF[pos];
            if (pc < 0) continue;
            if (stamp[pc] != g) { stamp[pc] = g; pcls++; }
            if (unit_stamp[pc] != panel) {
                unit_stamp[pc] = panel; prow++;
                slot_of[pc] = S.pcnt[panel]++;
                if (S.has_panel) S.prow[(size_t) panel * 2 * kPanel + slot_of[pc]] = pc;
            }
            sl[q] = slot_of[pc];
        }
        if (S.has_panel) S.slots[pos] = sl[0] | (sl[1] << 8);
    }
    S.P_cls = pcls; S.panel_rows = prow;
    // a row N^T[c] of panel I's transposed panel is only ever read by the tiles J <= I that hold a child of c
    S.first_use.assign((size_t) S.Kpad_prev, S.np);
    for (int64_t pos = K - 1; pos >= 0; pos--)
        for (int q = 0; q < 2; q++) {
            const int32_t pc = q ? S.pM[pos] : S.pF[pos];
            if (pc >= 0) S.first_use[pc] = (int32_t) (pos / kPanel);
        }
    if (S.has_sub) {
        const int64_t nsub = S.Kpad / kSub;
        S.prow8.assign((size_t) nsub * 2 * kSub, -1);
        S.pcnt8.assign((size_t) nsub, 0);
        S.slots8.assign((size_t) S.Kpad, kZeroSlot8 | (kZeroSlot8 << 8));
        std::fill(unit_stamp.begin(), unit_stamp.end(), -1);
        for (int64_t pos = 0; pos < K; pos++) {
            const int32_t sub = (int32_t) (pos / kSub);
            int32_t sl[2] = { kZeroSlot8, kZeroSlot8 };
            for (int q = 0; q < 2; q++) {
                const int32_t pc = q ? S.pM[pos] : S.pF[pos];
                if (pc < 0) continue;
                if (unit_stamp[pc] != sub) {
                    unit_stamp[pc] = sub; S.sub_rows++;
                    slot_of[pc] = S.pcnt8[sub]++;
                    S.prow8[(size_t) sub * 2 * kSub + slot_of[pc]] = pc;
                }
                sl[q] = slot_of[pc];
            }
            S.slots8[pos] = sl[0] | (sl[1] << 8);
        }
    }

    // ---- units of 16 classes: staged positions of their parent rows --------------------------------------
    if (S.has_units) {
        const int zero_row = (int) S.Kpad_prev;
        const int nu = 2 * S.np;
        S.urow.assign((size_t) nu * kUnitRowsMax, -1); S.ucnt.assign((size_t) nu, 0); S.uslots.assign((size_t) S.Kpad, 0);
        std::vector<int32_t> ustamp((size_t) S.Kpad_prev + 1, -1), uslot((size_t) S.Kpad_prev + 1, 0);
        for (int u = 0; u < nu; u++) {
            // distinct parent rows of the unit's 16 classes (an unknown parent = the matrix's all-zero row), numbered in
            // order of first use; gmask[4] = the rows each group of 8 reads (classes 0-7 / 8-15, father / mother)
            int n = 0;
            int32_t rows[34], rf[16], rm[16];
            uint64_t gmask[4] = { 0, 0, 0, 0 };
            for (int i = 0; i < 16; i++)
                for (int q = 0; q < 2; q++) {
                    int32_t r = q ? S.pM[u * 16 + i] : S.pF[u * 16 + i];
                    if (r < 0) r = zero_row;
                    if (ustamp[r] != u) { ustamp[r] = u; uslot[r] = n; rows[n++] = r; }
                    (q ? rm : rf)[i] = uslot[r];          // index into rows[]
                    gmask[(i >> 3) * 2 + q] |= (uint64_t) 1 << uslot[r];
                }
            // Staged position of every row.  The consumer reads, per 128-bit shared load, the father (or mother) rows of
            // 8 consecutive classes at the same column: conflict-free when those 8 staged positions are distinct mod 8
            // (rows are an odd number of 16-byte units apart).  Greedy colouring with 8 colours over the four groups,
            // colours balanced so that position = colour + 8 * k fits the stage.
            const int Wd = n * (512 + 2) * 8 <= kUnitStageBytes ? 512 : 256;
            const int cap = kUnitStageBytes / ((Wd + 2) * 8);
            int load[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, pos[34];
            unsigned used[4] = { 0, 0, 0, 0 };            // colours already taken inside each group
            int npos = 0;
            for (int r = 0; r < n; r++) {
                unsigned forbidden = 0;
                for (int g4i = 0; g4i < 4; g4i++) if (gmask[g4i] >> r & 1) forbidden |= used[g4i];
                int best = -1;
                for (int pass = 0; pass < 2 && best < 0; pass++)
                    for (int c = 0; c < 8; c++) {
                        if (c + 8 * load[c] >= cap) continue;                  // no free position of this colour
                        if (pass == 0 && (forbidden >> c & 1u)) continue;
                        if (best < 0 || load[c] < load[best]) best = c;
                    }
                pos[r] = best + 8 * load[best]; load[best]++;
                for (int g4i = 0; g4i < 4; g4i++) if (gmask[g4i] >> r & 1) used[g4i] |= 1u << best;
                S.urow[(size_t) u * kUnitRowsMax + pos[r]] = rows[r];          // (unused positions stay -1: not copied)
                npos = std::max(npos, pos[r] + 1);
            }
            for (int i = 0; i < 16; i++) S.uslots[u * 16 + i] = pos[rf[i]] | (pos[rm[i]] << 8);
            S.ucnt[u] = npos | (n << 8) | ((Wd == 512 ? 1 : 0) << 16);
        }
    }

    // ---- corrections: class pairs that share a parent INDIVIDUAL whose class has several members ----
    // uses = (parent individual, class, multiplicity), grouped by the parent's index in the previous cut (counting sort:
    // the classes of a group stay in ascending order)
    struct Use { int32_t a, cls, mult; };
    std::vector<Use> raw;
    for (int64_t pos = 0; pos < K; pos++) {
        const ClassRec &c = C.cls[nat_at[pos]];
        bool pushed0 = false;
        for (int q = 0; q < 2; q++) {
            const int32_t pi = q ? c.pi1 : c.pi0, pcn = q ? c.pc1 : c.pc0;
            if (pi < 0 || Cp->cls[pcn].size <= 1) continue;       // singleton parent class: diagonal already = self kinship
            if (q == 1 && pushed0 && pi == c.pi0) { raw.back().mult++; continue; }   // kept: the same individual twice
            raw.push_back({ Cp->midx(pi), (int32_t) pos, 1 });
            if (q == 0) pushed0 = true;
        }
    }
    if (!raw.empty()) {
        std::vector<int32_t> off((size_t) Cp->nm + 1, 0);
        for (const Use &u : raw) off[u.a + 1]++;
        for (int64_t a = 0; a < Cp->nm; a++) off[a + 1] += off[a];
        std::vector<Use> uses(raw.size());
        {
            std::vector<int32_t> cur(off.begin(), off.end() - 1);
            for (const Use &u : raw) uses[cur[u.a]++] = u;
        }
        struct Term { int32_t i, j, c, mult; };
        std::vector<Term> terms;
        int64_t budget = (int64_t) 1 << 27;
        for (int64_t a0 = 0; a0 < Cp->nm; a0++) {
            const size_t a = off[a0], b = off[a0 + 1];
            if (a == b) continue;
            const int64_t k = (int64_t) (b - a);
            budget -= k * (k + 1) / 2;
            if (budget < 0) { over_budget_ = true; return; }
            const int32_t pc = Cp->pos[Cp->nat_of[a0]];
            for (size_t x = a; x < b; x++)
                for (size_t y = a; y <= x; y++) {
                    if (x == y && (S.flags[uses[x].cls] & kFlagSingleton)) continue;   // diagonal of a singleton := d_new
                    terms.push_back({ uses[x].cls, uses[y].cls, pc, uses[x].mult * uses[y].mult });
                }
        }
        // order by (i, j, c): counting sort by i, then the few terms of a row by insertion
        std::vector<int32_t> ioff((size_t) K + 1, 0);
        for (const Term &t : terms) ioff[t.i + 1]++;
        for (int64_t i = 0; i < K; i++) ioff[i + 1] += ioff[i];
        std::vector<Term> st(terms.size());
        {
            std::vector<int32_t> cur(ioff.begin(), ioff.end() - 1);
            for (const Term &t : terms) st[cur[t.i]++] = t;
        }
        for (int64_t i = 0; i < K; i++) {
            const int32_t a = ioff[i], b = ioff[i + 1];
            if (b - a > 64) {
                std::sort(st.begin() + a, st.begin() + b, [](const Term &x, const Term &y) { return x.j != y.j ? x.j < y.j : x.c < y.c; });
                continue;
            }
            for (int32_t x = a + 1; x < b; x++) {
                const Term t = st[x];
                int32_t y = x - 1;
                while (y >= a && (st[y].j > t.j || (st[y].j == t.j && st[y].c > t.c))) { st[y + 1] = st[y]; y--; }
                st[y + 1] = t;
            }
        }
        S.pg_off.push_back(0);
        for (size_t a = 0; a < st.size();) {
            size_t b = a;
            while (b < st.size() && st[b].i == st[a].i && st[b].j == st[a].j) {
                if (!S.pt_cls.empty() && (size_t) S.pg_off.back() < S.pt_cls.size() && S.pt_cls.back() == st[b].c)
                    S.pt_mult.back() += st[b].mult;
                else { S.pt_cls.push_back(st[b].c); S.pt_mult.push_back(st[b].mult); }
                b++;
            }
            S.pg_i.push_back(st[a].i); S.pg_j.push_back(st[a].j);
            S.pg_off.push_back((int32_t) S.pt_cls.size());
            a = b;
        }
    } else S.pg_off.push_back(0);
}

inline void PhiPlanner::start(std::function<void(int)> after_cut) {
    after_cut_ = std::move(after_cut);
    const int G = G_;
    state_.reset(new std::atomic<int>[G]);
    for (int g = 0; g < G; g++) state_[g].store(0, std::memory_order_relaxed);
    next_.store(0);
    auto wait_state = [this](int g, int want) {          // the cut is being planned by a thread that claimed it earlier
        int spins = 0;
        for (;;) {
            const int s = state_[g].load(std::memory_order_acquire);
            if (s < 0) return false;
            if (s >= want) return true;
            if (++spins < 64) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    };
    auto worker = [this, G, wait_state]() {
        for (int g = next_.fetch_add(1); g < G; g = next_.fetch_add(1)) {
            if (failed_any_.load()) return;
            stage_b(g);
            state_[g].store(1, std::memory_order_release);
            if (g > 0 && !wait_state(g - 1, 1)) return;
            stage_c(g);
            if (failed_.load() || over_budget_.load()) { fail_all(); return; }
            if (after_cut_) after_cut_(g);
            state_[g].store(2, std::memory_order_release);
        }
    };
    const int nt = std::max(1, std::min(threads_, G));
    if (nt == 1) { worker(); return; }
    for (int t = 0; t < nt; t++) pool_.emplace_back(worker);
}

inline bool PhiPlanner::wait_cut(int g) {
    int spins = 0;
    for (;;) {
        const int s = state_[g].load(std::memory_order_acquire);
        if (s < 0) return false;
        if (s >= 2) return true;
        if (++spins < 64) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
}

inline bool PhiPlanner::finish() {
    join();
    mark("stages B, C");
    if (over_budget_) {
        // an individual of a multi-member sibship with tens of thousands of mates: the "same individual" correction list would
        // be quadratic in the mates.  Code 3: the caller plans again with one matrix row per individual (no corrections at all).
        P.error = "the sibship compression needs too many corrections for this pedigree (an individual of a multi-member sibship "
                  "has tens of thousands of mates); plan with compress=0";
        P.error_code = 3; return false;
    }
    if (failed_) {
        P.error = "internal: a parent is missing from the previous cut";
        P.error_code = 1; return false;
    }
    P.plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
    return true;
}

inline bool build_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m,
                           const int32_t *pro, int compress_opt, int order_opt, int world, PhiPlan &P) {
    PhiPlanner planner(n, sire, dam, m, pro, compress_opt, order_opt, world, P);
    if (!planner.begin()) return false;
    planner.start();
    return planner.finish();
}

}  // namespace gk
