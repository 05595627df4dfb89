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

struct MemberRec {
    int32_t ind;        // individual (rank)
    int32_t pi0, pi1;   // parent individuals, -1 missing; kept: both = ind
    int32_t pc0, pc1;   // their class rows in the previous matrix, -1 missing
    uint8_t kept;
};

struct ClassRec {
    int32_t pi0, pi1, pc0, pc1;
    int32_t first_member;   // lowest individual rank
    int32_t size;
    uint8_t kept;
    int32_t member_begin, member_end;   // range in the sorted member array
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
    // walk starts: classes whose parent rows have the fewest users first (path ends before path middles)
    std::vector<int32_t> starts(K), deg(K);
    for (int32_t i = 0; i < K; i++) {
        int32_t r[2]; int k = rows_of(cls[subset[i]], r);
        int32_t d = 0;
        for (int q = 0; q < k; q++) d += off[r[q] + 1] - off[r[q]];
        deg[i] = d; starts[i] = i;
    }
    std::stable_sort(starts.begin(), starts.end(), [&](int32_t a, int32_t b) { return deg[a] < deg[b]; });
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

// Run fn(g) for g in [g0, g1) on up to `threads` host threads (the cuts are planned independently).
template <class F>
inline void parallel_steps(int g0, int g1, int threads, F fn) {
    if (threads <= 1 || g1 - g0 <= 1) { for (int g = g0; g < g1; g++) fn(g); return; }
    std::atomic<int> next(g0);
    std::vector<std::thread> pool;
    const int nt = std::min(threads, g1 - g0);
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&]() { for (int g = next.fetch_add(1); g < g1; g = next.fetch_add(1)) fn(g); });
    for (auto &th : pool) th.join();
}

// Per-cut scratch of the planner.  Stage A fills members / classes (in NATURAL order: sorted by key), stage
// B the order of the classes, stage C the arrays the kernels read.
struct CutWork {
    std::vector<int32_t> members;                  // individuals of the cut, ascending
    std::vector<int32_t> nat_of;                   // natural class id of members[k]
    std::vector<ClassRec> cls;                     // natural order; pc0 / pc1 = NATURAL ids in the previous cut (stage B)
    std::vector<int32_t> pos;                      // natural id -> position in the class matrix
    int64_t n_kept = 0, P_ref = 0;
    int32_t base = 0;
    std::vector<int32_t> table;                    // table[ind - base] = natural id (built when the cut's rank range is compact)
    void build_table() {
        if (members.empty()) return;
        base = members.front();
        const int64_t range = (int64_t) members.back() - base + 1;
        if (range > 8 * (int64_t) members.size() + 1024) return;     // sparse cut: keep the binary search
        table.assign((size_t) range, -1);
        for (size_t k = 0; k < members.size(); k++) table[members[k] - base] = nat_of[k];
    }
    int32_t nat(int32_t ind) const {               // natural class id of an individual of this cut
        if (!table.empty()) return table[ind - base];
        const auto it = std::lower_bound(members.begin(), members.end(), ind);
        return nat_of[it - members.begin()];
    }
};

inline bool build_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m,
                           const int32_t *pro, int compress_opt, int order_opt, int world, PhiPlan &P) {
    auto t0 = std::chrono::steady_clock::now();
    P = PhiPlan();
    P.n = n; P.m = m;
    if (n <= 0 || m <= 0 || !sire || !dam || !pro) {
        P.error = "gk_phi: empty pedigree or proband list (the binding substitutes get_proband_ids for an empty list)";
        P.error_code = 1; return false;
    }
    for (int32_t i = 0; i < m; i++)
        if (pro[i] < 0 || pro[i] >= n) { P.error = "proband rank out of range"; P.error_code = 2; return false; }

    // ---- min / max ancestral path length from a proband (one reverse-topological sweep) ----
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
            if (p >= x) { P.error = "parent does not precede child in rank order"; P.error_code = 1; return false; }
            minp[p] = (minp[p] < 0) ? minp[x] + 1 : std::min(minp[p], minp[x] + 1);
            maxp[p] = std::max(maxp[p], maxp[x] + 1);
        }
    }
    const int G = top + 1;
    P.G = G;
    P.single_cut = (G == 1);
    P.ranked = (G - 1 > 26);
    P.compressed = P.ranked ? false : (compress_opt != 0);
    P.steps.resize(G);
    int threads = (int) std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, 24));
    if (const char *e = std::getenv("GK_PLAN_THREADS")) threads = std::max(1, std::atoi(e));

    const bool timing = std::getenv("GK_PLAN_TIMING") != nullptr;
    auto mark = [&](const char *what) {
        if (timing) std::fprintf(stderr, "[plan] %-28s %.1f ms\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    mark("path lengths");
    // ---- members of every cut (x belongs to cuts enter(x) = G-1-maxp .. leave(x) = G-1-minp), ascending ----
    std::vector<CutWork> W(G);
    {
        std::vector<int64_t> cnt(G + 1, 0);
        for (int32_t x = 0; x < n; x++) if (maxp[x] >= 0) { cnt[G - 1 - maxp[x]]++; cnt[G - minp[x]]--; }
        int64_t run = 0;
        for (int g = 0; g < G; g++) { run += cnt[g]; W[g].members.reserve((size_t) run); }
        for (int32_t x = 0; x < n; x++) if (maxp[x] >= 0)
            for (int g = G - 1 - maxp[x]; g <= G - 1 - minp[x]; g++) W[g].members.push_back(x);
    }
    mark("members");
    auto enter = [&](int32_t x) { return G - 1 - maxp[x]; };
    std::atomic<bool> failed(false), over_budget(false);
    // multi-GPU: lineage owner of every individual that takes part (one top-down sweep)
    std::vector<uint8_t> owner;
    if (world > 1) {
        owner.assign((size_t) n, 0);
        int32_t rr = 0;
        for (int32_t x = 0; x < n; x++) {
            if (maxp[x] < 0) continue;
            const int32_t f = sire[x], d = dam[x];
            if (f >= 0 && maxp[f] >= 0) owner[x] = owner[f];
            else if (d >= 0 && maxp[d] >= 0) owner[x] = owner[d];
            else owner[x] = (uint8_t) (rr++ % world);
        }
    }

    // ---- stage A (parallel): classes of every cut, in natural order --------------------------------
    parallel_steps(0, G, threads, [&](int g) {
        CutWork &C = W[g];
        const size_t nm = C.members.size();
        // new members with equal (father, mother) share a class; kept members and uncompressed jobs: singletons
        std::vector<std::pair<uint64_t, uint32_t>> keys(nm);
        for (size_t k = 0; k < nm; k++) {
            const int32_t x = C.members[k];
            const bool kept = enter(x) < g;
            uint64_t key;
            if (kept || !P.compressed) key = (uint64_t(1) << 63) | uint64_t(uint32_t(x));
            else key = (uint64_t(uint32_t(sire[x] + 1)) << 31) | uint64_t(uint32_t(dam[x] + 1));
            keys[k] = { key, (uint32_t) k };
            if (kept) C.n_kept++;
        }
        std::sort(keys.begin(), keys.end());          // ties: position = individual order
        C.nat_of.assign(nm, 0);
        std::vector<int32_t> parents;
        for (size_t i = 0; i < nm;) {
            size_t j = i + 1;
            while (j < nm && keys[j].first == keys[i].first) j++;
            const int32_t x = C.members[keys[i].second];
            ClassRec c;
            c.kept = enter(x) < g;
            c.pi0 = c.kept ? x : sire[x]; c.pi1 = c.kept ? x : dam[x];
            c.pc0 = c.pc1 = -1;
            c.first_member = x; c.size = (int32_t) (j - i);
            c.member_begin = (int32_t) i; c.member_end = (int32_t) j;
            for (size_t q = i; q < j; q++) C.nat_of[keys[q].second] = (int32_t) C.cls.size();
            if (!c.kept) { if (c.pi0 >= 0) parents.push_back(c.pi0); if (c.pi1 >= 0) parents.push_back(c.pi1); }
            C.cls.push_back(c);
            i = j;
        }
        std::sort(parents.begin(), parents.end());
        C.P_ref = (int64_t) (std::unique(parents.begin(), parents.end()) - parents.begin());
        C.build_table();
    });

    mark("stage A (classes)");
    // ---- stage B: parent classes (natural ids in the previous cut) and the order of the classes --------
    // The order only needs to know WHICH classes share a parent row, so the cuts are independent of one
    // another -- except for multi-GPU jobs, which group the classes by the rank that owns the first parent's
    // row (a POSITION in the previous matrix): those are ordered cut after cut.
    auto order_cut = [&](int g) {
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
            order.resize(K);
            for (int64_t i = 0; i < K; i++) order[i] = (int32_t) i;
            std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return C.cls[a].first_member > C.cls[b].first_member; });
        } else if (order_opt == 0 || g == 0) {
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
            if (world <= 1) {
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
                std::vector<std::vector<int32_t>> groups(world);
                for (int64_t i = 0; i < K; i++) {
                    const int32_t pi = C.cls[i].pi0 >= 0 ? C.cls[i].pi0 : C.cls[i].pi1;
                    const int q = pi >= 0 ? owner[pi] : owner[C.cls[i].first_member];
                    groups[q].push_back((int32_t) i);
                }
                for (int q = 0; q < world; q++) locality_order(C.cls, groups[q], K_prev, order);
            }
        }
        C.pos.assign(K, 0);
        for (int64_t p = 0; p < K; p++) C.pos[order[p]] = (int32_t) p;
    };
    parallel_steps(0, G, threads, order_cut);

    mark("stage B (order)");
    // ---- stage C (parallel): the arrays the kernels read -------------------------------------------
    parallel_steps(0, G, threads, [&](int g) {
        StepPlan &S = P.steps[g];
        const CutWork &C = W[g];
        const CutWork *Cp = g > 0 ? &W[g - 1] : nullptr;
        const int64_t K = (int64_t) C.cls.size();
        S.n_kept = C.n_kept; S.P_ref = C.P_ref;
        S.cut_size = (g == G - 1) ? (int64_t) m : (int64_t) C.members.size();
        S.K_prev = Cp ? (int64_t) Cp->cls.size() : 0;
        S.Kpad_prev = Cp ? round_up(std::max<int64_t>(S.K_prev, 1), kPanel) : 0;
        S.K = K;
        S.Kpad = round_up(std::max<int64_t>(K, 1), kPanel);
        S.np = (int) (S.Kpad / kPanel);
        S.pF.assign((size_t) S.Kpad, -1); S.pM.assign((size_t) S.Kpad, -1); S.flags.assign((size_t) S.Kpad, 0);
        S.cls_ind.resize(K); S.cls_size.resize(K);
        std::vector<int32_t> nat_at(K);                    // position -> natural id
        for (int64_t i = 0; i < K; i++) nat_at[C.pos[i]] = (int32_t) i;
        for (int64_t p = 0; p < K; p++) {
            const ClassRec &c = C.cls[nat_at[p]];
            if ((c.pi0 >= 0 && g > 0 && c.pc0 < 0) || (c.pi1 >= 0 && g > 0 && c.pc1 < 0)) failed = true;
            S.pF[p] = c.pc0 >= 0 ? Cp->pos[c.pc0] : -1;
            S.pM[p] = c.pc1 >= 0 ? Cp->pos[c.pc1] : -1;
            S.flags[p] = (c.size == 1 ? kFlagSingleton : 0) | (c.kept ? kFlagKept : 0) |
                         ((!c.kept && c.pc0 >= 0 && c.pc1 >= 0) ? kFlagBoth : 0);
            S.cls_ind[p] = c.first_member; S.cls_size[p] = c.size;
            if (c.kept) S.kept_cls++;
        }
        if (g == 0) return;

        // ---- distinct parent rows: overall, per panel (64 slots), per sub-panel of 8 classes (16 slots) ----
        std::vector<int32_t> stamp((size_t) S.K_prev, -1), unit_stamp((size_t) S.K_prev, -1), slot_of((size_t) S.K_prev, 0);
        S.prow.assign((size_t) S.np * 2 * kPanel, -1);
        S.pcnt.assign((size_t) S.np, 0);
        S.slots.assign((size_t) S.Kpad, kZeroSlot | (kZeroSlot << 8));
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
                    S.prow[(size_t) panel * 2 * kPanel + slot_of[pc]] = pc;
                }
                sl[q] = slot_of[pc];
            }
            S.slots[pos] = sl[0] | (sl[1] << 8);
        }
        S.P_cls = pcls; S.panel_rows = prow;
        // a row N^T[c] of panel I's transposed panel is only ever read by the tiles J <= I that hold a child of c
        S.first_use.assign((size_t) S.Kpad_prev, S.np);
        for (int64_t pos = K - 1; pos >= 0; pos--)
            for (int q = 0; q < 2; q++) {
                const int32_t pc = q ? S.pM[pos] : S.pF[pos];
                if (pc >= 0) S.first_use[pc] = (int32_t) (pos / kPanel);
            }
        {
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
        {
        const int zero_row = (int) S.Kpad_prev;
        const int nu = 2 * S.np;
        S.urow.assign((size_t) nu * kUnitRowsMax, -1); S.ucnt.assign((size_t) nu, 0); S.uslots.assign((size_t) S.Kpad, 0);
        std::vector<int32_t> ustamp((size_t) S.Kpad_prev + 1, -1), uslot((size_t) S.Kpad_prev + 1, 0);
        for (int u = 0; u < nu; u++) {
            // distinct parent rows of the unit's 16 classes (an unknown parent = the matrix's all-zero row)
            int n = 0;
            int32_t rows[34], rf[16], rm[16];
            for (int i = 0; i < 16; i++)
                for (int q = 0; q < 2; q++) {
                    int32_t r = q ? S.pM[u * 16 + i] : S.pF[u * 16 + i];
                    if (r < 0) r = zero_row;
                    if (ustamp[r] != u) { ustamp[r] = u; uslot[r] = n; rows[n++] = r; }
                    (q ? rm : rf)[i] = uslot[r];          // index into rows[]
                }
            // Staged position of every row.  The consumer reads, per 128-bit shared load, the father (or mother) rows of
            // 8 consecutive classes at the same column: conflict-free when those 8 staged positions are distinct mod 8
            // (rows are an odd number of 16-byte units apart).  Greedy colouring with 8 colours over the four groups
            // (classes 0-7 / 8-15, father / mother), colours balanced so that position = colour + 8 * k fits the stage.
            const int W = n * (512 + 2) * 8 <= kUnitStageBytes ? 512 : 256;
            const int cap = kUnitStageBytes / ((W + 2) * 8);
            int load[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, colour[34], pos[34];
            for (int r = 0; r < n; r++) {
                unsigned forbidden = 0;
                for (int g4 = 0; g4 < 4; g4++) {
                    const int32_t *par = (g4 & 1) ? rm : rf;
                    const int i0 = (g4 >> 1) * 8;
                    bool member = false;
                    for (int i = i0; i < i0 + 8; i++) if (par[i] == r) member = true;
                    if (!member) continue;
                    for (int i = i0; i < i0 + 8; i++) if (par[i] < r) forbidden |= 1u << colour[par[i]];
                }
                int best = -1;
                for (int pass = 0; pass < 2 && best < 0; pass++)
                    for (int c = 0; c < 8; c++) {
                        if (c + 8 * load[c] >= cap) continue;                  // no free position of this colour
                        if (pass == 0 && (forbidden >> c & 1u)) continue;
                        if (best < 0 || load[c] < load[best]) best = c;
                    }
                colour[r] = best; pos[r] = best + 8 * load[best]; load[best]++;
            }
            int npos = 0;
            for (int r = 0; r < n; r++) { S.urow[(size_t) u * kUnitRowsMax + pos[r]] = rows[r]; npos = std::max(npos, pos[r] + 1); }
            for (int q = 0; q < npos; q++) {                                     // unused positions: -1 (not copied)
                bool used = false;
                for (int r = 0; r < n; r++) if (pos[r] == q) used = true;
                if (!used) S.urow[(size_t) u * kUnitRowsMax + q] = -1;
            }
            for (int i = 0; i < 16; i++) S.uslots[u * 16 + i] = pos[rf[i]] | (pos[rm[i]] << 8);
            S.ucnt[u] = npos | (n << 8) | ((W == 512 ? 1 : 0) << 16);
        }
        }

        // ---- corrections: class pairs that share a parent INDIVIDUAL whose class has several members ----
        struct Use { int32_t ind, cls, mult; };
        struct Term { int32_t i, j, c, mult; };
        std::vector<Use> uses;
        std::vector<Term> terms;
        for (int64_t pos = 0; pos < K; pos++) {
            const ClassRec &c = C.cls[nat_at[pos]];
            bool pushed0 = false;
            for (int q = 0; q < 2; q++) {
                const int32_t pi = q ? c.pi1 : c.pi0, pcn = q ? c.pc1 : c.pc0;
                if (pi < 0 || Cp->cls[pcn].size <= 1) continue;       // singleton parent class: diagonal already = self kinship
                if (q == 1 && pushed0 && pi == c.pi0) { uses.back().mult++; continue; }   // kept: the same individual twice
                uses.push_back({ pi, (int32_t) pos, 1 });
                if (q == 0) pushed0 = true;
            }
        }
        std::sort(uses.begin(), uses.end(), [](const Use &a, const Use &b) {
            return ((uint64_t) (uint32_t) a.ind << 32 | (uint32_t) a.cls) < ((uint64_t) (uint32_t) b.ind << 32 | (uint32_t) b.cls);
        });
        int64_t budget = (int64_t) 1 << 27;
        for (size_t a = 0; a < uses.size();) {
            size_t b = a + 1;
            while (b < uses.size() && uses[b].ind == uses[a].ind) b++;
            const int64_t k = (int64_t) (b - a);
            budget -= k * (k + 1) / 2;
            if (budget < 0) { over_budget = true; return; }
            const int32_t pc = Cp->pos[Cp->nat(uses[a].ind)];
            for (size_t x = a; x < b; x++)
                for (size_t y = a; y <= x; y++) {
                    if (x == y && (S.flags[uses[x].cls] & kFlagSingleton)) continue;   // diagonal of a singleton := d_new
                    terms.push_back({ uses[x].cls, uses[y].cls, pc, uses[x].mult * uses[y].mult });
                }
            a = b;
        }
        std::sort(terms.begin(), terms.end(), [](const Term &a, const Term &b) {
            if (a.i != b.i) return a.i < b.i;
            if (a.j != b.j) return a.j < b.j;
            return a.c < b.c;
        });
        S.pg_off.push_back(0);
        for (size_t a = 0; a < terms.size();) {
            size_t b = a;
            while (b < terms.size() && terms[b].i == terms[a].i && terms[b].j == terms[a].j) {
                if (!S.pt_cls.empty() && (size_t) S.pg_off.back() < S.pt_cls.size() && S.pt_cls.back() == terms[b].c)
                    S.pt_mult.back() += terms[b].mult;
                else { S.pt_cls.push_back(terms[b].c); S.pt_mult.push_back(terms[b].mult); }
                b++;
            }
            S.pg_i.push_back(terms[a].i); S.pg_j.push_back(terms[a].j);
            S.pg_off.push_back((int32_t) S.pt_cls.size());
            a = b;
        }
    });
    mark("stage C (arrays)");
    if (over_budget) {
        // an individual of a multi-member sibship with tens of thousands of mates: the "same individual" correction list would
        // be quadratic in the mates.  Code 3: the caller plans again with one matrix row per individual (no corrections at all).
        P.error = "the sibship compression needs too many corrections for this pedigree (an individual of a multi-member sibship "
                  "has tens of thousands of mates); plan with compress=0";
        P.error_code = 3; return false;
    }
    if (failed) {
        P.error = "internal: a parent is missing from the previous cut";
        P.error_code = 1; return false;
    }

    // ---- final expansion map (caller order, duplicates expand to the same row) --------------
    P.final_cls.resize(m); P.final_ind.resize(m);
    const CutWork &L = W[G - 1];
    for (int32_t i = 0; i < m; i++) { P.final_ind[i] = pro[i]; P.final_cls[i] = L.pos[L.nat(pro[i])]; }
    P.plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

}  // namespace gk
