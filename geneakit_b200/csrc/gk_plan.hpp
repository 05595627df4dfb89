// Host planner of the B200 kinship path (pure C++17, no CUDA).
//
// Turns a rank-space pedigree + proband list into the per-generation TILE PLAN the sm_100a
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
//   * order: classes are stored sorted by the class index of their first known parent in the
//     previous matrix, so that the parents of T consecutive classes form a contiguous WINDOW
//     of T columns (coalesced / bulk-copied) plus a LIST of scattered rows (the other parent).
//   * tiles: per tile of T classes one fixed-size DESCRIPTOR (window start, list of scattered
//     parent rows, and the few other tiles that share a parent INDIVIDUAL with it: only those
//     tile pairs need the "same individual -> self kinship" rule of lib/compute.cpp:522-528),
//     and per class the local indices + individual ids of its two parents.
//
// Rounding: for G-1 <= 26 cuts below the founders every value is a dyadic rational that FP64
// holds exactly, so any evaluation order is bit-identical to the reference.  Deeper pedigrees
// switch to RANKED mode (no compression, per-pair choice of the summation grouping by rank,
// exactly as lib/compute.cpp:529-548 does).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

namespace gk {

constexpr int kMissing = 0xFFFF;   // la half-word for "parent unknown"
constexpr int kConfMax = 12;       // conflicting tiles stored inline per descriptor
constexpr int kTile = 32;          // classes per tile (at most)
constexpr int kWin = 32;           // window width (columns of the previous matrix)
constexpr int kListCap = 40;       // scattered parent rows per tile (a tile is closed early rather than exceed it)
// Tile descriptor, int32 words (16-byte aligned blocks, fetched with cp.async):
//   [0] c0  [1] nl | count << 16  [2] nconf (-1: always apply the identity rule)  [3] first class
//   [4, 4+cap)  list   [.., +kConfMax) conflicting tiles   [.., +T) la   [.., +2T) pind   [.., +T) rank
constexpr int kDescList = 4;
constexpr int kDescConf = kDescList + kListCap;
constexpr int kDescLa = kDescConf + kConfMax;
constexpr int kDescPind = kDescLa + kTile;
constexpr int kDescRank = kDescPind + 2 * kTile;
constexpr int kDescWords = kDescRank + kTile;      // 184 words = 736 bytes
static_assert(kDescWords % 4 == 0, "descriptor must be a multiple of 16 bytes");

struct StepPlan {
    int64_t K = 0, ld = 0, K_prev = 0;
    int nt = 0, nr_max = 0;
    int64_t cut_size = 0;     // |cut_g| as the reference counts it (last cut: m, duplicates included)
    int64_t n_kept = 0;       // members also in the previous cut
    int64_t P_ref = 0;        // distinct individuals of cut g-1 that are a parent of a NEW member
    int64_t P_cls = 0;        // distinct rows of matrix g-1 read by this step
    int64_t kept_cls = 0;     // singleton classes of kept individuals
    std::vector<int32_t> cls_rank;  // K (ranked mode only; host copy, the device reads the descriptors)
    std::vector<int32_t> desc;      // nt * kDescWords
};

struct PhiPlan {
    int32_t n = 0, m = 0, G = 0;
    int tile = kTile, win = kWin, lmax = kListCap;
    bool compressed = true, ranked = false, single_cut = false;
    std::vector<StepPlan> steps;            // steps[g], g = 0 .. G-1 (steps[0]: K only)
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
    int32_t pi0, pi1;   // parent individuals (window parent first), -1 missing; kept: both = ind
    int32_t pc0, pc1;   // their class rows in the previous matrix, -1 missing
    uint8_t kept;
};

struct ClassRec {
    int32_t pi0, pi1, pc0, pc1;
    int32_t first_member;   // lowest individual rank (tie-break; rank in ranked mode)
    uint8_t section;        // 0 = some member is the window parent of a class of the next cut
    uint8_t kept;
    int32_t member_begin, member_end;   // range in the sorted member array
};

inline bool build_phi_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m,
                           const int32_t *pro, int compress_opt, int tile_opt, PhiPlan &P) {
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
    (void) tile_opt;                   // one tile shape: 32 classes, 32-column window, <= 40 list rows
    const int T = kTile, W = kWin;
    P.steps.resize(G);

    // members entering each cut: enter(x) = G-1-maxp, last cut containing x = G-1-minp
    std::vector<int64_t> new_off(G + 1, 0);
    for (int32_t x = 0; x < n; x++) if (maxp[x] >= 0) new_off[G - 1 - maxp[x] + 1]++;
    for (int g = 0; g < G; g++) new_off[g + 1] += new_off[g];
    std::vector<int32_t> new_mem(new_off[G]);
    {
        std::vector<int64_t> w(new_off.begin(), new_off.end() - 1);
        for (int32_t x = 0; x < n; x++) if (maxp[x] >= 0) new_mem[w[G - 1 - maxp[x]]++] = x;
    }
    auto leave = [&](int32_t x) { return G - 1 - minp[x]; };

    std::vector<int32_t> cls_prev(n, -1), cls_cur(n, -1);
    std::vector<int32_t> win_stamp(n, -1);            // step stamp: individual is a window parent for the next cut
    std::vector<int64_t> use_stamp(n, -1);            // (step << 32 | tile): individual already recorded for this tile
    std::vector<std::pair<int32_t, int32_t>> use_pairs, conf_pairs;
    std::vector<int32_t> par_stamp(n, -1);            // distinct-parent counting
    std::vector<int32_t> prow_stamp;                  // distinct parent rows of previous matrix
    std::vector<int32_t> row_tile_stamp, row_loc;     // per previous-matrix row: local list slot in the current tile

    std::vector<int32_t> cur;                         // members of the current cut (individuals)
    std::vector<MemberRec> mem;
    std::vector<ClassRec> cls;
    std::vector<int32_t> order;

    for (int g = 0; g < G; g++) {
        StepPlan &S = P.steps[g];
        // ---- members of cut g ------------------------------------------------------------
        std::vector<int32_t> nxt;
        nxt.reserve(cur.size() + (new_off[g + 1] - new_off[g]));
        mem.clear();
        for (int32_t x : cur) if (leave(x) >= g) {
            MemberRec r; r.ind = x; r.kept = 1; r.pi0 = r.pi1 = x; r.pc0 = r.pc1 = cls_prev[x];
            mem.push_back(r); nxt.push_back(x);
        }
        S.n_kept = (int64_t) mem.size();
        int64_t pref = 0;
        for (int64_t e = new_off[g]; e < new_off[g + 1]; e++) {
            int32_t x = new_mem[e];
            MemberRec r; r.ind = x; r.kept = 0;
            int32_t f = sire[x], d = dam[x];
            if (f >= 0) { r.pi0 = f; r.pi1 = d; } else { r.pi0 = d; r.pi1 = -1; }
            r.pc0 = r.pi0 >= 0 ? cls_prev[r.pi0] : -1;
            r.pc1 = r.pi1 >= 0 ? cls_prev[r.pi1] : -1;
            if ((r.pi0 >= 0 && r.pc0 < 0) || (r.pi1 >= 0 && r.pc1 < 0)) {
                P.error = "internal: parent of a new cut member is not in the previous cut"; P.error_code = 1; return false;
            }
            for (int q = 0; q < 2; q++) {
                int32_t p = q ? r.pi1 : r.pi0;
                if (p >= 0 && par_stamp[p] != g) { par_stamp[p] = g; pref++; }
            }
            mem.push_back(r); nxt.push_back(x);
        }
        S.P_ref = pref;
        S.cut_size = (g == G - 1) ? (int64_t) m : (int64_t) mem.size();
        S.K_prev = g > 0 ? P.steps[g - 1].K : 0;

        // ---- window parents of the next cut (section flag) -------------------------------
        if (g + 1 < G) {
            for (const MemberRec &r : mem) if (leave(r.ind) >= g + 1) win_stamp[r.ind] = g;   // kept next: own row
            for (int64_t e = new_off[g + 1]; e < new_off[g + 2]; e++) {
                int32_t y = new_mem[e];
                int32_t w = sire[y] >= 0 ? sire[y] : dam[y];
                if (w >= 0) win_stamp[w] = g;
            }
        }

        // ---- classes: new members with equal (pi0, pi1) share a row; kept are singletons ----
        auto key_of = [&](const MemberRec &r) -> uint64_t {
            if (r.kept || !P.compressed)           // unique per individual
                return (uint64_t(1) << 63) | uint64_t(uint32_t(r.ind));
            return (uint64_t(uint32_t(r.pi0 + 1)) << 31) | uint64_t(uint32_t(r.pi1 + 1));
        };
        std::sort(mem.begin(), mem.end(), [&](const MemberRec &a, const MemberRec &b) {
            uint64_t ka = key_of(a), kb = key_of(b);
            if (ka != kb) return ka < kb;
            return a.ind < b.ind;
        });
        cls.clear();
        for (int32_t i = 0; i < (int32_t) mem.size();) {
            int32_t j = i + 1;
            uint64_t k = key_of(mem[i]);
            while (j < (int32_t) mem.size() && key_of(mem[j]) == k) j++;
            ClassRec c;
            c.pi0 = mem[i].pi0; c.pi1 = mem[i].pi1; c.pc0 = mem[i].pc0; c.pc1 = mem[i].pc1;
            c.first_member = mem[i].ind; c.kept = mem[i].kept; c.member_begin = i; c.member_end = j;
            c.section = 1;
            for (int32_t q = i; q < j; q++) if (win_stamp[mem[q].ind] == g) { c.section = 0; break; }
            cls.push_back(c);
            i = j;
        }
        const int64_t K = (int64_t) cls.size();
        order.resize(K);
        for (int64_t i = 0; i < K; i++) order[i] = (int32_t) i;
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            const ClassRec &x = cls[a], &y = cls[b];
            if (x.section != y.section) return x.section < y.section;
            if (x.pc0 != y.pc0) return x.pc0 < y.pc0;
            if (x.pc1 != y.pc1) return x.pc1 < y.pc1;
            return x.first_member < y.first_member;
        });
        S.K = K;
        S.ld = round_up(std::max<int64_t>(K, 1), 16);
        for (int64_t pos = 0; pos < K; pos++) {
            const ClassRec &c = cls[order[pos]];
            for (int32_t q = c.member_begin; q < c.member_end; q++) cls_cur[mem[q].ind] = (int32_t) pos;
        }
        if (P.ranked) {
            S.cls_rank.resize(K);
            for (int64_t pos = 0; pos < K; pos++) S.cls_rank[pos] = cls[order[pos]].first_member;
        }

        // ---- tiles (only steps that read a previous matrix) -------------------------------
        if (g > 0) {
            const int DW = kDescWords;
            S.desc.clear();
            prow_stamp.assign((size_t) S.K_prev, -1);
            row_tile_stamp.assign((size_t) S.K_prev, -1);
            row_loc.assign((size_t) S.K_prev, 0);
            use_pairs.clear();
            int64_t pcls = 0, keptc = 0;
            int t = -1, nl = 0, cnt = 0;
            int32_t c0 = 0;
            uint8_t tile_section = 0;
            auto open_tile = [&](int64_t pos) {
                t++;
                S.desc.resize((size_t) (t + 1) * DW, 0);
                int32_t *d = &S.desc[(size_t) t * DW];
                for (int k = 0; k < T; k++) { d[kDescLa + k] = kMissing | (kMissing << 16); d[kDescPind + 2 * k] = d[kDescPind + 2 * k + 1] = -1; }
                // window start: first known father class from here on inside this section (sorted => the minimum)
                int32_t cmin = -1;
                for (int64_t q = pos; q < K && q < pos + T; q++) {
                    const ClassRec &c = cls[order[q]];
                    if (c.section != cls[order[pos]].section) break;
                    if (c.pc0 >= 0) { cmin = c.pc0; break; }
                }
                c0 = cmin < 0 ? 0 : (cmin & ~1);                 // even: 16-byte aligned row segments
                nl = 0; cnt = 0; tile_section = cls[order[pos]].section;
                d[0] = c0; d[3] = (int32_t) pos;
            };
            auto close_tile = [&]() {
                int32_t *d = &S.desc[(size_t) t * DW];
                d[1] = nl | (cnt << 16);
                S.nr_max = std::max(S.nr_max, W + nl);
            };
            for (int64_t pos = 0; pos < K; pos++) {
                const ClassRec &c = cls[order[pos]];
                // how many new list rows would this class add to the open tile?
                auto extra = [&]() {
                    int e = 0;
                    int32_t seen = -1;
                    for (int q = 0; q < 2; q++) {
                        const int32_t pi = q ? c.pi1 : c.pi0, pc = q ? c.pc1 : c.pc0;
                        if (pi < 0 || (pc >= c0 && pc < c0 + W) || row_tile_stamp[pc] == t || pc == seen) continue;
                        seen = pc; e++;
                    }
                    return e;
                };
                if (t < 0 || cnt == T || c.section != tile_section || nl + extra() > kListCap) {
                    if (t >= 0) close_tile();
                    open_tile(pos);
                }
                int32_t *d = &S.desc[(size_t) t * DW];
                if (c.kept) keptc++;
                int la[2] = { kMissing, kMissing };
                for (int q = 0; q < 2; q++) {
                    const int32_t pi = q ? c.pi1 : c.pi0, pc = q ? c.pc1 : c.pc0;
                    d[kDescPind + 2 * cnt + q] = pi;
                    if (pi < 0) continue;
                    if (prow_stamp[pc] != g) { prow_stamp[pc] = g; pcls++; }
                    const int64_t us = ((int64_t) g << 32) | (uint32_t) t;
                    if (use_stamp[pi] != us) { use_stamp[pi] = us; use_pairs.emplace_back(pi, t); }
                    if (pc >= c0 && pc < c0 + W) { la[q] = pc - c0; continue; }
                    if (row_tile_stamp[pc] != t) {
                        row_tile_stamp[pc] = t; row_loc[pc] = W + nl;
                        d[kDescList + nl++] = pc;
                    }
                    la[q] = row_loc[pc];
                }
                d[kDescLa + cnt] = la[0] | (la[1] << 16);
                d[kDescRank + cnt] = P.ranked ? c.first_member : 0;
                cnt++;
            }
            if (t >= 0) close_tile();
            S.nt = t + 1;
            S.P_cls = pcls; S.kept_cls = keptc;
            // tiles that share a parent individual: only those pairs apply the identity rule
            std::sort(use_pairs.begin(), use_pairs.end());
            conf_pairs.clear();
            for (size_t i = 0; i < use_pairs.size();) {
                size_t j = i + 1;
                while (j < use_pairs.size() && use_pairs[j].first == use_pairs[i].first) j++;
                for (size_t a = i; a < j; a++)
                    for (size_t b = i; b < j; b++)
                        if (a != b) conf_pairs.emplace_back(use_pairs[a].second, use_pairs[b].second);
                i = j;
            }
            std::sort(conf_pairs.begin(), conf_pairs.end());
            conf_pairs.erase(std::unique(conf_pairs.begin(), conf_pairs.end()), conf_pairs.end());
            for (size_t i = 0; i < conf_pairs.size();) {
                size_t j = i;
                const int tt = conf_pairs[i].first;
                while (j < conf_pairs.size() && conf_pairs[j].first == tt) j++;
                int32_t *d = &S.desc[(size_t) tt * DW];
                if ((int) (j - i) > kConfMax) d[2] = -1;
                else {
                    d[2] = (int32_t) (j - i);
                    for (size_t q = i; q < j; q++) d[kDescConf + (q - i)] = conf_pairs[q].second;
                }
                i = j;
            }
        }
        cur.swap(nxt);
        cls_prev.swap(cls_cur);   // cls_cur now holds stale entries; every member is rewritten before use
    }

    // ---- final expansion map (caller order, duplicates expand to the same row) --------------
    P.final_cls.resize(m); P.final_ind.resize(m);
    for (int32_t i = 0; i < m; i++) { P.final_ind[i] = pro[i]; P.final_cls[i] = cls_prev[pro[i]]; }
    P.plan_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

}  // namespace gk
