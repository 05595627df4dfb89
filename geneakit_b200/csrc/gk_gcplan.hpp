// Host planner of the gen.gc propagation (pure C++17).
//
// The reference (lib/compute.cpp:817-862) enumerates, per ancestor, every descending path by
// DFS and stops at the first proband on each path: exponential under inbreeding.  Its RESULT is
// the linear top-down propagation  v(anc)=1, v(i) = 1/2 pass(father) + 1/2 pass(mother),
// pass(x) = 0 if x is a proband else v(x)  (SURVEY 8a row a11), which is the same row
// gather-average as the kinship recurrence.  We run it level by level over the same vertex cuts
// as gen.phi, with one matrix row per live individual and one column per requested ancestor.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace gk {

struct GcStep {
    int64_t rows = 0;
    std::vector<int32_t> p0, p1;        // parent row in the previous matrix, -1 none; kept: (old row, -2)
    std::vector<int32_t> seed_off, seed_col;
};

struct GcPlan {
    int32_t n = 0, m = 0, a = 0, G = 0;
    int64_t ld = 0, max_rows = 0;
    std::vector<GcStep> steps;
    std::vector<int32_t> row_of;        // m: row of caller proband i in the last matrix, -1 -> zeros
    std::string error;
    int error_code = 0;
};

inline bool build_gc_plan(int32_t n, const int32_t *sire, const int32_t *dam, int32_t m, const int32_t *pro,
                          int32_t a, const int32_t *anc, GcPlan &P) {
    P = GcPlan();
    P.n = n; P.m = m; P.a = a;
    if (n <= 0 || m <= 0 || a <= 0 || !sire || !dam || !pro || !anc) {
        P.error = "gk_gc: empty pedigree, proband or ancestor list"; P.error_code = 1; return false;
    }
    for (int32_t i = 0; i < m; i++) if (pro[i] < 0 || pro[i] >= n) { P.error = "proband rank out of range"; P.error_code = 2; return false; }
    for (int32_t j = 0; j < a; j++) if (anc[j] < 0 || anc[j] >= n) { P.error = "ancestor rank out of range"; P.error_code = 2; return false; }
    std::vector<int32_t> minp(n, -1), maxp(n, -1);
    std::vector<uint8_t> is_pro(n, 0);
    for (int32_t i = 0; i < m; i++) { minp[pro[i]] = 0; maxp[pro[i]] = std::max(maxp[pro[i]], 0); is_pro[pro[i]] = 1; }
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
    P.ld = (a + 15) / 16 * 16;
    // ancestor columns per individual (CSR)
    std::vector<int32_t> aoff(n + 1, 0), acol(a);
    for (int32_t j = 0; j < a; j++) aoff[anc[j] + 1]++;
    for (int32_t x = 0; x < n; x++) aoff[x + 1] += aoff[x];
    {
        std::vector<int32_t> w(aoff.begin(), aoff.end() - 1);
        for (int32_t j = 0; j < a; j++) acol[w[anc[j]]++] = j;
    }
    std::vector<std::vector<int32_t>> new_at(G);
    for (int32_t x = 0; x < n; x++) if (maxp[x] >= 0) new_at[G - 1 - maxp[x]].push_back(x);
    std::vector<int32_t> row_prev(n, -1), row_cur(n, -1), cur;
    P.steps.resize(G);
    for (int g = 0; g < G; g++) {
        GcStep &S = P.steps[g];
        std::vector<int32_t> nxt;
        S.seed_off.push_back(0);
        auto add_row = [&](int32_t x, int32_t p0, int32_t p1, bool seed) {
            row_cur[x] = (int32_t) nxt.size();
            nxt.push_back(x);
            S.p0.push_back(p0); S.p1.push_back(p1);
            if (seed) for (int32_t k = aoff[x]; k < aoff[x + 1]; k++) S.seed_col.push_back(acol[k]);
            S.seed_off.push_back((int32_t) S.seed_col.size());
        };
        for (int32_t x : cur) if (G - 1 - minp[x] >= g) add_row(x, row_prev[x], -2, false);   // kept: seed travels with the copy
        for (int32_t x : new_at[g]) {
            const int32_t f = sire[x], d = dam[x];
            const int32_t p0 = (f >= 0 && !is_pro[f]) ? row_prev[f] : -1;   // pass(x) = 0 for probands (:820-821)
            const int32_t p1 = (d >= 0 && !is_pro[d]) ? row_prev[d] : -1;
            add_row(x, p0, p1, true);
        }
        S.rows = (int64_t) nxt.size();
        P.max_rows = std::max(P.max_rows, S.rows);
        cur.swap(nxt);
        row_prev.swap(row_cur);
    }
    // read-then-reset of lib/compute.cpp:854-859: a repeated proband id reads 0 the second time
    P.row_of.assign(m, -1);
    std::vector<uint8_t> seen(n, 0);
    for (int32_t i = 0; i < m; i++) {
        if (!seen[pro[i]]) { seen[pro[i]] = 1; P.row_of[i] = row_prev[pro[i]]; }
    }
    return true;
}

}  // namespace gk
