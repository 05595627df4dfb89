/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Not product code: nothing under
 * geneakit_b200/ may include, link, dlopen or call this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Plain-C restatement of the reference's dense kinship path and of gen.gc, in RANK space
 * (individual k = position k of Pedigree::ids; sire[k]/dam[k] = parent rank or -1).
 * Citations are relative to /root/reference/geneakit/.  Parity is PINNED: tests/test_oracle.py
 * checks this file against the reference's golden vectors (test/test_geneaJi.py:26-41,
 * test/test_genea140.py:29-30,52-54, test/test_custom.py:11-15) and against outputs of the
 * unmodified reference compiled here (oracle/_ref/libgkref.so, fixtures in tests/golden/).
 *
 * The restatement is deliberately literal (same recursion, same branch order, same
 * `kinship += 0.5 * x` accumulation) so that FP64 rounding on deep pedigrees (> 26
 * generations) is reproduced, not just the exact dyadic case.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------------
 * Loader order: lib/create.cpp:121-155 populate_pedigree_dfs + :158-199 convert_pedigree.
 * Input in FILE order: fidx[i]/midx[i] = file position of the parent or -1 (id 0 = unknown).
 * Output order[r] = file position of the individual that receives rank r
 * (father subtree, then mother subtree, then self; roots tried in file order).
 * Iterative to survive deep pedigrees.
 * ------------------------------------------------------------------------------------ */
int gko_dfs_order(int n, const int *fidx, const int *midx, int *order) {
    unsigned char *state = (unsigned char *) calloc((size_t) n, 1); /* 0 new,1 f done,2 m done,3 emitted */
    int *stack = (int *) malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
    if (!state || !stack) { free(state); free(stack); return -1; }
    int rank = 0;
    for (int root = 0; root < n; root++) {
        if (state[root] == 3) continue;
        int sp = 0;
        stack[sp++] = root;
        while (sp > 0) {
            int x = stack[sp - 1];
            if (state[x] == 3) { sp--; continue; }
            if (state[x] == 0) {
                state[x] = 1;
                if (fidx[x] >= 0 && state[fidx[x]] != 3) {
                    if (state[fidx[x]] != 0) { free(state); free(stack); return -2; } /* cycle */
                    stack[sp++] = fidx[x];
                    continue;
                }
            }
            if (state[x] == 1) {
                state[x] = 2;
                if (midx[x] >= 0 && state[midx[x]] != 3) {
                    if (state[midx[x]] != 0) { free(state); free(stack); return -2; }
                    stack[sp++] = midx[x];
                    continue;
                }
            }
            state[x] = 3;
            order[rank++] = x;
            sp--;
        }
    }
    free(state);
    free(stack);
    return rank == n ? 0 : -3;
}

/* lib/identify.cpp:55-65 get_proband_ids: childless individuals.  Writes RANKS (caller maps to
 * ids and sorts by id, as the reference does).  Returns the count. */
int gko_childless(int n, const int *sire, const int *dam, int *out) {
    unsigned char *has = (unsigned char *) calloc((size_t) n + 1, 1);
    for (int k = 0; k < n; k++) {
        if (sire[k] >= 0) has[sire[k]] = 1;
        if (dam[k] >= 0) has[dam[k]] = 1;
    }
    int c = 0;
    for (int k = 0; k < n; k++) if (!has[k]) out[c++] = k;
    free(has);
    return c;
}

/* lib/identify.cpp:28-38 get_founder_ids: both parents unknown. */
int gko_parentless(int n, const int *sire, const int *dam, int *out) {
    int c = 0;
    for (int k = 0; k < n; k++) if (sire[k] < 0 && dam[k] < 0) out[c++] = k;
    return c;
}

/* ---------------------------------------------------------------------------------------
 * Vertex cuts: lib/compute.cpp:119-135 cut_vertices and helpers :28-114.
 *   generations[k]  (:43-52)  = individuals reachable from a proband by SOME path of length k
 *   bottom_up[i]    (:55-73)  = union of generations[0 .. G-1-i]   (after the reverse at :71)
 *   top_down[i]     (:76-94)  = union of generations[G-1-i .. G-1] (generations reversed at :78)
 *   cut[i]          (:97-114) = bottom_up[i] intersect top_down[i]
 *   cut[G-1] := the caller's proband vector, order and duplicates kept (:132)
 * With first[x]/last[x] = smallest/largest level holding x, membership of cut i is
 * first[x] <= G-1-i <= last[x].  Members are emitted in rank order (the reference's order is
 * pointer order, which no result depends on).
 * Returns G (number of cuts) or <0.  cut_off has G+1 entries; members are ranks.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    int G;
    int64_t *off;   /* G+1 */
    int *mem;       /* off[G] */
} gko_cuts_t;

static void cuts_free(gko_cuts_t *c) { free(c->off); free(c->mem); c->off = NULL; c->mem = NULL; }

static int cuts_build(int n, const int *sire, const int *dam, int m, const int *pro, gko_cuts_t *out) {
    int *first = (int *) malloc(sizeof(int) * (size_t) n);
    int *last = (int *) malloc(sizeof(int) * (size_t) n);
    int *cur = (int *) malloc(sizeof(int) * (size_t) n);
    int *nxt = (int *) malloc(sizeof(int) * (size_t) n);
    int *stamp = (int *) malloc(sizeof(int) * (size_t) n);
    if (!first || !last || !cur || !nxt || !stamp) return -1;
    for (int k = 0; k < n; k++) { first[k] = -1; last[k] = -1; stamp[k] = -1; }
    int ncur = 0;
    for (int i = 0; i < m; i++) {
        int p = pro[i];
        if (p < 0 || p >= n) { free(first); free(last); free(cur); free(nxt); free(stamp); return -2; }
        if (stamp[p] != 0) { stamp[p] = 0; cur[ncur++] = p; }
    }
    int G = 0;
    while (ncur > 0) {               /* get_generations :43-52 */
        for (int i = 0; i < ncur; i++) {
            int x = cur[i];
            if (first[x] < 0) first[x] = G;
            last[x] = G;
        }
        int nn = 0;
        for (int i = 0; i < ncur; i++) {  /* get_previous_generation :28-40 */
            int x = cur[i];
            int f = sire[x], d = dam[x];
            if (f >= 0 && stamp[f] != G + 1) { stamp[f] = G + 1; nxt[nn++] = f; }
            if (d >= 0 && stamp[d] != G + 1) { stamp[d] = G + 1; nxt[nn++] = d; }
        }
        int *t = cur; cur = nxt; nxt = t;
        ncur = nn;
        G++;
    }
    out->G = G;
    out->off = (int64_t *) calloc((size_t) G + 1, sizeof(int64_t));
    /* count */
    for (int i = 0; i + 1 < G; i++) {
        int L = G - 1 - i;
        int64_t c = 0;
        for (int x = 0; x < n; x++) if (first[x] >= 0 && first[x] <= L && L <= last[x]) c++;
        out->off[i + 1] = c;
    }
    if (G > 0) out->off[G] = m;
    for (int i = 0; i < G; i++) out->off[i + 1] += out->off[i];
    out->mem = (int *) malloc(sizeof(int) * (size_t) (out->off[G] > 0 ? out->off[G] : 1));
    for (int i = 0; i + 1 < G; i++) {
        int L = G - 1 - i;
        int64_t w = out->off[i];
        for (int x = 0; x < n; x++) if (first[x] >= 0 && first[x] <= L && L <= last[x]) out->mem[w++] = x;
    }
    if (G > 0) memcpy(out->mem + out->off[G - 1], pro, sizeof(int) * (size_t) m);
    free(first); free(last); free(cur); free(nxt); free(stamp);
    return G;
}

int gko_cut_sizes(int n, const int *sire, const int *dam, int m, const int *pro, int *sizes, int cap) {
    gko_cuts_t c;
    int G = cuts_build(n, sire, dam, m, pro, &c);
    if (G < 0) return G;
    for (int i = 0; i < G && i < cap; i++) sizes[i] = (int) (c.off[i + 1] - c.off[i]);
    cuts_free(&c);
    return G;
}

/* lib/compute.cpp:591-635 get_required_memory_for_kinships */
double gko_required_gb(int n, const int *sire, const int *dam, int m, const int *pro) {
    gko_cuts_t c;
    int G = cuts_build(n, sire, dam, m, pro, &c);
    if (G < 2) { if (G >= 0) cuts_free(&c); return 0.0; }
    int64_t best = -1; int64_t s1 = 0, s2 = 0;
    for (int i = 0; i + 1 < G; i++) {
        int64_t a = c.off[i + 1] - c.off[i], b = c.off[i + 2] - c.off[i + 1];
        if (a + b > best) { best = a + b; s1 = a; s2 = b; }   /* first maximum wins, :622-628 */
    }
    cuts_free(&c);
    return (1.0 * s1 * s1 + 1.0 * s2 * s2) * sizeof(double) / 1e9;
}

/* ---------------------------------------------------------------------------------------
 * Pair rule: lib/compute.cpp:491-550 compute_kinship, branch for branch.
 * prev[x] = slot of x in the previous cut's matrix A (n_prev x n_prev) or -1.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    const int *sire, *dam, *prev;
    const double *A;
    int64_t na;
} pair_ctx;

static double pair_kinship(const pair_ctx *c, int i1, int i2) {
    double kinship = 0.0;
    const int p1 = c->prev[i1], p2 = c->prev[i2];
    if (p1 != -1 && p2 != -1) {                      /* :498-500 */
        kinship = c->A[(int64_t) p1 * c->na + p2];
    } else if (p1 != -1) {                           /* :501-511 */
        if (c->sire[i2] >= 0) kinship += 0.5 * pair_kinship(c, i1, c->sire[i2]);
        if (c->dam[i2] >= 0) kinship += 0.5 * pair_kinship(c, i1, c->dam[i2]);
    } else if (p2 != -1) {                           /* :512-521 */
        if (c->sire[i1] >= 0) kinship += 0.5 * pair_kinship(c, c->sire[i1], i2);
        if (c->dam[i1] >= 0) kinship += 0.5 * pair_kinship(c, c->dam[i1], i2);
    } else if (i1 == i2) {                           /* :522-528 same individual (rank equal) */
        kinship = 0.5;
        if (c->sire[i1] >= 0 && c->dam[i2] >= 0)
            kinship += 0.5 * pair_kinship(c, c->sire[i1], c->dam[i2]);
    } else if (i1 < i2) {                            /* :529-538 expand the higher rank */
        if (c->sire[i2] >= 0) kinship += 0.5 * pair_kinship(c, i1, c->sire[i2]);
        if (c->dam[i2] >= 0) kinship += 0.5 * pair_kinship(c, i1, c->dam[i2]);
    } else {                                         /* :539-548 */
        if (c->sire[i1] >= 0) kinship += 0.5 * pair_kinship(c, c->sire[i1], i2);
        if (c->dam[i1] >= 0) kinship += 0.5 * pair_kinship(c, c->dam[i1], i2);
    }
    return kinship;
}

/* lib/compute.cpp:638-700 compute_kinships.  out = m*m doubles, row-major, caller order. */
int gko_phi(int n, const int *sire, const int *dam, int m, const int *pro, double *out) {
    gko_cuts_t c;
    int G = cuts_build(n, sire, dam, m, pro, &c);
    if (G <= 0) return G < 0 ? G : -4;
    int64_t n0 = c.off[1] - c.off[0];
    double *A = (double *) calloc((size_t) (n0 * n0 > 0 ? n0 * n0 : 1), sizeof(double));   /* :655-661 */
    if (!A) { cuts_free(&c); return -1; }
    for (int64_t i = 0; i < n0; i++) A[i * n0 + i] = 0.5;
    int *prev = (int *) malloc(sizeof(int) * (size_t) n);
    for (int k = 0; k < n; k++) prev[k] = -1;          /* data.hpp:49-56 default -1 */
    int64_t na = n0;
    for (int g = 0; g + 1 < G; g++) {                  /* :673-698 */
        const int *cut = c.mem + c.off[g];
        for (int64_t i = 0; i < na; i++) prev[cut[i]] = (int) i;   /* :678-681 (stale slots stay, as upstream) */
        const int *nx = c.mem + c.off[g + 1];
        int64_t nb = c.off[g + 2] - c.off[g + 1];
        double *B = (double *) calloc((size_t) (nb * nb > 0 ? nb * nb : 1), sizeof(double));
        if (!B) { free(A); free(prev); cuts_free(&c); return -1; }
        pair_ctx ctx = { sire, dam, prev, A, na };
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nb; i++) {             /* :553-568 */
            int x = nx[i];
            double k = 0.5;
            if (sire[x] >= 0 && dam[x] >= 0) k += 0.5 * pair_kinship(&ctx, sire[x], dam[x]);
            B[i * nb + i] = k;
        }
        #pragma omp parallel for schedule(dynamic, 16)
        for (int64_t i = 0; i < nb; i++) {             /* :571-588 */
            for (int64_t j = 0; j < i; j++) {
                double k = pair_kinship(&ctx, nx[i], nx[j]);
                B[i * nb + j] = k;
                B[j * nb + i] = k;
            }
        }
        free(A);
        A = B;
        na = nb;
    }
    memcpy(out, A, sizeof(double) * (size_t) (na * na));
    free(A); free(prev); cuts_free(&c);
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * gen.gc, literal: lib/compute.cpp:817-827 add_contribution (DFS over children, stop at the
 * first proband on each path) and :832-862 compute_genetic_contributions, including the
 * "read then reset" of :854-859 (a duplicated proband id reads 0 the second time).
 * Exponential under inbreeding, like upstream -- small cases only.
 * ------------------------------------------------------------------------------------ */
typedef struct { const int64_t *coff; const int *child; const unsigned char *is_pro; double *acc; } gc_ctx;

static void add_contribution(const gc_ctx *c, int x, int depth) {
    if (c->is_pro[x]) {
        c->acc[x] += pow(0.5, depth);
    } else {
        for (int64_t e = c->coff[x]; e < c->coff[x + 1]; e++) add_contribution(c, c->child[e], depth + 1);
    }
}

int gko_gc_dfs(int n, const int *sire, const int *dam, int m, const int *pro,
               int a, const int *anc, double *out) {
    /* children lists in the order Individual's constructor appends them (individual.hpp:50-62):
     * individuals are created in rank order; father's list gets the child, then mother's. */
    int64_t *coff = (int64_t *) calloc((size_t) n + 2, sizeof(int64_t));
    for (int k = 0; k < n; k++) {
        if (sire[k] >= 0) coff[sire[k] + 1]++;
        if (dam[k] >= 0) coff[dam[k] + 1]++;
    }
    for (int k = 0; k < n; k++) coff[k + 1] += coff[k];
    int *child = (int *) malloc(sizeof(int) * (size_t) (coff[n] > 0 ? coff[n] : 1));
    int64_t *w = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 1));
    memcpy(w, coff, sizeof(int64_t) * (size_t) (n + 1));
    for (int k = 0; k < n; k++) {
        if (sire[k] >= 0) child[w[sire[k]]++] = k;
        if (dam[k] >= 0) child[w[dam[k]]++] = k;
    }
    unsigned char *is_pro = (unsigned char *) calloc((size_t) n + 1, 1);
    double *acc = (double *) calloc((size_t) n + 1, sizeof(double));
    for (int i = 0; i < m; i++) {
        if (pro[i] < 0 || pro[i] >= n) { free(coff); free(child); free(w); free(is_pro); free(acc); return -2; }
        is_pro[pro[i]] = 1;
    }
    gc_ctx ctx = { coff, child, is_pro, acc };
    for (int j = 0; j < a; j++) {
        if (anc[j] < 0 || anc[j] >= n) { free(coff); free(child); free(w); free(is_pro); free(acc); return -2; }
        add_contribution(&ctx, anc[j], 0);
        for (int i = 0; i < m; i++) {
            out[(int64_t) i * a + j] = acc[pro[i]];
            acc[pro[i]] = 0.0;
        }
    }
    free(coff); free(child); free(w); free(is_pro); free(acc);
    return 0;
}

/* gen.gc, linear form: v(anc)=1; in rank order
 * v(i) = 0.5*pass(sire) + 0.5*pass(dam), pass(x) = 0 if x is a proband else v(x).
 * Bit-identical to the DFS while depth <= 52 (sums of distinct-path powers of two stay exact);
 * checked against gko_gc_dfs and the reference in tests/test_oracle.py.  Polynomial time. */
int gko_gc(int n, const int *sire, const int *dam, int m, const int *pro,
           int a, const int *anc, double *out) {
    unsigned char *is_pro = (unsigned char *) calloc((size_t) n + 1, 1);
    unsigned char *dup = (unsigned char *) calloc((size_t) m + 1, 1);
    for (int i = 0; i < m; i++) {
        if (pro[i] < 0 || pro[i] >= n) { free(is_pro); free(dup); return -2; }
        if (is_pro[pro[i]]) dup[i] = 1;      /* a later occurrence of the same proband */
        is_pro[pro[i]] = 1;
    }
    for (int j = 0; j < a; j++) if (anc[j] < 0 || anc[j] >= n) { free(is_pro); free(dup); return -2; }
    int err = 0;
    #pragma omp parallel
    {
        double *v = (double *) malloc(sizeof(double) * (size_t) (n + 1));
        if (!v) { err = -1; }
        else {
            #pragma omp for schedule(dynamic, 4)
            for (int j = 0; j < a; j++) {
                int src = anc[j];
                v[src] = 1.0;
                for (int k = src + 1; k < n; k++) {
                    double s = 0.0;
                    int f = sire[k], d = dam[k];
                    if (f >= src && !is_pro[f]) s += 0.5 * v[f];
                    if (d >= src && !is_pro[d]) s += 0.5 * v[d];
                    v[k] = s;
                }
                /* :854-859: value read once per proband id, later duplicates read the reset 0 */
                for (int i = 0; i < m; i++)
                    out[(int64_t) i * a + j] = (dup[i] || pro[i] < src) ? 0.0 : v[pro[i]];
            }
        }
        free(v);
    }
    free(is_pro); free(dup);
    return err;
}

int gko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void gko_set_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void) t;
#endif
}
