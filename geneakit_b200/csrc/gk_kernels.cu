// Hand-written sm_100a kernels of the dense-kinship path (no Triton, no library kernels).
//
// The step kernel replaces the reference's hot loop nest
//   compute_kinship_between_probands (lib/compute.cpp:571-588)  [n'(n'-1)/2 calls of]
//   compute_kinship                  (lib/compute.cpp:491-550)
//   compute_kinship_with_oneself     (lib/compute.cpp:553-568)
// for one pair of adjacent vertex cuts.  It is an HBM-bound FP64 gather-average: per output
// cell 4 gathered reads, 3 DADD + 1 DMUL, one 8-byte store.  Tensor cores are not used.
//
// PERSISTENT kernel, 2 CTAs per SM, each walking the unordered tile pairs (I >= J) of the new
// matrix with a grid stride, software-pipelined: while pair k is computed from stage buffer
// k&1, the parent rows of pair k+1 are landing in the other buffer and the descriptors of pair
// k+2 in a third descriptor slot.  Per pair:
//   1. stage E[rows(I)][cols(J)] of the PREVIOUS matrix.  rows(I)/cols(J) = the tile's
//      contiguous parent WINDOW + its LIST of scattered parent rows:
//        L[r][0..W)   = G[row r of I][window of J]      one 256-byte BULK async copy per row
//        R[j][0..W)   = G[list row j of J][window of I] (the window x list block through the
//                       symmetric entry)                one 256-byte BULK async copy per row
//        R[j][W + r]  = G[list row r of I][list col j of J]   8-byte cp.async gathers (L2 hits)
//      Bulk copies (cp.async.bulk -> UBLKCP, the TMA engine) complete on an mbarrier.
//   2. out[a][b] = 1/4 * ((p + q) + (r + s)) from shared memory (RANKED: grouping chosen per pair
//      by rank, reproducing the reference's rounding on pedigrees deeper than 26); unknown
//      parents index an all-zero row; tile pairs that share a parent individual apply
//      "same individual -> self kinship" (:522-528),
//   3. store tile (I,J) from registers and, transposed through shared memory, its mirror (J,I):
//      both coalesced.
// Pairs are enumerated I-major, so the CTAs resident at any time share rows(I): the scattered
// list rows are then served from the 126 MB L2 instead of HBM.
#include <cstdlib>
#include "gk_kernels.cuh"
#include "gk_plan.hpp"

namespace gk {

namespace {

constexpr int kThreads = 256;
constexpr int T = kTile, W = kWin, CAP = kListCap, DW = kDescWords;
constexpr int ZIDX = W + CAP;            // local index of the all-zero row / column (unknown parent)
constexpr int PL = W;                    // L pitch: (W + CAP + 1) rows x W doubles
constexpr int PR = W + CAP + 2;          // R pitch: (CAP + 1) rows x 74 doubles (16-byte aligned rows)
constexpr int OFFR = (W + CAP + 1) * PL; // R starts after L inside one stage buffer
constexpr int kStageDoubles = OFFR + (CAP + 1) * PR;
// NSTAGE = 2: double-buffered pipeline, 2 CTAs / SM.  NSTAGE = 1: one buffer (the transpose
// buffer aliases it), 4 CTAs / SM: the other resident CTAs hide the staging latency instead.
constexpr size_t step_smem(int nstage) {
    return (size_t) (nstage * kStageDoubles + (nstage > 1 ? T * (T + 1) : 0)) * sizeof(double) +
           (size_t) 3 * 2 * DW * sizeof(int32_t) + 2 * sizeof(unsigned long long);
}
static_assert(kStageDoubles % 2 == 0 && OFFR % 2 == 0 && PR % 2 == 0, "16-byte alignment of staged rows");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- mbarrier + bulk async copy (TMA engine, 1-D) ---------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "GK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra GK_DONE;\n"
        "bra GK_WAIT;\n"
        "GK_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_256(void *smem, const void *gmem, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(W * 8), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void pair_to_tiles(long long k, int &I, int &J) {
    int i = (int) ((sqrt(8.0 * (double) k + 1.0) - 1.0) * 0.5);
    while ((long long) i * (i + 1) / 2 > k) --i;
    while ((long long) (i + 1) * (i + 2) / 2 <= k) ++i;
    I = i;
    J = (int) (k - (long long) i * (i + 1) / 2);
}
__device__ __forceinline__ void advance_pair(int &I, int &J, int step) {   // pair index += step (row-major triangle)
    J += step;
    while (J > I) { J -= I + 1; ++I; }
}

// descriptors of tiles I and J -> slot (2*DW int32), 16-byte cp.async
__device__ __forceinline__ void load_desc(int32_t *slot, const int32_t *desc, int I, int J, int tid) {
    constexpr int CH = DW / 4;
    if (tid < 2 * CH) {
        const int t = tid < CH ? I : J, c = tid < CH ? tid : tid - CH;
        cp_async16(slot + (tid < CH ? 0 : DW) + 4 * c, desc + (size_t) t * DW + 4 * c);
    }
}

// Ask the L2 to stream in the scattered parent rows of a tile that the sweep reaches two waves
// from now: turns the first touch of those rows (otherwise random 32-byte sector misses of the
// list x list gathers) into sequential DRAM reads.
__device__ __forceinline__ void prefetch_list_rows(const int32_t *desc, int tile, const double *Gp, long long ldp, int lane) {
    const int32_t *d = desc + (size_t) tile * DW;
    const int nl = d[1] & 0xFFFF;
    if (lane < nl || lane + 32 < nl) {
        const unsigned row_bytes = (unsigned) (ldp * 8);
        for (int r = lane; r < nl; r += 32) {
            const char *p = reinterpret_cast<const char *>(Gp + (long long) d[kDescList + r] * ldp);
            for (unsigned off = 0; off < row_bytes; off += 65536u) {
                const unsigned n = row_bytes - off < 65536u ? row_bytes - off : 65536u;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p + off), "r"(n) : "memory");
            }
        }
    }
}

// stage E[rows(I)][cols(J)] of the previous matrix into one stage buffer (all asynchronous)
__device__ __forceinline__ void issue_stage(double *S, unsigned long long *bar, const int32_t *dI, const int32_t *dJ,
                                            const double *__restrict__ Gp, long long ldp, int lastrow, int tid) {
    const int c0I = dI[0], nlI = dI[1] & 0xFFFF, c0J = dJ[0], nlJ = dJ[1] & 0xFFFF;
    const int32_t *listI = dI + kDescList, *listJ = dJ + kDescList;
    const int nrI = W + nlI;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_expect_tx(bar, (unsigned) ((nrI + nlJ) * W * 8));
    // bulk copies, interleaved over the 8 warps (the issue loop is serial inside a warp):
    // copy c < nrI        : L[c][:]        = G[row c of I][window of J]
    // copy nrI <= c       : R[c - nrI][:W] = G[list row of J][window of I]
    {
        const int c = lane * 8 + warp;
        if (c < nrI) {
            int row = c < W ? c0I + c : listI[c - W];
            row = row > lastrow ? lastrow : row;
            bulk_copy_256(S + c * PL, Gp + (long long) row * ldp + c0J, bar);
        } else if (c < nrI + nlJ) {
            const int j = c - nrI;
            bulk_copy_256(S + OFFR + j * PR, Gp + (long long) listJ[j] * ldp + c0I, bar);
        }
    }
    {   // R[j][W + r] = G[list row r of I][list col j of J]: 8-byte gathers (L2-resident across the I sweep)
        const char *rowp[CAP / 8];
#pragma unroll
        for (int u = 0; u < CAP / 8; u++) {
            const int r = warp + 8 * u;
            rowp[u] = r < nlI ? reinterpret_cast<const char *>(Gp + (long long) listI[r] * ldp) : nullptr;
        }
        for (int j = lane; j < nlJ; j += 32) {
            const unsigned col8 = (unsigned) listJ[j] * 8u;
            double *dst = S + OFFR + j * PR + W + warp;
#pragma unroll
            for (int u = 0; u < CAP / 8; u++)
                if (rowp[u]) cp_async8(dst + 8 * u, rowp[u] + col8);
        }
    }
}

template <bool RANKED, int NSTAGE>
__global__ void __launch_bounds__(kThreads, NSTAGE == 1 ? 4 : 2) phi_step_kernel(const StepDev s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Sbuf = reinterpret_cast<double *>(smem_raw);                  // NSTAGE stage buffers
    double *O = NSTAGE > 1 ? Sbuf + NSTAGE * kStageDoubles : Sbuf;         // T x (T+1) transpose buffer (aliased when NSTAGE == 1)
    int32_t *Dbuf = reinterpret_cast<int32_t *>(Sbuf + NSTAGE * kStageDoubles + (NSTAGE > 1 ? T * (T + 1) : 0));   // 3 slots x (desc I, desc J)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(Dbuf + 3 * 2 * DW);   // one mbarrier per stage

    const int tid = threadIdx.x;
    const long long np = s.npairs;
    const int stride = (int) gridDim.x;
    const double *__restrict__ Gp = s.Gprev;
    const long long ldp = s.ld_prev, ldn = s.ld_new;
    const int lastrow = (int) (s.K_prev - 1);
    long long k = blockIdx.x;
    if (k >= np) return;

    // ---- one-time setup: zero rows / columns for unknown parents, mbarriers ---------------------
    for (int b = 0; b < NSTAGE; b++) {
        double *S = Sbuf + b * kStageDoubles;
        for (int i = tid; i < PL; i += kThreads) S[ZIDX * PL + i] = 0.0;            // L zero row
        for (int i = tid; i < PR; i += kThreads) S[OFFR + CAP * PR + i] = 0.0;       // R zero row
        for (int i = tid; i < CAP; i += kThreads) S[OFFR + i * PR + ZIDX] = 0.0;     // R zero column
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    int I, J, In, Jn, Inn, Jnn;              // tiles of pair k, k+stride, k+2*stride
    pair_to_tiles(k, I, J);
    In = I; Jn = J; advance_pair(In, Jn, stride);
    Inn = In; Jnn = Jn; advance_pair(Inn, Jnn, stride);

    // ---- prologue: descriptors of the first pair, then its rows + the second pair's descriptors
    load_desc(Dbuf, s.desc, I, J, tid);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    issue_stage(Sbuf, &bars[0], Dbuf, Dbuf + DW, Gp, ldp, lastrow, tid);
    if (k + stride < np) load_desc(Dbuf + 2 * DW, s.desc, In, Jn, tid);
    cp_async_commit();

    for (int it = 0; k < np; k += stride, ++it) {
        const int32_t *dI = Dbuf + (it % 3) * 2 * DW, *dJ = dI + DW;
        double *S = Sbuf + (NSTAGE > 1 ? (it & 1) : 0) * kStageDoubles;
        cp_async_wait_all();
        if (NSTAGE > 1) mbar_wait(&bars[it & 1], (unsigned) ((it >> 1) & 1));
        else mbar_wait(&bars[0], (unsigned) (it & 1));
        __syncthreads();       // pair k staged, next descriptors landed, everyone done with pair k-1
        if (NSTAGE > 1) {
            // ---- keep the memory system busy: rows of pair k+1, descriptors of pair k+2 -----------
            if (k + stride < np) {
                const int32_t *nI = Dbuf + ((it + 1) % 3) * 2 * DW;
                issue_stage(Sbuf + ((it + 1) & 1) * kStageDoubles, &bars[(it + 1) & 1], nI, nI + DW, Gp, ldp, lastrow, tid);
                if (k + 2 * (long long) stride < np) load_desc(Dbuf + ((it + 2) % 3) * 2 * DW, s.desc, Inn, Jnn, tid);
            }
            cp_async_commit();
        }

        // ---- pair k -----------------------------------------------------------------------------
        const int c0I = dI[0], cntI = dI[1] >> 16, cntJ = dJ[1] >> 16;
        const long long startI = dI[3], startJ = dJ[3];
        const int32_t *listI = dI + kDescList;
        bool slow = (I == J) || dI[2] < 0 || dJ[2] < 0;      // do the two tiles share a parent individual?
        if (!slow)
            for (int c = 0; c < dI[2]; c++) slow |= (dI[kDescConf + c] == J);

        constexpr int CPT = T * T / kThreads;                 // 4 cells per thread: rows a = tid/32 + 8c, column b = tid%32
        const int b = tid & 31, a_base = tid >> 5;
        const int lb = dJ[kDescLa + b];
        const int b0 = min(lb & 0xFFFF, ZIDX), b1 = min((lb >> 16) & 0xFFFF, ZIDX);
        // E(ra, cb) = cb < W ? L[ra][cb] : R[cb - W][ra]  ==  S[base(cb) + ra * stride(cb)]
        const int base0 = b0 < W ? b0 : OFFR + (b0 - W) * PR, str0 = b0 < W ? PL : 1;
        const int base1 = b1 < W ? b1 : OFFR + (b1 - W) * PR, str1 = b1 < W ? PL : 1;
        double v[CPT];
        if (!slow) {
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int a = a_base + 8 * c;
                const int la = dI[kDescLa + a];
                const int a0 = min(la & 0xFFFF, ZIDX), a1 = min((la >> 16) & 0xFFFF, ZIDX);
                const double p = S[base0 + a0 * str0], q = S[base0 + a1 * str0];
                const double r = S[base1 + a0 * str1], t = S[base1 + a1 * str1];
                if (RANKED) {
                    // lib/compute.cpp:529-548: the higher-ranked individual is expanded first, so the
                    // inner sums run over the LOWER-ranked one's parents.
                    v[c] = (dI[kDescRank + a] < dJ[kDescRank + b]) ? 0.25 * (__dadd_rn(p, q) + __dadd_rn(r, t))
                                                                  : 0.25 * (__dadd_rn(p, r) + __dadd_rn(q, t));
                } else {
                    v[c] = 0.25 * ((p + q) + (r + t));
                }
            }
        } else {
            const int ib0 = dJ[kDescPind + 2 * b], ib1 = dJ[kDescPind + 2 * b + 1];
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const int a = a_base + 8 * c;
                const int la = dI[kDescLa + a];
                const int a0 = min(la & 0xFFFF, ZIDX), a1 = min((la >> 16) & 0xFFFF, ZIDX);
                const int ia0 = dI[kDescPind + 2 * a], ia1 = dI[kDescPind + 2 * a + 1];
                double p = S[base0 + a0 * str0], q = S[base0 + a1 * str0];
                double r = S[base1 + a0 * str1], t = S[base1 + a1 * str1];
                // same individual on both sides -> its self kinship (lib/compute.cpp:522-528)
                if (ia0 >= 0 && (ia0 == ib0 || ia0 == ib1)) {
                    const double d = s.dprev[a0 < W ? c0I + a0 : listI[a0 - W]];
                    if (ia0 == ib0) p = d;
                    if (ia0 == ib1) r = d;
                }
                if (ia1 >= 0 && (ia1 == ib0 || ia1 == ib1)) {
                    const double d = s.dprev[a1 < W ? c0I + a1 : listI[a1 - W]];
                    if (ia1 == ib0) q = d;
                    if (ia1 == ib1) t = d;
                }
                if (RANKED) {
                    v[c] = (dI[kDescRank + a] < dJ[kDescRank + b]) ? 0.25 * (__dadd_rn(p, q) + __dadd_rn(r, t))
                                                                  : 0.25 * (__dadd_rn(p, r) + __dadd_rn(q, t));
                } else {
                    v[c] = 0.25 * ((p + q) + (r + t));
                }
            }
        }
        {
            double *g = s.Gnew + (startI + a_base) * ldn + startJ + b;
            const bool colok = b < cntJ;
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                if (colok && a_base + 8 * c < cntI) *g = v[c];
                g += 8 * ldn;
                if (NSTAGE > 1) O[(a_base + 8 * c) * (T + 1) + b] = v[c];
            }
        }
        if (I == J) {
            // self kinship of the new classes (lib/compute.cpp:553-568): a kept individual copies its
            // value (bitwise what the reference recomputes); a new one gets 1/2 + 1/2 phi(father, mother)
            if (tid < cntI) {
                const int la = dI[kDescLa + tid];
                const int a0 = la & 0xFFFF, a1 = (la >> 16) & 0xFFFF;
                const int i0 = dI[kDescPind + 2 * tid], i1 = dI[kDescPind + 2 * tid + 1];
                double d = 0.5;
                if (a0 != 0xFFFF && i0 == i1) d = s.dprev[a0 < W ? c0I + a0 : listI[a0 - W]];
                else if (a0 != 0xFFFF && a1 != 0xFFFF) d = 0.5 + 0.5 * (a1 < W ? S[a0 * PL + a1] : S[OFFR + (a1 - W) * PR + a0]);
                s.dnew[startI + tid] = d;
            }
            if (NSTAGE == 1) __syncthreads();      // all reads of the single stage buffer are done
        } else {
            __syncthreads();
            if (NSTAGE == 1) {                     // the transpose buffer aliases the (now dead) stage buffer
#pragma unroll
                for (int c = 0; c < CPT; c++) O[(a_base + 8 * c) * (T + 1) + b] = v[c];
                __syncthreads();
            }
            // mirror tile (J,I): thread (row bb = a_base + 8c of tile J, column aa = b of tile I); lanes along aa -> coalesced
            double *g = s.Gnew + (startJ + a_base) * ldn + startI + b;
            const bool colok = b < cntI;
            const double *o = O + b * (T + 1) + a_base;
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                if (colok && a_base + 8 * c < cntJ) *g = o[8 * c];
                g += 8 * ldn;
            }
            if (NSTAGE == 1) __syncthreads();      // before the next stage overwrites the aliased buffer
        }
        if (NSTAGE == 1) {
            if (k + stride < np) {
                const int32_t *nI = Dbuf + ((it + 1) % 3) * 2 * DW;
                // (the aliased transpose buffer covers L rows 0..32 only: the zero rows / columns survive)
                issue_stage(S, &bars[0], nI, nI + DW, Gp, ldp, lastrow, tid);
                if (k + 2 * (long long) stride < np) load_desc(Dbuf + ((it + 2) % 3) * 2 * DW, s.desc, Inn, Jnn, tid);
            }
            cp_async_commit();
        }
        // L2 prefetch duty: the CTA that will open tile row Inn (two strides ahead) streams its list rows in
        if (s.prefetch && Jnn == 0 && (tid >> 5) == 7 && k + 2 * (long long) stride < np) prefetch_list_rows(s.desc, Inn, Gp, ldp, tid & 31);
        I = In; J = Jn; In = Inn; Jn = Jnn;
        advance_pair(Inn, Jnn, stride);
    }
}

// Founders' matrix (lib/compute.cpp:655-661): classes of cut 0 are unrelated; self kinship 1/2.
__global__ void init_kernel(double *G0, double *d0, long long K0, long long ld0) {
    const long long total = K0 * ld0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total;
         i += (long long) gridDim.x * blockDim.x)
        G0[i] = 0.0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < K0;
         i += (long long) gridDim.x * blockDim.x)
        d0[i] = 0.5;
}

// Final expansion to the caller's proband order (duplicates, ancestors-as-probands) fused with
// the gen.phiMean reduction (geneakit/compute.py:124-132).  Probands of one class (sibship) have
// identical rows, so one CTA gathers a 1024-column slice of the class row ONCE (8-byte gathers,
// the row stays in L2) and then streams it, coalesced, to every local result row of that class,
// patching the "same individual" entries with the self kinship.  Each CTA also emits one
// (sum, diagonal sum) partial for the mean.
__global__ void __launch_bounds__(256) expand_kernel(const ExpandDev e) {
    const int a = blockIdx.x / e.nchunks, chunk = blockIdx.x - a * e.nchunks;
    const int c = e.cact[a];
    const long long j0 = (long long) chunk * kExpandCols + threadIdx.x;
    double v[4];
    int uj[4];
    const double *row = e.G + (long long) c * e.ld;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long j = j0 + 256 * k;
        const bool ok = j < e.m;
        uj[k] = ok ? e.ind[j] : -2;
        v[k] = ok ? row[e.cls[j]] : 0.0;
    }
    const double dc = e.d[c];
    double sum = 0.0, dsum = 0.0;
    for (int q = e.coff[a]; q < e.coff[a + 1]; q++) {
        const long long i = e.cmem[q];
        const int ui = e.ind[i];
        double *orow = e.out + (i - e.row0) * e.m;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long j = j0 + 256 * k;
            if (j < e.m) {
                const double val = (uj[k] == ui) ? dc : v[k];
                orow[j] = val;
                sum += val;
                if (i == j) dsum += val;
            }
        }
    }
    // warp shuffle -> shared -> one partial per CTA (fixed order: deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    }
    __shared__ double ws[8], wd[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = sum; wd[threadIdx.x >> 5] = dsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < 8; w++) { x += ws[w]; y += wd[w]; }
        e.partials[2 * (long long) blockIdx.x] = x;
        e.partials[2 * (long long) blockIdx.x + 1] = y;
    }
}

// All probands are top founders (single cut): 0.5*I, even for duplicated probands
// (lib/compute.cpp:655-661).  partials[0..1] = (sum, diagonal sum) of the local rows.
__global__ void eye_kernel(double *out, long long m, long long row0, long long row1, double *partials) {
    const long long total = (row1 - row0) * m;
    for (long long x = blockIdx.x * (long long) blockDim.x + threadIdx.x; x < total; x += (long long) gridDim.x * blockDim.x) {
        const long long i = row0 + x / m, j = x % m;
        out[x] = (i == j) ? 0.5 : 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { partials[0] = 0.5 * (double) (row1 - row0); partials[1] = partials[0]; }
}

__global__ void __launch_bounds__(256) mean_stage1_kernel(const double *partials, long long nblocks, double *scratch) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nblocks; i += 256LL * gridDim.x) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { scratch[2 * blockIdx.x] = sa[0]; scratch[2 * blockIdx.x + 1] = sb[0]; }
}

__global__ void __launch_bounds__(256) mean_finish_kernel(const double *partials, int nblocks, double *sums) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sums[0] = sa[0]; sums[1] = sb[0]; }
}

// gen.gc propagation (linear form of lib/compute.cpp:817-862): one CTA per
// member row, 16-byte vectorised, coalesced along the ancestor columns.
__global__ void __launch_bounds__(256) gc_step_kernel(const GcStepDev g) {
    const long long row = blockIdx.x;
    const int p0 = g.p0[row], p1 = g.p1[row];
    double2 *dst = reinterpret_cast<double2 *>(g.Vnew + row * g.ld);
    const long long n2 = g.ld / 2;
    if (p1 == -2) {          // kept: copy
        const double2 *a = reinterpret_cast<const double2 *>(g.Vprev + (long long) p0 * g.ld);
        for (long long c = threadIdx.x; c < n2; c += 256) dst[c] = a[c];
    } else {
        const double2 *a = p0 >= 0 ? reinterpret_cast<const double2 *>(g.Vprev + (long long) p0 * g.ld) : nullptr;
        const double2 *b = p1 >= 0 ? reinterpret_cast<const double2 *>(g.Vprev + (long long) p1 * g.ld) : nullptr;
        for (long long c = threadIdx.x; c < n2; c += 256) {
            double2 v = make_double2(0.0, 0.0);
            if (a) { const double2 x = a[c]; v.x += 0.5 * x.x; v.y += 0.5 * x.y; }
            if (b) { const double2 y = b[c]; v.x += 0.5 * y.x; v.y += 0.5 * y.y; }
            dst[c] = v;
        }
    }
    __syncthreads();
    for (int k = g.seed_off[row] + threadIdx.x; k < g.seed_off[row + 1]; k += 256)
        g.Vnew[row * g.ld + g.seed_col[k]] = 1.0;
}

__global__ void __launch_bounds__(256) gc_gather_kernel(const double *V, long long ld, const int32_t *row_of,
                                                        double *out, long long m, long long a) {
    const long long i = blockIdx.x;
    const int r = row_of[i];
    for (long long c = threadIdx.x; c < a; c += 256) out[i * a + c] = r >= 0 ? V[(long long) r * ld + c] : 0.0;
}

}  // namespace

static int g_stages = 2, g_prefetch = 1;
int step_prefetch_default() { return g_prefetch; }
size_t step_smem_bytes() { return step_smem(g_stages); }

static int g_step_grid[2][2] = { { 0, 0 }, { 0, 0 } };    // [ranked][nstage - 1]

template <bool RANKED, int NSTAGE>
static cudaError_t configure_one(int sms) {
    cudaError_t e;
    const size_t smem = step_smem(NSTAGE);
    if ((e = cudaFuncSetAttribute(phi_step_kernel<RANKED, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem))) return e;
    int occ = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phi_step_kernel<RANKED, NSTAGE>, kThreads, smem))) return e;
    g_step_grid[RANKED ? 1 : 0][NSTAGE - 1] = sms * (occ > 0 ? occ : 1);     // persistent: one wave, every CTA resident
    return cudaSuccess;
}

cudaError_t configure_kernels() {
    cudaError_t e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev))) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev))) return e;
    if ((e = configure_one<false, 1>(sms))) return e;
    if ((e = configure_one<false, 2>(sms))) return e;
    if ((e = configure_one<true, 1>(sms))) return e;
    if ((e = configure_one<true, 2>(sms))) return e;
    const char *env = getenv("GK_STAGES");
    if (env && (env[0] == '1' || env[0] == '2')) g_stages = env[0] - '0';
    env = getenv("GK_PREFETCH");
    if (env && (env[0] == '0' || env[0] == '1')) g_prefetch = env[0] - '0';
    return cudaSuccess;
}

cudaError_t launch_step(const StepDev &s, bool ranked, cudaStream_t st) {
    if (s.nt <= 0) return cudaSuccess;
    long long grid = g_step_grid[ranked ? 1 : 0][g_stages - 1];
    if (grid <= 0) grid = 296;
    if (grid > s.npairs) grid = s.npairs;
    const size_t smem = step_smem(g_stages);
    if (g_stages == 1) {
        if (ranked) phi_step_kernel<true, 1><<<(unsigned) grid, kThreads, smem, st>>>(s);
        else phi_step_kernel<false, 1><<<(unsigned) grid, kThreads, smem, st>>>(s);
    } else {
        if (ranked) phi_step_kernel<true, 2><<<(unsigned) grid, kThreads, smem, st>>>(s);
        else phi_step_kernel<false, 2><<<(unsigned) grid, kThreads, smem, st>>>(s);
    }
    return cudaGetLastError();
}

cudaError_t launch_init(double *G0, double *d0, int64_t K0, int64_t ld0, cudaStream_t st) {
    long long total = (long long) K0 * ld0;
    int blocks = (int) ((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    init_kernel<<<blocks, 256, 0, st>>>(G0, d0, K0, ld0);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandDev &e, cudaStream_t st) {
    if (e.nact <= 0) return cudaSuccess;
    expand_kernel<<<(unsigned) ((long long) e.nact * e.nchunks), 256, 0, st>>>(e);
    return cudaGetLastError();
}

cudaError_t launch_eye(double *out, int64_t m, int64_t row0, int64_t row1, double *partials, cudaStream_t st) {
    eye_kernel<<<148 * 4, 256, 0, st>>>(out, m, row0, row1, partials);
    return cudaGetLastError();
}

cudaError_t launch_mean_finish(const double *partials, int64_t nblocks, double *scratch, double *sums, cudaStream_t st) {
    mean_stage1_kernel<<<256, 256, 0, st>>>(partials, nblocks, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    mean_finish_kernel<<<1, 256, 0, st>>>(scratch, 256, sums);
    return cudaGetLastError();
}

cudaError_t launch_gc_step(const GcStepDev &g, cudaStream_t st) {
    if (g.rows <= 0) return cudaSuccess;
    gc_step_kernel<<<(unsigned) g.rows, 256, 0, st>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_gc_gather(const double *V, int64_t ld, const int32_t *row_of, double *out, int64_t m,
                             int64_t a, cudaStream_t st) {
    if (m <= 0 || a <= 0) return cudaSuccess;
    gc_gather_kernel<<<(unsigned) m, 256, 0, st>>>(V, ld, row_of, out, m, a);
    return cudaGetLastError();
}

}  // namespace gk
