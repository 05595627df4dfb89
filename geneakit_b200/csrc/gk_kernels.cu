// Hand-written sm_100a kernels of the dense-kinship path (no Triton, no library kernels).
//
// The step kernel replaces the reference's hot loop nest
//   compute_kinship_between_probands (lib/compute.cpp:571-588)  [n'(n'-1)/2 calls of]
//   compute_kinship                  (lib/compute.cpp:491-550)
// for one pair of adjacent vertex cuts; the diagonal kernel replaces
//   compute_kinship_with_oneself     (lib/compute.cpp:553-568)
// and applies the "same individual" rule (:522-528).  The step is an HBM-bound FP64
// gather-average, B = P A P^T with two 1/2 entries per row of P.  Tensor cores are not used.
//
// FACTORISED, two phases per PANEL I of 32 new classes, inside ONE persistent launch:
//   phase 1 (I, cb):  N[i][c] = 1/2 (A[F_i][c] + A[M_i][c]),  i in I, c in a block of 64 columns.
//            Reads 64-column segments of whole rows of A (512 B, coalesced; rows may live on a peer
//            GPU: the row table holds CUDA-IPC mappings, the loads cross NVLink), transposes the
//            32 x 64 tile through shared memory and stores N^T[c][0..32) (16 KB contiguous) into
//            one of kRing panel buffers that stay resident in the 126 MB L2.
//   phase 2 (I, J):   B[j][i] = 1/2 (N^T[F_j][i] + N^T[M_j][i]),  j in tile J, i in I.
//            Reads two 256-byte rows of the L2-resident panel per class j, stores the 32 x 32 tile
//            (J,I) directly and its mirror (I,J) transposed through shared memory: both coalesced.
//            Only tiles J <= I are evaluated (full_rows = 0); a rank that owns a row block of a
//            sharded matrix evaluates every J and stores only its own rows (full_rows = 1).
// Work items are dealt round-robin to the resident CTAs in QUEUE ORDER
//   P1(0), P1(1), P2(0), P1(2), P2(1), ...
// and an item waits only on items that precede it in the queue (phase 2 of panel I on the ncb
// phase-1 items of I; phase 1 of panel I on phase 2 of panel I - kRing, whose ring slot it
// overwrites), so the wait can never deadlock: the lowest unfinished item never waits.
// No 8-byte gathers, no index arithmetic per element: every global access is a 256/512-byte
// segment.  In RANKED mode (pedigrees deeper than 26 generations) the planner orders the classes
// by descending rank, which makes "J <= I" the reference's rounding order (gk_plan.hpp).
#include <cstdlib>
#include <cstring>
#include "gk_kernels.cuh"
#include "gk_plan.hpp"

namespace gk {

namespace {

constexpr int kWarps = kStepWarps;        // 3 per SM sub-partition, one CTA per SM
constexpr int kThreads = kWarps * 32;
constexpr int T = kPanel;                 // 32
constexpr int PS = T + 2;                 // staging pitch (doubles): 16-byte aligned rows, conflict-free 128-bit transposed reads
constexpr int kStageDoubles = (2 * T + 1) * PS; // up to 64 distinct parent rows of a panel per warp + one all-zero row (slot 64: unknown parent)
constexpr size_t kStepSmem = (size_t) kWarps * kStageDoubles * sizeof(double);

// ---- second step kernel (multi-GPU): phase 1 through the TMA engine ---------------------------------
constexpr int NW2 = kStepP2Warps;         // phase-2 warps per CTA (LDGSTS, one private stage each)
constexpr int NC1 = kStepP1Consumers;     // phase-1 consumer warps per CTA
constexpr int NS1 = kStepP1Stages;        // phase-1 TMA stages per CTA
static_assert(NC1 % NS1 == 0 && (32 / (NC1 / NS1)) % 4 == 0, "consumer warps per stage must divide the columns");
constexpr int kThreadsTma = (NW2 + NC1 + 1) * 32;   // + the phase-1 producer warp
constexpr int S1 = kP1Classes;            // phase 1: classes per item (8)
constexpr int C1 = kP1Cols;               // phase 1: columns per item (256): 2 KB row segments, the size the TMA engine needs
constexpr int P1PITCH = C1 * 8 + 16;      // bytes between staged rows: 516 words = 4 mod 32 -> conflict-free (class, column) reads
constexpr int P1STAGE = (2 * S1 + 1) * P1PITCH; // up to 16 distinct parent rows of the 8 classes + one all-zero row (slot 16)
constexpr size_t kStepSmemTma = (size_t) NW2 * kStageDoubles * sizeof(double) + (size_t) NS1 * P1STAGE + 2 * NS1 * sizeof(unsigned long long);

__device__ __forceinline__ int ld_relaxed(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Poll a completion counter.  Relaxed loads (an acquire would invalidate the whole L1 on every
// iteration); everything read after the wait comes from L2 (bulk copies / stores to fresh lines).
__device__ __forceinline__ void wait_count(const int *p, int target) {
    unsigned ns = 64;
    while (ld_relaxed(p) < target) { __nanosleep(ns); if (ns < 1024) ns *= 2; }
}
__device__ __forceinline__ void signal_count(int *p) {
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(p) : "memory");
}

// L2 eviction priorities: the panel ring must stay resident (evict_last), the streams must not push it out
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_normal() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(double *p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;\n" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

// 16-byte asynchronous copy global -> shared through L2 only (LDGSTS); bytes = 0 zero-fills (unknown parent)
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, unsigned bytes, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(dst), "l"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- mbarrier + bulk asynchronous copy (the TMA engine, 1-D) -----------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "GK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra GK_DONE;\n"
        "bra GK_WAIT;\n"
        "GK_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy(unsigned dst, const void *src, unsigned bytes, unsigned bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// Row r of a row-sharded matrix: shard q = r / rows_per_shard (every rank owns the same number of
// panels), local row r - q * rows_per_shard.  One shard: no lookup at all.
struct RowLookup {
    double *const *bases;     // shared-memory copy of RowTable::base
    double *base0;
    int rps;                  // rows per shard
    int n;
    long long ld;
    __device__ __forceinline__ const double *operator()(int row) const {
        if (n == 1) return base0 + (long long) row * ld;
        const int q = row / rps;
        return bases[q] + (long long) (row - q * rps) * ld;
    }
};

// One CTA per SM, WARP-AUTONOMOUS: every warp walks the work queue on its own (items gw, gw + nwarps,
// ...) with a private 17 KB shared-memory stage.  An item's 64 parent rows (256 bytes each: 32
// columns of rows of the previous matrix for phase 1, rows of the L2-resident panel for phase 2)
// are staged with asynchronous 16-byte copies (LDGSTS: two rows per instruction, no registers
// held), then the warp produces the 32 x 32 tile(s) straight from shared memory.  No CTA-wide barrier.
// (Measured alternatives, all slower on B200: 1-D TMA bulk copies sustain only ~1 row per 70 cycles
// per SM, 4x too slow for 256-byte rows; dedicated producer warps feeding mbarrier stages serialise
// behind the dependency waits; issuing the next item's copies before publishing the current one
// delays the release fence by a full memory latency.)
struct StepItem {
    int info, local;
    int r0, r1;      // lane l: distinct parent rows l and 32 + l of the item's panel (-1 past the end)
    int n;           // how many distinct rows
    int slots;       // lane = class: slot of its father row | slot of its mother row << 8
};

__global__ void __launch_bounds__(kThreads, 1) phi_step_kernel(const StepDev s) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double *sprev[kMaxShards];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kMaxShards) sprev[threadIdx.x] = s.prev.base[threadIdx.x];
    double *const S = smem + (size_t) warp * kStageDoubles;
    for (int i = lane; i < PS; i += 32) S[kZeroSlot * PS + i] = 0.0;       // the zero row is never overwritten
    __syncthreads();
    const RowLookup prev{ sprev, s.prev.base[0], s.prev.n > 1 ? s.prev.row0[1] : 1, s.prev.n, s.ld_prev };
    const unsigned long long pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    const unsigned long long pol_out = s.out_policy == 0 ? pol_stream : (s.out_policy == 2 ? pol_keep : policy_evict_normal());
    const int half = lane >> 4, l16 = lane & 15;
    // copy r: distinct rows 2r (lanes 0-15) and 2r + 1 (lanes 16-31) of the panel -> slots 2r, 2r + 1
    const unsigned sdst = smem_u32(S) + (unsigned) ((half * PS + l16 * 2) * sizeof(double));
    double *const out = s.out_base;              // own rows of the new matrix: global row r at out + (r - out_row0) * ld_new
    const long long ldn = s.ld_new;
    const long long nw = (long long) gridDim.x * kWarps;
    int g = 0;
    int ready_panel = -1;        // phase 1 of this panel is known complete
    int freed_panel = -1;        // phase 2 of this panel is known complete (its ring slot is free)

    StepItem cur;
    for (long long q = (long long) blockIdx.x * kWarps + warp; q < s.total_items; q += nw) {
        // ---- decode: group -> (phase, panel), position in the group, the panel's distinct parent rows ----
        while (q >= s.grp_off[g + 1]) ++g;
        cur.info = s.grp_info[g];
        cur.local = (int) (q - s.grp_off[g]);
        const int I = cur.info & 0x3FFFFFFF, local = cur.local;
        const bool p1 = (cur.info >> 30) == 0;
        const int P = p1 ? I : local;                        // the panel / tile whose parent rows are staged
        cur.r0 = s.prow[P * 2 * T + lane]; cur.r1 = s.prow[P * 2 * T + T + lane];
        cur.n = s.pcnt[P];
        cur.slots = s.slots[P * T + lane];
        double *buf = s.ring + (long long) ((I - s.panel_begin) % s.ring_n) * s.ring_stride;
        if (!p1 && I != ready_panel) {          // phase 1 of panel I must be complete before its rows are read
            if (lane == 0 && !s.debug_nowait) wait_count(s.done1 + I, s.ncb);
            ready_panel = I;
            __syncwarp();
        }
        // ---- stage the distinct rows: 32 columns of rows of the previous matrix (phase 1) or whole rows
        //      of the L2-resident panel (phase 2), two rows per instruction --------------------------------
        {
            const unsigned long long pol = p1 ? pol_stream : pol_keep;
            const long long coff = p1 ? (long long) local * T + l16 * 2 : (long long) l16 * 2;
            const int nops = (cur.n + 1) >> 1;
#pragma unroll 4
            for (int r = 0; r < nops; r++) {
                const int row = __shfl_sync(0xffffffffu, r < 16 ? cur.r0 : cur.r1, (2 * r + half) & 31);
                const double *src = row < 0 ? s.zero_row : (p1 ? prev(row) : buf + (long long) row * T) + coff;
                cp_async16(sdst + (unsigned) (2 * r * PS * sizeof(double)), src, row >= 0 ? 16u : 0u, pol);
            }
        }
        // lane = class: its father / mother rows as 128-bit columns pairs
        const double2 *sf = reinterpret_cast<const double2 *>(S + (cur.slots & 0xFF) * PS);
        const double2 *sm = reinterpret_cast<const double2 *>(S + ((cur.slots >> 8) & 0xFF) * PS);
        int *done;
        if (p1) {
            // -------- phase 1: N^T[c][i], c in [c0, c0 + 32): lane = class i, two columns per 128-bit read
            const int old = I - s.ring_n;    // the ring slot must have been drained by phase 2 of its previous panel
            if (old >= s.panel_begin && old > freed_panel) {
                if (lane == 0 && !s.debug_nowait) wait_count(s.done2 + old, s.full_rows ? s.np : old + 1);
                freed_panel = old;
            }
            cp_async_wait_all();
            __syncwarp();
            double *dst = buf + (long long) local * T * T + lane;
            // rows of the panel that no tile J <= I will read are not written (half of them on average)
            const unsigned need = s.full_rows ? 0xffffffffu : __ballot_sync(0xffffffffu, s.first_use[local * T + lane] <= I);
#pragma unroll 8
            for (int c2 = 0; c2 < T / 2; c2++) {
                const double2 x = sf[c2], y = sm[c2];
                if (need >> (2 * c2) & 1u) st_hint(dst + (2 * c2) * T, 0.5 * (x.x + y.x), pol_keep);
                if (need >> (2 * c2 + 1) & 1u) st_hint(dst + (2 * c2 + 1) * T, 0.5 * (x.y + y.y), pol_keep);
            }
            done = s.done1 + I;
        } else {
            // -------- phase 2: tile (J, I) and its mirror ---------------------------------------
            const int J = local;
            double *drow = out + ((long long) J * T - s.out_row0) * ldn + I * T + lane;    // row j of tile J, columns of panel I
            double *mrow = out + ((long long) I * T - s.out_row0) * ldn + J * T + lane;    // row i of panel I, columns of tile J
            cp_async_wait_all();
            __syncwarp();
            if (J != I || s.full_rows) {
                if (!s.full_rows && J >= s.panel_begin) {     // (a tile J of another rank is pulled by its owner: pull_mirror_kernel)
                    // direct: lane = class i of the panel, row j: B[j][i] = 1/2 (N^T[F_j][i] + N^T[M_j][i])
#pragma unroll 8
                    for (int j = 0; j < T; j++) {
                        const int sl = __shfl_sync(0xffffffffu, cur.slots, j);
                        st_hint(drow + (long long) j * ldn, 0.5 * (S[(sl & 0xFF) * PS + lane] + S[((sl >> 8) & 0xFF) * PS + lane]), pol_out);
                    }
                }
                // mirror: lane = class j of the tile, rows i, i + 1: B[i][j] = the same value
#pragma unroll 8
                for (int i2 = 0; i2 < T / 2; i2++) {
                    const double2 x = sf[i2], y = sm[i2];
                    st_hint(mrow + (long long) (2 * i2) * ldn, 0.5 * (x.x + y.x), pol_out);
                    st_hint(mrow + (long long) (2 * i2 + 1) * ldn, 0.5 * (x.y + y.y), pol_out);
                }
            } else {
                // diagonal tile: V[j][i] is valid where j <= i (in RANKED mode that is the reference's
                // grouping); the other triangle is its mirror.  out[r][lane] = r <= lane ? V[r][lane] : V[lane][r]
                const double *sfl = reinterpret_cast<const double *>(sf), *sml = reinterpret_cast<const double *>(sm);
#pragma unroll 4
                for (int r = 0; r < T; r++) {
                    const int sl = __shfl_sync(0xffffffffu, cur.slots, r);
                    const double a = 0.5 * (S[(sl & 0xFF) * PS + lane] + S[((sl >> 8) & 0xFF) * PS + lane]);   // V[r][lane]
                    const double b = 0.5 * (sfl[r] + sml[r]);                                                    // V[lane][r]
                    st_hint(mrow + (long long) r * ldn, r <= lane ? a : b, pol_out);
                }
            }
            done = s.done2 + I;
        }
        __syncwarp();                                    // every lane is done reading the stage (and has issued its stores)
        if (lane == 0) signal_count(done);
    }
}

// SECOND STEP KERNEL, used by the ranks of a multi-GPU job (the parent rows of phase 1 are mostly on a
// peer: the TMA engine's deep asynchronous queue hides the NVLink latency that stalls LDGSTS issue;
// measured at 2 GPUs: 186 ms per pass against 261 ms with the kernel above).  One CTA per SM, two
// engines at once:
//
//  PHASE 1 through the TMA engine.  An item = 8 classes of a panel x 256 columns: the producer warp has
//  the TMA engine copy the 16 parent-row segments (2 KB each -- measured on B200: 1-D bulk copies cost
//  ~72 cycles per SM whatever their size below 1 KB and only reach HBM rate from 2 KB up, random
//  rows included) into one of NS1 shared-memory stages, completion on the stage's mbarrier, nothing
//  held in registers, the next stages already in flight.  NC1 consumer warps average father and
//  mother rows straight out of the stage -- lane = (class, column) -- and store the 64-byte pieces
//  N^T[c][8 classes] of the L2-resident panel.
//
//  PHASE 2 through the LSU.  NW2 autonomous warps, each with a private 17 KB stage: an item's 64
//  panel rows (256 bytes each, L2 hits) are staged with asynchronous 16-byte copies (LDGSTS, two
//  rows per instruction), then the warp produces tile (J, I) and its mirror straight from shared
//  memory.
//
// The two item streams are dealt round-robin over the CTAs and meet only in the per-panel completion
// counters: phase 2 of panel I waits for its n1 phase-1 items, phase 1 of panel I for phase 2 of
// panel I - ring_n (whose ring slot it overwrites).  Every wait is on strictly earlier work, so the
// persistent grid cannot deadlock.  No CTA-wide barrier in the loops.
__global__ void __launch_bounds__(kThreadsTma, 1) phi_step_tma_kernel(const StepDev s) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double *sprev[kMaxShards], *sout[kMaxShards];
    double *smem2 = reinterpret_cast<double *>(smem_raw);                                   // NW2 phase-2 stages
    unsigned char *smem1 = smem_raw + (size_t) NW2 * kStageDoubles * sizeof(double);        // NS1 phase-1 stages
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem1 + (size_t) NS1 * P1STAGE);   // full[NS1], empty[NS1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kMaxShards) { sprev[threadIdx.x] = s.prev.base[threadIdx.x]; sout[threadIdx.x] = s.outt.base[threadIdx.x]; }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS1; i++) { mbar_init(smem_u32(bars + i), 1); mbar_init(smem_u32(bars + NS1 + i), NC1 / NS1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // the all-zero rows (unknown parent) of every stage are never overwritten
    for (int i = threadIdx.x; i < NS1 * (P1PITCH / 8); i += blockDim.x)
        reinterpret_cast<double *>(smem1 + (size_t) (i / (P1PITCH / 8)) * P1STAGE + (size_t) kZeroSlot8 * P1PITCH)[i % (P1PITCH / 8)] = 0.0;
    for (int i = threadIdx.x; i < NW2 * PS; i += blockDim.x) smem2[(size_t) (i / PS) * kStageDoubles + kZeroSlot * PS + i % PS] = 0.0;
    __syncthreads();
    const RowLookup prev{ sprev, s.prev.base[0], s.prev.n > 1 ? s.prev.row0[1] : 1, s.prev.n, s.ld_prev };
    const unsigned long long pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    const long long n1_cta = s.total1 > blockIdx.x ? (s.total1 - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == NW2 + NC1) {
        // ============================ phase-1 producer (TMA) ====================================
        for (long long k = 0; k < n1_cta; k++) {
            const long long idx = blockIdx.x + k * gridDim.x;
            const int panel = s.panel_begin + (int) (idx / s.n1pp), rem = (int) (idx % s.n1pp);
            const int sg = rem / s.ncb256, cb = rem - sg * s.ncb256;
            const int w = min(C1, s.Kpad_prev - cb * C1);                 // columns of this block (multiple of 32)
            const int stage = (int) (k % NS1);
            const unsigned full = smem_u32(bars + stage), empty = smem_u32(bars + NS1 + stage);
            const int sub = panel * (T / S1) + sg;
            const int n = s.pcnt8[sub];                                   // distinct parent rows of the 8 classes
            const int row = lane < 2 * S1 ? s.prow8[sub * 2 * S1 + lane] : -1;
            if (lane == 0) {
                mbar_wait(empty, (unsigned) (((k / NS1) & 1) ^ 1));        // the consumers are done with this stage
                mbar_arrive_expect_tx(full, (unsigned) (n * w * sizeof(double)));
            }
            __syncwarp();
            if (lane < n)
                bulk_copy(smem_u32(smem1 + (size_t) stage * P1STAGE + (size_t) lane * P1PITCH), prev(row) + (long long) cb * C1,
                          (unsigned) (w * sizeof(double)), full, pol_stream);
        }
    } else if (warp >= NW2) {
        // ============================ phase-1 consumers =========================================
        const int i = lane & (S1 - 1), cc = lane >> 3;                    // lane = (class i of the item, column c mod 4)
        int freed_panel = -1;                    // phase 2 of this panel is known complete (its ring slot is free)
        // NC1 / NS1 consumer warps serve each stage (a warp only ever waits for the NEXT phase of its
        // stage's barrier); they split the item's columns between them
        constexpr int kPer = NC1 / NS1;
        const int cw = warp - NW2, my_stage = cw / kPer, part = cw % kPer;
        for (long long k = my_stage; k < n1_cta; k += NS1) {
            const long long idx = blockIdx.x + k * gridDim.x;
            const int panel = s.panel_begin + (int) (idx / s.n1pp), rem = (int) (idx % s.n1pp);
            const int sg = rem / s.ncb256, cb = rem - sg * s.ncb256;
            const int w = min(C1, s.Kpad_prev - cb * C1);
            const int stage = (int) (k % NS1);
            const unsigned full = smem_u32(bars + stage), empty = smem_u32(bars + NS1 + stage);
            const int old = panel - s.ring_n;    // the ring slot must have been drained by phase 2 of its previous panel
            if (old >= s.panel_begin && old > freed_panel) {
                if (lane == 0 && !s.debug_nowait) wait_count(s.done2 + old, s.full_rows ? s.np : old + 1);
                freed_panel = old;
                __syncwarp();
            }
            mbar_wait(full, (unsigned) ((k / NS1) & 1));
            const int sl = s.slots8[panel * T + sg * S1 + i];
            const double *rf = reinterpret_cast<const double *>(smem1 + (size_t) stage * P1STAGE + (size_t) (sl & 0xFF) * P1PITCH) + cc;
            const double *rm = reinterpret_cast<const double *>(smem1 + (size_t) stage * P1STAGE + (size_t) ((sl >> 8) & 0xFF) * P1PITCH) + cc;
            double *dst = s.ring + (long long) ((panel - s.panel_begin) % s.ring_n) * s.ring_stride +
                          ((long long) cb * C1 + cc) * T + sg * S1 + i;
            const int cbeg = part * (w / kPer), cend = cbeg + w / kPer;      // w is a multiple of 32
#pragma unroll 8
            for (int c = cbeg; c < cend; c += 4) st_hint(dst + (long long) c * T, 0.5 * (rf[c] + rm[c]), pol_keep);
            __syncwarp();
            if (lane == 0) { mbar_arrive(empty); signal_count(s.done1 + panel); }
        }
    } else {
        // ============================ phase-2 warps (LDGSTS) ====================================
        double *const S = smem2 + (size_t) warp * kStageDoubles;
        const int half = lane >> 4, l16 = lane & 15;
        // copy r: distinct rows 2r (lanes 0-15) and 2r + 1 (lanes 16-31) of the tile -> slots 2r, 2r + 1
        const unsigned sdst = smem_u32(S) + (unsigned) ((half * PS + l16 * 2) * sizeof(double));
        double *const out = s.out_base;              // own rows of the new matrix: global row r at out + (r - out_row0) * ld_new
        const long long ldn = s.ld_new;
        const unsigned long long pol_out = s.out_policy == 0 ? pol_stream : (s.out_policy == 2 ? pol_keep : policy_evict_normal());
        const long long nw2 = (long long) gridDim.x * NW2;
        int g = 0, ready_panel = -1;
        for (long long q = (long long) blockIdx.x * NW2 + warp; q < s.total2; q += nw2) {
            while (q >= s.grp_off2[g + 1]) ++g;
            const int I = s.panel_begin + g, J = (int) (q - s.grp_off2[g]);
            const int r0 = s.prow[J * 2 * T + lane], r1 = s.prow[J * 2 * T + T + lane], n = s.pcnt[J];
            const int slots = s.slots[J * T + lane];
            const double *buf = s.ring + (long long) ((I - s.panel_begin) % s.ring_n) * s.ring_stride;
            if (I != ready_panel) {              // phase 1 of panel I must be complete before its rows are read
                if (lane == 0 && !s.debug_nowait) wait_count(s.done1 + I, s.n1pp * (NC1 / NS1));
                ready_panel = I;
                __syncwarp();
            }
            const int nops = (n + 1) >> 1;
#pragma unroll 4
            for (int r = 0; r < nops; r++) {
                const int row = __shfl_sync(0xffffffffu, r < 16 ? r0 : r1, (2 * r + half) & 31);
                const double *src = row < 0 ? s.zero_row : buf + (long long) row * T + l16 * 2;
                cp_async16(sdst + (unsigned) (2 * r * PS * sizeof(double)), src, row >= 0 ? 16u : 0u, pol_keep);
            }
            const double2 *sf = reinterpret_cast<const double2 *>(S + (slots & 0xFF) * PS);
            const double2 *sm = reinterpret_cast<const double2 *>(S + ((slots >> 8) & 0xFF) * PS);
            // row j of tile J, columns of panel I: tile J may belong to another rank (ranked multi-GPU jobs evaluate
            // J <= I and store the mirror into the owner's rows: peer stores over NVLink)
            const RowLookup outl{ sout, s.outt.base[0], s.outt.n > 1 ? s.outt.row0[1] : 1, s.outt.n, ldn };
            double *drow = const_cast<double *>(outl(J * T)) + I * T + lane;
            double *mrow = out + ((long long) I * T - s.out_row0) * ldn + J * T + lane;    // row i of panel I, columns of tile J
            cp_async_wait_all();
            __syncwarp();
            if (J != I || s.full_rows) {
                // (tile J of another rank -- a ranked multi-GPU job evaluates J <= I only -- is not stored by this rank: its
                // owner pulls the transposed piece out of this rank's rows afterwards, pull_mirror_kernel)
                if (!s.full_rows && J >= s.panel_begin) {
#pragma unroll 8
                    for (int j = 0; j < T; j++) {
                        const int sl = __shfl_sync(0xffffffffu, slots, j);
                        st_hint(drow + (long long) j * ldn, 0.5 * (S[(sl & 0xFF) * PS + lane] + S[((sl >> 8) & 0xFF) * PS + lane]), pol_out);
                    }
                }
#pragma unroll 8
                for (int i2 = 0; i2 < T / 2; i2++) {
                    const double2 x = sf[i2], y = sm[i2];
                    st_hint(mrow + (long long) (2 * i2) * ldn, 0.5 * (x.x + y.x), pol_out);
                    st_hint(mrow + (long long) (2 * i2 + 1) * ldn, 0.5 * (x.y + y.y), pol_out);
                }
            } else {
                const double *sfl = reinterpret_cast<const double *>(sf), *sml = reinterpret_cast<const double *>(sm);
#pragma unroll 4
                for (int r = 0; r < T; r++) {
                    const int sl = __shfl_sync(0xffffffffu, slots, r);
                    const double a = 0.5 * (S[(sl & 0xFF) * PS + lane] + S[((sl >> 8) & 0xFF) * PS + lane]);   // V[r][lane]
                    const double b = 0.5 * (sfl[r] + sml[r]);                                                    // V[lane][r]
                    st_hint(mrow + (long long) r * ldn, r <= lane ? a : b, pol_out);
                }
            }
            __syncwarp();                                    // every lane is done reading the stage (and has issued its stores)
            if (lane == 0) signal_count(s.done2 + I);
        }
    }
}

// THIRD STEP KERNEL (one GPU, the default for large cuts).  What was measured on B200 to get here
// (scripts/micro/gather4_bench.cu, the GK_PROF role counters, ncu): reads that hit L2 run at 19 TB/s, DRAM reads at 7.3,
// stores at 5.9-7.5 TB/s, so the step is bound by DRAM time = reads of A + writes of B (they ADD: ~4.3 ms at c4); a 1-D
// TMA bulk copy costs ~105 cycles of the SM's TMA engine whatever its size, tile::gather4 ~50 cycles per row but packs the
// four rows at a 128-byte-aligned pitch (every row on the same banks), and per-item chains of dependent global loads or a
// gpu-scope fence on a critical warp cost more than the data movement itself.  Hence, one CTA per SM, four kinds of warps:
//
//  PRODUCER (1 warp): phase-1 items = one UNIT (16 classes = half a panel) x 512 columns of the previous matrix.  One bulk
//  copy (UBLKCP) per distinct parent row of the unit -- 4 KB each, ~20 per item -- into a 2-stage FIFO; staged rows are
//  (W + 2) * 8 bytes apart (an odd number of 16-byte units) at positions the planner chose so that the rows eight
//  consecutive classes read are distinct mod 8: bank-conflict-free 128-bit reads.  Items are decoded two ahead; the unit
//  of an item is found in a 32-entry window of the item prefix (one coalesced load per ~16 items).
//  CONSUMERS (4 warps): all work on the oldest full stage, lane = (class, column-pair parity): two 128-bit shared loads,
//  N = 1/2 (father + mother), 128-byte half rows N^T[c][16 classes] straight into the L2-resident panel ring (only the rows
//  some tile J <= I will read).  No transposition, no barrier wider than a warp; a finished item bumps a shared counter.
//  PHASE-2 WARPS (10): tiles (I, J <= I) are taken from a global queue in order, so the whole GPU drains the oldest panel.
//  Per class j the two 256-byte ring rows N^T[F_j], N^T[M_j] are loaded straight from L2 into registers (lane = class of
//  panel I; four batches of eight classes software-pipelined over two register buffers), row j of tile (J, I) is stored,
//  and the mirror tile (I, J) leaves through a 16 x 33 scratch as 128-byte pieces.  Finished tiles go to a mailbox.
//  PUBLISHER (1 warp): snapshots the consumers' counters and the mailboxes, executes ONE fence.acq_rel.gpu, then updates
//  done1 / done2 -- the release / acquire chain worker -> shared memory -> publisher -> fence carries the workers' stores,
//  and no worker ever stalls on a membar.
//
// Dependencies as in the first kernel: phase 2 of panel I waits for done1[I]; phase 1 of panel I for done2 of panel
// I - ring_n; every wait is on strictly earlier work and all CTAs are resident.
constexpr int G4NS = kG4Stages, G4NC = kG4Consumers, G4P2 = kG4P2Warps;
constexpr int kThreadsG4 = (G4NC + G4P2 + 2) * 32;        // consumers, phase-2 warps, the producer, the publisher
constexpr int G4SCRP = 33;                                 // scratch pitch (doubles)
constexpr int G4SCR = 16 * G4SCRP;                         // scratch doubles per warp
constexpr size_t kStepSmemG4 = (size_t) G4NS * kG4StageBytes + (size_t) G4P2 * G4SCR * sizeof(double) + 2 * G4NS * sizeof(unsigned long long) + 64;

__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void signal_add(int *p, int v) {
    asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(kThreadsG4, 1) phi_step_g4_kernel(const StepDev s) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int spub[8][2];                           // (panel, blocks) of the last 8 items issued, for the publisher
    __shared__ volatile int ctl[G4NC + 1];               // items consumed by each consumer warp; items published
    __shared__ volatile int p2box[G4P2][2];              // phase-2 warp -> publisher: (panel << 8 | tiles, sequence number)
    __shared__ volatile int p2ack[G4P2];                 // sequence number the publisher has published
    __shared__ volatile int p2live;                      // phase-2 warps still running
    unsigned char *stages = smem_raw;
    double *scratch = reinterpret_cast<double *>(smem_raw + (size_t) G4NS * kG4StageBytes);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(scratch + (size_t) G4P2 * G4SCR);   // full[NS], empty[NS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x <= G4NC) ctl[threadIdx.x] = 0;
    if (threadIdx.x < G4P2) { p2box[threadIdx.x][0] = 0; p2box[threadIdx.x][1] = 0; p2ack[threadIdx.x] = 0; }
    if (threadIdx.x == 0) p2live = G4P2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < G4NS; i++) { mbar_init(smem_u32(bars + i), 1); mbar_init(smem_u32(bars + G4NS + i), G4NC); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    __shared__ double *sprev[kMaxShards];
    if (threadIdx.x < kMaxShards) sprev[threadIdx.x] = s.prev.base[threadIdx.x];
    __syncthreads();
    const RowLookup prev{ sprev, s.prev.base[0], s.prev.n > 1 ? s.prev.row0[1] : 1, s.prev.n, s.ld_prev };
    const unsigned long long pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    // debug_nowait (timing experiments only, results are garbage): bit 0 = skip the dependency waits, bit 1 = no phase 1, bit 2 = no phase 2
    const long long n_cta = (s.debug_nowait & 2) ? 0 : (s.total_g4 > blockIdx.x ? (s.total_g4 - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    const long long total2 = (s.debug_nowait & 4) ? 0 : s.total2;
    const int nblk_panel = 2 * (s.Kpad_prev / T);        // blocks (16 classes x 32 columns) of a panel = its done1 target

    // item t of this launch -> (unit, column block): the units' item prefix is searched 32 entries at a time (one coalesced
    // load per ~16 items instead of a chain of dependent loads per item)
    struct ItemCursor {
        int base = 0;                 // unit of the window's first entry
        long long lo = 0, hi = 0;     // lo = item_off[base]; lane l holds hi = item_off[base + 1 + l]
        bool loaded = false;
    };
    auto locate = [&](ItemCursor &c, const long long t, const int nunits, int &unit, int &cb) {
        for (;;) {
            if (!c.loaded) {
                c.lo = s.item_off[c.base];
                const int u = c.base + 1 + lane;
                c.hi = u <= nunits ? s.item_off[u] : 0x7fffffffffffffffLL;
                c.loaded = true;
            }
            const unsigned m = __ballot_sync(0xffffffffu, c.hi > t);
            if (m) {
                const int l = __ffs(m) - 1;
                const long long start = l ? __shfl_sync(0xffffffffu, c.hi, l - 1) : c.lo;
                unit = c.base + l; cb = (int) (t - start);
                return;
            }
            c.base += 32; c.loaded = false;
        }
    };
    const int nunits = 2 * s.np;

    if (warp == G4NC + G4P2) {
        // ============================ producer: one bulk copy per staged row ===========================
        const long long pt0 = clock64();
        long long p_stage = 0, p_ring = 0;
        ItemCursor cur;
        // descriptors are decoded two items ahead (their global loads -- prefix window, row list -- land before they are needed)
        int dU[2] = { 0, 0 }, dcb[2] = { 0, 0 }, dn[2] = { 0, 0 }, dr0[2] = { 0, 0 }, dr1[2] = { 0, 0 };
        auto decode = [&](const long long k, const int q) {
            if (k >= n_cta) return;
            int unit, cb;
            locate(cur, blockIdx.x + k * gridDim.x, nunits, unit, cb);
            dU[q] = unit; dcb[q] = cb; dn[q] = s.pcnt4[unit];
            const int32_t *pr = s.prow4 + (long long) unit * kG4RowsMax;
            dr0[q] = pr[lane]; dr1[q] = lane < kG4RowsMax - 32 ? pr[32 + lane] : -1;
        };
        decode(0, 0);
        decode(1, 1);
        int checked_panel = -1;
#pragma unroll 1
        for (long long k0 = 0; k0 < n_cta; k0 += 2) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const long long k = k0 + q;
                if (k >= n_cta) break;
                const int U = dU[q], cb = dcb[q], I = U >> 1;
                const int npos = dn[q] & 0xFF, nrows = (dn[q] >> 8) & 0xFF;     // staged positions (some unused), rows copied
                const int stage = (int) (k % G4NS);
                if (lane == 0) {
                    // the stage's previous item must have been consumed by every consumer warp, and its publication slot freed
                    if (k >= G4NS) {
                        const long long c0 = clock64();
                        for (;;) {
                            int mn = ctl[0];
#pragma unroll
                            for (int w = 1; w < G4NC; w++) mn = min(mn, ctl[w]);
                            if (mn > (int) (k - G4NS) && (int) k - ctl[G4NC] < 8) break;
                            __nanosleep(32);
                        }
                        p_stage += clock64() - c0;
                    }
                    if (I != checked_panel) {              // first item of a panel here: its ring slot must have been drained
                        const int old = I - s.ring_n;
                        const long long c0 = clock64();
                        if (old >= s.panel_begin && !(s.debug_nowait & 1)) wait_count(s.done2 + old, s.full_rows ? s.np - (s.panel_end - 1 - old) : old + 1);
                        p_ring += clock64() - c0;
                    }
                }
                checked_panel = I;
                __syncwarp();
                const int W = (dn[q] >> 16) ? 512 : 256;
                const int w = min(W, s.Kpad_prev - cb * W);                    // columns of this item (a multiple of 32)
                const unsigned full = smem_u32(bars + stage);
                if (lane == 0) {
                    spub[k & 7][0] = I; spub[k & 7][1] = w / T;
                    __threadfence_block();
                    mbar_arrive_expect_tx(full, (unsigned) (nrows * w * sizeof(double)));
                }
                __syncwarp();
                const unsigned pitch = (unsigned) (W + 2) * 8u;
                const unsigned sbase = smem_u32(stages + (size_t) stage * kG4StageBytes);
                // one bulk copy per staged row: 4 KB (2 KB) segments -- the TMA engine spends ~105 cycles per copy whatever its size
                // (row Kpad_prev = the all-zero row of an unknown parent: a small buffer of zeros, whatever the column block)
                if (lane < npos && dr0[q] >= 0)
                    bulk_copy(sbase + (unsigned) lane * pitch, dr0[q] < s.Kpad_prev ? prev(dr0[q]) + (long long) cb * W : s.zero_row,
                              (unsigned) (w * sizeof(double)), full, pol_stream);
                if (lane + 32 < npos && dr1[q] >= 0)
                    bulk_copy(sbase + (unsigned) (lane + 32) * pitch, dr1[q] < s.Kpad_prev ? prev(dr1[q]) + (long long) cb * W : s.zero_row,
                              (unsigned) (w * sizeof(double)), full, pol_stream);
                decode(k + 2, q);
            }
        }
        if (s.prof && lane == 0) {
            atomicAdd(s.prof + 0, (unsigned long long) (clock64() - pt0));
            atomicAdd(s.prof + 2, (unsigned long long) p_stage); atomicAdd(s.prof + 3, (unsigned long long) p_ring);
        }
    } else if (warp == G4NC + G4P2 + 1) {
        // ============================ publisher: makes consumed items visible to the other SMs =========
        // (one gpu-scope fence per item, off the producer's and the consumers' critical paths: the consumers only bump a
        // shared-memory counter, the release/acquire chain consumer -> counter -> this warp -> fence.gpu carries their stores)
        // One loop iteration = snapshot of everything finished (phase-1 items consumed by all consumer warps, phase-2
        // mailboxes), ONE fence, then the counter updates: the fence rate does not grow with the number of messages.
        if (lane == 0) {
            long long k = 0;                         // next phase-1 item to publish
            int seen[G4P2];
#pragma unroll
            for (int w = 0; w < G4P2; w++) seen[w] = 0;
            for (;;) {
                const int live = p2live;
                int mn = ctl[0];
#pragma unroll
                for (int w = 1; w < G4NC; w++) mn = min(mn, ctl[w]);
                int msg[G4P2], seq[G4P2];
                bool any = mn > (int) k;
#pragma unroll
                for (int w = 0; w < G4P2; w++) { seq[w] = p2box[w][1]; msg[w] = p2box[w][0]; any |= seq[w] != seen[w]; }
                if (any) {
                    __threadfence_block();
                    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                    for (; k < mn; k++) {
                        asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;\n" ::"l"(s.done1 + spub[k & 7][0]), "r"(spub[k & 7][1]) : "memory");
                    }
                    ctl[G4NC] = (int) k;
#pragma unroll
                    for (int w = 0; w < G4P2; w++)
                        if (seq[w] != seen[w]) {
                            asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;\n" ::"l"(s.done2 + (msg[w] >> 8)), "r"(msg[w] & 0xFF) : "memory");
                            seen[w] = seq[w];
                            p2ack[w] = seq[w];
                        }
                } else {
                    if (k >= n_cta && live == 0) break;
                    __nanosleep(32);
                }
            }
        }
    } else if (warp < G4NC) {
        // ============================ phase-1 consumers ===============================================
        // The stages are a FIFO: the producer runs up to kG4Stages items ahead, and ALL consumer warps work on the oldest
        // full stage together, each on its own 32-column blocks, so loading an item overlaps with consuming the previous ones.
        const long long ct0 = clock64();
        long long c_wait = 0;
        const int cls = lane & 15, par = lane >> 4;       // lane = (class of the unit, parity of the column pair)
        // the next item, decoded one iteration ahead (its global loads fly while the current item is consumed)
        int nU = 0, ncb = 0, nW = 0, npk = 0;
        ItemCursor cur;
        auto decode_next = [&](long long k) {
            if (k >= n_cta) return;
            locate(cur, blockIdx.x + k * gridDim.x, nunits, nU, ncb);
            nW = (s.pcnt4[nU] >> 16) ? 512 : 256;
            npk = s.slots4[nU * 16 + cls];
        };
        decode_next(0);
        for (long long k = 0; k < n_cta; k++) {
            const int stage = (int) (k % G4NS);
            const unsigned S = smem_u32(stages + (size_t) stage * kG4StageBytes);
            const unsigned full = smem_u32(bars + stage);
            const int U = nU, cb = ncb, W = nW, pk = npk;
            const int I = U >> 1;
            const int w = min(W, s.Kpad_prev - cb * W);
            const int nb64 = (w + 63) >> 6;
            double *buf = s.ring + (long long) ((I - s.panel_begin) % s.ring_n) * s.ring_stride + (U & 1) * 16 + cls;
            // lane = class: its father / mother rows in the stage.  Staged rows are (W + 2) * 8 bytes apart -- an odd number
            // of 16-byte units: the 128-bit reads of 8 consecutive classes (whose rows are mostly distinct mod 8: the
            // planner numbers them in order of first use) do not collide on a bank.
            const unsigned pitch = (unsigned) (W + 2) * 8u;
            const unsigned aF = S + (unsigned) (pk & 0xFF) * pitch + par * 16u, aM = S + (unsigned) ((pk >> 8) & 0xFF) * pitch + par * 16u;
            // rows of the panel some tile J <= I reads, for this warp's (at most two) blocks of 64 columns
            int fu[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = cb * W + (warp + (q >> 1) * G4NC) * 64 + (q & 1) * 32 + lane;
                fu[q] = c < s.Kpad_prev ? s.first_use[c] : 0x7fffffff;
            }
            decode_next(k + 1);
            const long long cw0 = clock64();
            mbar_wait(full, (unsigned) ((k / G4NS) & 1));
            c_wait += clock64() - cw0;
#pragma unroll
            for (int bi = 0; bi < 2; bi++) {
                const int b = warp + bi * G4NC;
                if (b < nb64) {
                    const int c0 = cb * W + b * 64;
                    const unsigned need_lo = __ballot_sync(0xffffffffu, s.full_rows ? fu[2 * bi] != 0x7fffffff : fu[2 * bi] <= I),
                                   need_hi = __ballot_sync(0xffffffffu, s.full_rows ? fu[2 * bi + 1] != 0x7fffffff : fu[2 * bi + 1] <= I);
                    double *dst = buf + (long long) c0 * T;
                    const unsigned bF = aF + (unsigned) b * 512u, bM = aM + (unsigned) b * 512u;
#pragma unroll
                    for (int u0 = 0; u0 < 16; u0 += 4) {
                        double2 x[4], y[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) { x[u] = lds_f64x2(bF + (u0 + u) * 32); y[u] = lds_f64x2(bM + (u0 + u) * 32); }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int c = 4 * (u0 + u) + 2 * par;                     // column of x[u].x inside the block
                            const unsigned need = c < 32 ? need_lo >> c : need_hi >> (c - 32);
                            if (need & 1u) st_hint(dst + c * T, 0.5 * (x[u].x + y[u].x), pol_keep);
                            if (need & 2u) st_hint(dst + (c + 1) * T, 0.5 * (x[u].y + y[u].y), pol_keep);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); ctl[warp] = (int) k + 1; }
        }
        if (s.prof && lane == 0) { atomicAdd(s.prof + 4, (unsigned long long) (clock64() - ct0)); atomicAdd(s.prof + 5, (unsigned long long) c_wait); }
    } else {
        // ============================ phase-2 warps (L2 -> registers) ==================================
        const int w2 = warp - G4NC;
        double *scr = scratch + (size_t) w2 * G4SCR;
        double *const out = s.out_base;
        const long long ldn = s.ld_new;
        const unsigned long long pol_out = s.out_policy == 0 ? pol_stream : (s.out_policy == 2 ? pol_keep : policy_evict_normal());
        const int hi = lane >> 4, lo = lane & 15;
        const bool nostore = (s.debug_nowait & 8) != 0;
        // finished tiles go to the publisher warp through a shared-memory mailbox (lane 0 only): the release / acquire chain
        // this warp -> mailbox -> publisher -> fence.gpu makes the warp's stores visible before the counter moves
        int myseq = 0;
        auto post = [&](const int panel, const int count) {
            while (p2ack[w2] != myseq) __nanosleep(32);        // the previous message has been taken
            __threadfence_block();
            p2box[w2][0] = (panel << 8) | count;
            __threadfence_block();
            p2box[w2][1] = ++myseq;
        };
        const long long nw2 = (long long) gridDim.x * G4P2;
        int g = 0, ready_panel = -1;
        const long long qt0 = clock64();
        long long q_wait = 0, q_sig = 0;
        // tiles are dealt in chunks of kG4Chunk consecutive queue items (mostly one panel I, tiles J, J + 1, ...): one
        // gpu-scope fence and one counter update per chunk and panel instead of one per tile
        const int CH = s.chunk2 > 0 ? s.chunk2 : 1;
        const long long nchunks = (total2 + CH - 1) / CH;
        // chunks are taken from a global counter in queue order, so every phase-2 warp of the GPU works on the OLDEST panel
        // that still has tiles (a static deal left most warps waiting for panels phase 1 had not reached yet while the
        // ready panel's tiles sat with a few busy warps, and phase 1 stalled on the ring slot that panel occupied)
        for (;;) {
            long long cq = 0;
            if (lane == 0) cq = (long long) atomicAdd(s.queue2, 1ULL);
            cq = __shfl_sync(0xffffffffu, cq, 0);
            if (cq >= nchunks) break;
            int sig_panel = -1, sig_count = 0;
            const long long q_end = (cq + 1) * CH < total2 ? (cq + 1) * CH : total2;
            for (long long q = cq * CH; q < q_end; q++) {
                while (q >= s.grp_off2[g + 1]) ++g;
                const int I = s.panel_begin + g;
                int J = (int) (q - s.grp_off2[g]);
                // a rank of a sharded job: every tile J except those of its own panels above the diagonal (tile (I, J < I) of the
                // own block stores both pieces, as on one GPU)
                if (s.full_rows && J > I) J += s.panel_end - 1 - I;
                const double *buf = s.ring + (long long) ((I - s.panel_begin) % s.ring_n) * s.ring_stride + lane;
                const int pf = s.pF[J * T + lane], pm = s.pM[J * T + lane];
                if (I != sig_panel && sig_count > 0) {    // the chunk crosses into the next panel: publish the tiles of the previous one
                    __syncwarp();
                    const long long s0 = clock64();
                    if (lane == 0) post(sig_panel, sig_count);
                    __syncwarp();
                    q_sig += clock64() - s0;
                    sig_count = 0;
                }
                sig_panel = I;
                if (I != ready_panel) {              // phase 1 of panel I must be complete before its rows are read
                    const long long w0 = clock64();
                    if (lane == 0 && !s.debug_nowait) wait_count(s.done1 + I, nblk_panel);
                    ready_panel = I;
                    __syncwarp();
                    q_wait += clock64() - w0;
                }
                double *drow = out + ((long long) J * T - s.out_row0) * ldn + I * T + lane;    // row j of tile J, columns of panel I
                double *mrow = out + ((long long) I * T - s.out_row0) * ldn + J * T;           // row i of panel I, columns of tile J
                // rows of tile J are stored only if they are this rank's: a sharded job stores the transposed pieces into its own rows
                // and, inside its own block (J < I), the rows of tile J as well; ranked sharded jobs evaluate J <= I and the owner
                // of another rank's J pulls the piece out of this rank's rows (pull_mirror_kernel)
                const bool own_j = J >= s.panel_begin && (!s.full_rows || J < I);
                if (J != I || s.full_rows) {
                    // (a rank of a sharded job evaluates every tile J for its own panels and stores only its own rows: the
                    // transposed pieces)
                    // Four batches of eight classes, software-pipelined over two register buffers: the loads of batch b + 1
                    // are in flight while batch b is averaged and stored (a warp alone otherwise sits through one L2
                    // latency per batch).  No predicates: an unknown parent reads a row of zeros.
                    double xa[8], ya[8], xb[8], yb[8];
                    const double *zrow = s.zero_row + lane;              // the row of an unknown parent
                    auto issue = [&](double (&x)[8], double (&y)[8], const int bt) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int rf = __shfl_sync(0xffffffffu, pf, 8 * bt + i), rm = __shfl_sync(0xffffffffu, pm, 8 * bt + i);
                            x[i] = __ldcg(rf >= 0 ? buf + (long long) rf * T : zrow);
                            y[i] = __ldcg(rm >= 0 ? buf + (long long) rm * T : zrow);
                        }
                    };
                    auto consume = [&](double (&x)[8], double (&y)[8], const int bt) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const double v = 0.5 * (x[i] + y[i]);
                            if (!nostore && own_j) st_hint(drow + (long long) (8 * bt + i) * ldn, v, pol_out);
                            scr[((8 * bt + i) & 15) * G4SCRP + lane] = v;
                        }
                    };
                    // mirror of 16 rows j: rows 2p (lanes 0-15) and 2p + 1 (lanes 16-31) of panel I, columns 16h .. 16h + 15 of tile J
                    auto mirror = [&](const int h) {
                        __syncwarp();
#pragma unroll
                        for (int p0 = 0; p0 < 16; p0 += 8) {
                            double v[8];
#pragma unroll
                            for (int p = 0; p < 8; p++) v[p] = scr[lo * G4SCRP + 2 * (p0 + p) + hi];
#pragma unroll
                            for (int p = 0; p < 8; p++) if (!nostore || v[p] == 12345.678) st_hint(mrow + (long long) (2 * (p0 + p) + hi) * ldn + 16 * h + lo, v[p], pol_out);
                        }
                        __syncwarp();
                    };
                    issue(xa, ya, 0);
                    issue(xb, yb, 1);
                    consume(xa, ya, 0);
                    issue(xa, ya, 2);
                    consume(xb, yb, 1);
                    issue(xb, yb, 3);
                    mirror(0);
                    consume(xa, ya, 2);
                    consume(xb, yb, 3);
                    mirror(1);
                } else {
                    // diagonal tile (one in I + 1): V[j][i] is valid where j <= i (in RANKED mode that is the reference's grouping);
                    // the other triangle is its mirror
#pragma unroll 4
                    for (int j = 0; j < T; j++) {
                        const int rf = __shfl_sync(0xffffffffu, pf, j), rm = __shfl_sync(0xffffffffu, pm, j);
                        const double x = rf >= 0 ? __ldcg(buf + (long long) rf * T) : 0.0, y = rm >= 0 ? __ldcg(buf + (long long) rm * T) : 0.0;
                        const double v = 0.5 * (x + y);
                        if (j <= lane) st_hint(drow + (long long) j * ldn, v, pol_out);
                        if (j < lane) st_hint(mrow + (long long) lane * ldn + j, v, pol_out);
                    }
                }
                sig_count++;
            }
            __syncwarp();
            const long long s0 = clock64();
            if (lane == 0 && sig_count > 0) post(sig_panel, sig_count);
            __syncwarp();
            q_sig += clock64() - s0;
        }
        __syncwarp();
        if (lane == 0) { while (p2ack[w2] != myseq) __nanosleep(32); atomicSub((int *) &p2live, 1); }
        if (s.prof && lane == 0) {
            atomicAdd(s.prof + 7, (unsigned long long) (clock64() - qt0)); atomicAdd(s.prof + 8, (unsigned long long) q_wait);
            atomicAdd(s.prof + 10, (unsigned long long) q_sig);
        }
    }
}

__device__ __forceinline__ double *row_ptr(const RowTable &t, int row, long long ld) {
    double *base = t.base[0];
    int r0 = 0;
#pragma unroll
    for (int q = 1; q < kMaxShards; q++)
        if (q < t.n && row >= t.row0[q]) { base = t.base[q]; r0 = t.row0[q]; }
    return base + (long long) (row - r0) * ld;
}

// Self kinship of the new classes (lib/compute.cpp:553-568) and the "same individual" corrections
// (:522-528) on top of the step kernel's B = P A P^T.
__global__ void __launch_bounds__(256) phi_diag_kernel(const DiagDev d) {
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int c = d.c_begin + t0; c < d.c_end; c += stride) {
        const int F = d.pF[c], M = d.pM[c], fl = d.flags[c];
        double self = 0.5;
        if (fl & kFlagKept) self = d.dprev[F];                      // bitwise what the reference recomputes
        else if (fl & kFlagBoth) self = 0.5 + 0.5 * row_ptr(d.prev, F, d.ld_prev)[M];
        d.dnew[c] = self;
        for (int q = 0; q < d.n_peer; q++) d.dnew_peer[q][c] = self;
        if (fl & kFlagSingleton) row_ptr(d.out, c, d.ld_new)[c] = self;
    }
    for (int gidx = t0; gidx < d.ngroups; gidx += stride) {
        const int i = d.pg_i[gidx], j = d.pg_j[gidx];
        const bool own_i = i >= d.c_begin && i < d.c_end, own_j = j >= d.c_begin && j < d.c_end;
        if (!own_i && !own_j) continue;
        double add = 0.0;
        for (int k = d.pg_off[gidx]; k < d.pg_off[gidx + 1]; k++) {
            const int c = d.pt_cls[k];
            add += 0.25 * (double) d.pt_mult[k] * (d.dprev[c] - row_ptr(d.prev, c, d.ld_prev)[c]);
        }
        if (own_i) { double *p = row_ptr(d.out, i, d.ld_new) + j; *p = *p + add; }
        if (own_j && i != j) { double *p = row_ptr(d.out, j, d.ld_new) + i; *p = *p + add; }
    }
}

// Founders' matrix (lib/compute.cpp:655-661): classes of cut 0 are unrelated; self kinship 1/2.
// A class of several founders keeps 0 on its diagonal (two DISTINCT members), a singleton 1/2.
__global__ void init_kernel(double *G0, double *d0, const int32_t *flags, long long K0, long long Kpad0) {
    const long long total = Kpad0 * Kpad0;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total;
         i += (long long) gridDim.x * blockDim.x) {
        const long long r = i / Kpad0, c = i - r * Kpad0;
        G0[i] = (r == c && r < K0 && (flags[r] & kFlagSingleton)) ? 0.5 : 0.0;
    }
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < Kpad0;
         i += (long long) gridDim.x * blockDim.x)
        d0[i] = 0.5;
}

// Multi-GPU start of an uncompressed job: the rows [row_begin, row_begin + rows) of the founders'
// matrix are zero (memset by the caller) except the diagonal; d0 is replicated on every rank.
__global__ void init_rows_kernel(double *G0, double *d0, const int32_t *flags, long long K0, long long Kpad0,
                                 long long row_begin, long long rows, long long dlen) {
    const long long stride = (long long) gridDim.x * blockDim.x, t0 = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    for (long long r = t0; r < rows; r += stride) {
        const long long c = row_begin + r;
        if (c < K0 && (flags[c] & kFlagSingleton)) G0[r * Kpad0 + c] = 0.5;
    }
    for (long long i = t0; i < dlen; i += stride) d0[i] = 0.5;
}

// Ranked multi-GPU jobs evaluate the tiles J <= I only (the reference's rounding order) and every rank stores into its OWN
// rows.  The tiles (J own, I of a higher rank) of this rank's rows are the transposes of tiles the owners of I computed into
// THEIR rows: read them out of the owners' HBM (256-byte row segments over NVLink), transpose through shared memory, store
// into the own rows.  One warp per tile; runs after the barrier that follows the step, and a second barrier follows it.
__global__ void __launch_bounds__(128) pull_mirror_kernel(const RowTable t, long long ld, int pb, int pe, int np, double *own, long long own_row0) {
    __shared__ double tile[4][T][T + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // unit of work = one panel I of a higher rank x 8 consecutive own panels J: the warp walks along the 32 source rows
    // (sequential 256-byte pieces of the same remote pages), and the warps of a CTA share those rows
    constexpr int kRun = 8;
    const long long runs = (pe - pb + kRun - 1) / kRun, nunits = (long long) (np - pe) * runs;
    for (long long u = blockIdx.x * 4LL + warp; u < nunits; u += gridDim.x * 4LL) {
        const int Ip = pe + (int) (u / runs), J0 = pb + (int) (u % runs) * kRun, J1 = min(J0 + kRun, pe);
        const double *src[T];
#pragma unroll
        for (int k = 0; k < T; k++) src[k] = row_ptr(t, Ip * T + k, ld) + lane;
        for (int Jp = J0; Jp < J1; Jp++) {
            double v[T];
#pragma unroll
            for (int k = 0; k < T; k++) v[k] = __ldcg(src[k] + (long long) Jp * T);
#pragma unroll
            for (int k = 0; k < T; k++) tile[warp][k][lane] = v[k];
            __syncwarp();
            double *dst = own + ((long long) Jp * T - own_row0) * ld + (long long) Ip * T + lane;
#pragma unroll 8
            for (int k = 0; k < T; k++) __stcs(dst + (long long) k * ld, tile[warp][lane][k]);
            __syncwarp();
        }
    }
}

// Multi-GPU expansion, even m: the class rows this rank's probands need, fetched from their owners TRANSPOSED --
// GT[c][a] = G[rows[a]][c] for the active classes a in [a_begin, a_begin + a_count) and every class column c -- so that
// expand_sym_kernel finds the values of 8 consecutive active classes for a column in one 64-byte segment.  Same access
// order as pull_mirror_kernel: a warp walks 8 column panels along its 32 source rows.
__global__ void __launch_bounds__(128) fetch_rows_t_kernel(const RowTable t, long long ld, const int32_t *rows, int a_begin, int a_count,
                                                           int nact, int ncp /* column panels */, double *GT, long long ldt) {
    __shared__ double tile[4][T][T + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kRun = 8;
    const long long runs = (ncp + kRun - 1) / kRun, nblk = (a_count + T - 1) / T, nunits = nblk * runs;
    for (long long u = blockIdx.x * 4LL + warp; u < nunits; u += gridDim.x * 4LL) {
        const int ab = a_begin + (int) (u / runs) * T, J0 = (int) (u % runs) * kRun, J1 = min(J0 + kRun, ncp);
        const double *src[T];
#pragma unroll
        for (int k = 0; k < T; k++) src[k] = row_ptr(t, rows[min(ab + k, nact - 1)], ld) + lane;     // (past the end: any valid row, never used)
        for (int Jp = J0; Jp < J1; Jp++) {
            double v[T];
#pragma unroll
            for (int k = 0; k < T; k++) v[k] = __ldcg(src[k] + (long long) Jp * T);
#pragma unroll
            for (int k = 0; k < T; k++) tile[warp][k][lane] = v[k];
            __syncwarp();
            double *dst = GT + (long long) Jp * T * ldt + ab + lane;
            if (ab + lane < a_begin + a_count) {
#pragma unroll 8
                for (int k = 0; k < T; k++) dst[(long long) k * ldt] = tile[warp][lane][k];
            }
            __syncwarp();
        }
    }
}

// Multi-GPU expansion: copy the class rows this rank's probands need from their owners (peer
// reads over NVLink, 16-byte vectorised, one CTA per row) into a compact local matrix.
__global__ void __launch_bounds__(256) fetch_rows_kernel(const RowTable t, long long ld, const int32_t *rows, double *dst, int first) {
    const int a = first + (int) blockIdx.x;             // position in the list of active classes
    const double2 *src = reinterpret_cast<const double2 *>(row_ptr(t, rows[a], ld));
    double2 *out = reinterpret_cast<double2 *>(dst + (long long) a * ld);
    for (long long c = threadIdx.x; c < ld / 2; c += 256) out[c] = src[c];
}

// Final expansion to the caller's proband order (duplicates, ancestors-as-probands) fused with
// the gen.phiMean reduction (geneakit/compute.py:124-132).  Probands of one class (sibship) have
// identical rows, so one CTA gathers a 1024-column slice of the class row ONCE (8-byte gathers,
// the row stays in L2) and then streams it, coalesced, to every local result row of that class,
// patching the "same individual" entries with the self kinship.  Each CTA also emits one
// (sum, diagonal sum) partial for the mean.
constexpr int kExpCPT = kExpandCols / 256;     // result columns per thread
constexpr int kExpMembers = 64;                // result rows of a class whose (row, individual) are staged in shared memory
__global__ void __launch_bounds__(256) expand_kernel(const ExpandDev e) {
    __shared__ int mem_i[kExpMembers], mem_u[kExpMembers];
    const int a = e.a_begin + blockIdx.x / e.nchunks, chunk = blockIdx.x % e.nchunks;     // (a launch covers a range of the active classes)
    const int c = e.cact[a];
    const long long j0 = (long long) chunk * kExpandCols + threadIdx.x;
    const int q0 = e.coff[a], nm = e.coff[a + 1] - q0;
    if (threadIdx.x < nm && threadIdx.x < kExpMembers) {      // independent of the gathers below: both chains overlap
        const int i = e.cmem[q0 + threadIdx.x];
        mem_i[threadIdx.x] = i; mem_u[threadIdx.x] = e.ind[i];
    }
    double v[kExpCPT];
    int uj[kExpCPT];
    const double *row = e.G + (long long) (e.compact ? a : c) * e.ld;
#pragma unroll
    for (int k = 0; k < kExpCPT; k++) {
        const long long j = j0 + 256 * k;
        const bool ok = j < e.m;
        uj[k] = ok ? e.ind[j] : -2;
        v[k] = ok ? row[e.cls[j]] : 0.0;
    }
    const double dc = e.d[c];
    double sum = 0.0, dsum = 0.0;
    __syncthreads();
    for (int q = 0; q < nm; q++) {
        long long i; int ui;
        if (q < kExpMembers) { i = mem_i[q]; ui = mem_u[q]; } else { i = e.cmem[q0 + q]; ui = e.ind[i]; }
        double *orow = e.out + (i - e.row0) * e.m;
#pragma unroll
        for (int k = 0; k < kExpCPT; k++) {
            const long long j = j0 + 256 * k;
            if (j < e.m) {
                const double val = (uj[k] == ui) ? dc : v[k];
                __stcs(orow + j, val);
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
        e.partials[2 * ((long long) a * e.nchunks + chunk)] = x;
        e.partials[2 * ((long long) a * e.nchunks + chunk) + 1] = y;
    }
}

// The same expansion for ONE GPU, where every class row is local: the class matrix is symmetric, so the eight values
// G[a0 .. a0 + 7][cls_j] that eight consecutive classes need for column j are ONE contiguous 64-byte segment of row cls_j --
// G[cls_j][a0 .. a0 + 7].  A CTA = 8 classes x kExpandSymCols result columns; lane = (column j, 16-byte part of the segment):
// a warp load touches 8 segments instead of 32 scattered 8-byte words (the gather rate, one divergent line per clock per
// SM, was the bound of the kernel above), and every result row of those classes is written in 64-byte pieces.  CTAs of the
// same 8 classes are adjacent in the grid: the 64-byte column block they read stays in L2 (2.8 MB at c4).
constexpr int kExpSymMembers = 32;             // result rows per class staged in shared memory
constexpr int kExpSymIter = 256;               // columns per iteration: two pairs per thread
constexpr int kExpSymCols = 2048;              // result columns per CTA
__global__ void __launch_bounds__(256, 3) expand_sym_kernel(const ExpandDev e) {
    __shared__ int s_n[8], s_q0[8];
    __shared__ int s_i[8][kExpSymMembers], s_u[8][kExpSymMembers];
    const int nchunks = (int) ((e.m + kExpSymCols - 1) / kExpSymCols);
    const int A = e.a_begin / 8 + blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;      // (a launch covers a range of the active classes)
    const int a0 = A * 8, K = e.nact;
    if (threadIdx.x < 8) {
        const int c = a0 + threadIdx.x;
        s_q0[threadIdx.x] = c < K ? e.coff[c] : 0;
        s_n[threadIdx.x] = c < K ? e.coff[c + 1] - e.coff[c] : 0;
    }
    __syncthreads();
    double sum = 0.0, dsum = 0.0;
    {
        // warp = class: its result rows (and their individuals) into shared memory; the diagonal entries out[i][i] that fall
        // into this CTA's columns are the self kinship of the class
        const int cl = threadIdx.x >> 5;
        for (int k = threadIdx.x & 31; k < s_n[cl]; k += 32) {
            const int i = e.cmem[s_q0[cl] + k];
            if (k < kExpSymMembers) { s_i[cl][k] = i; s_u[cl][k] = e.ind[i]; }
            if (i / kExpSymCols == chunk) dsum += e.d[e.cact[a0 + cl]];
        }
    }
    __syncthreads();
    // lane = (pair of adjacent columns, 16-byte part of the segment): the 8 lanes of a class pair write 128 contiguous bytes of
    // a result row per instruction (m is even: the rows are 16-byte aligned)
    const int r = threadIdx.x & 3, q = threadIdx.x >> 2;
    const int ca = 2 * r, cb = 2 * r + 1;                                // this thread's two classes (within the block of 8)
    const double da = a0 + ca < K ? e.d[e.cact[a0 + ca]] : 0.0, db = a0 + cb < K ? e.d[e.cact[a0 + cb]] : 0.0;
    const long long jbase = (long long) chunk * kExpSymCols + 2 * q;
    constexpr int U = kExpSymIter / 128, NIT = kExpSymCols / kExpSymIter;
    struct Batch { int uj[U][2]; double2 v[U][2]; };
    auto load = [&](Batch &b, const int it) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long j = jbase + (long long) it * kExpSymIter + u * 128 + h;
                const bool ok = j < e.m;
                b.uj[u][h] = ok ? e.ind[j] : -2;
                const int c = ok ? e.cls[j] : 0;
                b.v[u][h] = __ldcg(reinterpret_cast<const double2 *>(e.G + (long long) c * e.ld + a0 + ca));
            }
    };
    auto store = [&](const Batch &b, const int it) {
        const long long j0 = jbase + (long long) it * kExpSymIter;
        const long long left = e.m - j0;                                 // even: a pair of columns is inside or outside together
        const int nu = left <= 0 ? 0 : (int) (left + 127 < 128LL * U ? (left + 127) / 128 : U);
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int cl = half ? cb : ca;
            const double dc = half ? db : da;
            const int nm = s_n[cl];
            for (int k = 0; k < nm; k++) {
                long long i; int ui;
                if (k < kExpSymMembers) { i = s_i[cl][k]; ui = s_u[cl][k]; } else { i = e.cmem[s_q0[cl] + k]; ui = e.ind[i]; }
                double2 *rowp = reinterpret_cast<double2 *>(e.out + (i - e.row0) * e.m + j0);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (u < nu) {
                        double2 val;
                        val.x = (b.uj[u][0] == ui) ? dc : (half ? b.v[u][0].y : b.v[u][0].x);
                        val.y = (b.uj[u][1] == ui) ? dc : (half ? b.v[u][1].y : b.v[u][1].x);
                        __stcs(rowp + u * 64, val);
                        sum += val.x + val.y;
                    }
            }
        }
    };
    // the segments of the next 256 columns are in flight while the current ones are written
    Batch b0, b1;
    load(b0, 0);
#pragma unroll 1
    for (int it = 0; it < NIT; it += 2) {
        load(b1, it + 1);
        store(b0, it);
        if (it + 2 < NIT) load(b0, it + 2);
        store(b1, it + 1);
    }
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
        e.partials[2 * ((long long) A * nchunks + chunk)] = x;
        e.partials[2 * ((long long) A * nchunks + chunk) + 1] = y;
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
    for (int k = g.seed_off[row] + threadIdx.x; k < g.seed_off[row + 1]; k += 256) {
        const int c = g.seed_col[k] - g.col0;
        if (c >= 0 && c < g.ncols) g.Vnew[row * g.ld + c] = 1.0;
    }
}

__global__ void __launch_bounds__(256) gc_gather_kernel(const double *V, long long ld, const int32_t *row_of,
                                                        double *out, long long m, long long a) {
    const long long i = blockIdx.x;
    const int r = row_of[i];
    for (long long c = threadIdx.x; c < a; c += 256) out[i * a + c] = r >= 0 ? V[(long long) r * ld + c] : 0.0;
}

// ---- device-side consumers of the result: the m x m matrix never has to cross PCIe -------------------
// One warp per result row, two passes (count, then fill at the row's offset): entries come out in the reference's order
// (row by row, columns ascending).  mode 0 = gen.phiOver (geneakit/compute.py:170-178): j < i and value > threshold;
// mode 1 = the CSR of gen.phi(sparse=True) (lib/compute.cpp:402-413,458-476): j <= i and |2 value| > 1e-9.
__global__ void __launch_bounds__(256) select_kernel(const SelectDev q, const int fill) {
    const int lane = threadIdx.x & 31;
    const long long row = q.row0 + (long long) blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= q.row1) return;
    const double *src = q.M + (row - q.row0) * q.m;
    const long long jend = q.mode == 0 ? row : row + 1;
    long long n = 0;
    const long long base = fill ? q.row_off[row - q.row0] : 0;
    for (long long j0 = 0; j0 < jend; j0 += 32) {
        const long long j = j0 + lane;
        double v = 0.0;
        bool ok = false;
        if (j < jend) { v = src[j]; ok = q.mode == 0 ? v > q.thr : fabs(2.0 * v) > q.thr; }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (fill && ok) {
            const long long at = base + n + __popc(m & ((1u << lane) - 1));
            q.oj[at] = (int32_t) j;
            if (q.mode == 0) { q.oi[at] = (int32_t) row; q.ov[at] = v; }
            else q.ovf[at] = (float) (2.0 * v) / 2.0f;
        }
        n += __popc(m);
    }
    if (!fill && lane == 0) q.row_count[row - q.row0] = n;
}

// gen.phiCI's statistic (geneakit/compute.py:226-229) for one bootstrap resample per CTA: the mean of the entries < 0.5 of
// the sub-matrix phi[idx][:, idx].  Fixed assignment of entries to threads and a fixed reduction tree: deterministic.
__global__ void __launch_bounds__(256) ci_kernel(const double *M, long long ld, const int32_t *idx, int n, double *stat) {
    __shared__ double ssum[256];
    __shared__ unsigned long long scnt[256];
    const int32_t *ix = idx + (long long) blockIdx.x * n;
    double s = 0.0;
    unsigned long long c = 0;
    for (int a = threadIdx.x >> 5; a < n; a += 8) {
        const double *row = M + (long long) ix[a] * ld;
        for (int b = threadIdx.x & 31; b < n; b += 32) {
            const double v = row[ix[b]];
            if (v < 0.5) { s += v; c++; }
        }
    }
    ssum[threadIdx.x] = s; scnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) stat[blockIdx.x] = scnt[0] ? ssum[0] / (double) scnt[0] : nan("");
}

}  // namespace

static int g_step_grid = 0;
int step_grid() { return g_step_grid; }

cudaError_t configure_kernels() {
    cudaError_t e;
    int dev = 0, sms = 0, occ = 0;
    if ((e = cudaGetDevice(&dev))) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kStepSmem))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phi_step_kernel, kThreads, kStepSmem))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kStepSmemTma))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_g4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kStepSmemG4))) return e;
    // the three step kernels are launched with the same persistent grid, and their waits rely on every CTA being resident:
    // the grid is sized by the kernel that fits the fewest CTAs per SM
    {
        int occ_tma = 0, occ_g4 = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tma, phi_step_tma_kernel, kThreadsTma, kStepSmemTma))) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g4, phi_step_g4_kernel, kThreadsG4, kStepSmemG4))) return e;
        occ = occ < occ_tma ? occ : occ_tma;
        occ = occ < occ_g4 ? occ : occ_g4;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
    }
    const char *env = getenv("GK_CTAS_PER_SM");
    if (env && atoi(env) > 0 && atoi(env) < occ) occ = atoi(env);
    g_step_grid = sms * (occ > 0 ? occ : 1);     // persistent: one wave, every CTA resident (the waits rely on it)
    return cudaSuccess;
}

cudaError_t launch_step(const StepDev &s, cudaStream_t st) {
    if (s.total_items <= 0) return cudaSuccess;
    long long grid = g_step_grid > 0 ? g_step_grid : 148;
    const long long need = (s.total_items + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    phi_step_kernel<<<(unsigned) grid, kThreads, kStepSmem, st>>>(s);
    return cudaGetLastError();
}

cudaError_t launch_step_tma(const StepDev &s, cudaStream_t st) {
    if (s.total1 + s.total2 <= 0) return cudaSuccess;
    const long long grid = g_step_grid > 0 ? g_step_grid : 148;
    phi_step_tma_kernel<<<(unsigned) grid, kThreadsTma, kStepSmemTma, st>>>(s);
    return cudaGetLastError();
}

cudaError_t launch_step_g4(const StepDev &s, cudaStream_t st) {
    if (s.total_g4 + s.total2 <= 0) return cudaSuccess;
    const long long grid = g_step_grid > 0 ? g_step_grid : 148;
    phi_step_g4_kernel<<<(unsigned) grid, kThreadsG4, kStepSmemG4, st>>>(s);
    return cudaGetLastError();
}

cudaError_t launch_diag(const DiagDev &d, cudaStream_t st) {
    long long work = (long long) (d.c_end - d.c_begin);
    if (d.ngroups > work) work = d.ngroups;
    if (work <= 0) return cudaSuccess;
    long long blocks = (work + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    phi_diag_kernel<<<(unsigned) blocks, 256, 0, st>>>(d);
    return cudaGetLastError();
}

cudaError_t launch_init(double *G0, double *d0, const int32_t *flags, int64_t K0, int64_t Kpad0, cudaStream_t st) {
    long long total = (long long) Kpad0 * Kpad0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    init_kernel<<<(unsigned) blocks, 256, 0, st>>>(G0, d0, flags, K0, Kpad0);
    return cudaGetLastError();
}

cudaError_t launch_init_rows(double *G0, double *d0, const int32_t *flags, int64_t K0, int64_t Kpad0,
                             int64_t row_begin, int64_t rows, int64_t dlen, cudaStream_t st) {
    init_rows_kernel<<<148 * 2, 256, 0, st>>>(G0, d0, flags, K0, Kpad0, row_begin, rows, dlen);
    return cudaGetLastError();
}

cudaError_t launch_fetch_rows(const RowTable &t, int64_t ld, const int32_t *rows, int32_t first, int32_t count, double *dst, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    fetch_rows_kernel<<<(unsigned) count, 256, 0, st>>>(t, ld, rows, dst, first);
    return cudaGetLastError();
}

cudaError_t launch_pull_mirror(const RowTable &t, int64_t ld, int32_t pb, int32_t pe, int32_t np, double *own, int64_t own_row0, cudaStream_t st) {
    if (pe <= pb || np <= pe) return cudaSuccess;
    pull_mirror_kernel<<<148 * 8, 128, 0, st>>>(t, ld, pb, pe, np, own, own_row0);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandDev &e, int32_t a_begin, int32_t a_count, cudaStream_t st) {
    if (a_count <= 0) return cudaSuccess;
    ExpandDev x = e;
    x.a_begin = a_begin;
    expand_kernel<<<(unsigned) ((long long) a_count * e.nchunks), 256, 0, st>>>(x);
    return cudaGetLastError();
}

int64_t expand_sym_blocks(int32_t nact, int64_t m) { return (int64_t) ((nact + 7) / 8) * ((m + kExpSymCols - 1) / kExpSymCols); }

cudaError_t launch_expand_sym(const ExpandDev &e, int32_t a_begin, int32_t a_count, cudaStream_t st) {
    if (a_count <= 0) return cudaSuccess;
    ExpandDev x = e;
    x.a_begin = a_begin;                                   // a multiple of 8
    expand_sym_kernel<<<(unsigned) expand_sym_blocks(a_count, e.m), 256, 0, st>>>(x);
    return cudaGetLastError();
}

cudaError_t launch_fetch_rows_t(const RowTable &t, int64_t ld, const int32_t *rows, int32_t a_begin, int32_t a_count, int32_t nact,
                                int32_t column_panels, double *GT, int64_t ldt, cudaStream_t st) {
    if (a_count <= 0) return cudaSuccess;
    fetch_rows_t_kernel<<<148 * 8, 128, 0, st>>>(t, ld, rows, a_begin, a_count, nact, column_panels, GT, ldt);
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

cudaError_t launch_select(const SelectDev &q, int fill, cudaStream_t st) {
    const long long rows = q.row1 - q.row0;
    if (rows <= 0) return cudaSuccess;
    select_kernel<<<(unsigned) ((rows + 7) / 8), 256, 0, st>>>(q, fill);
    return cudaGetLastError();
}

cudaError_t launch_ci(const double *M, int64_t ld, const int32_t *idx, int32_t b, int32_t n, double *stat, cudaStream_t st) {
    if (b <= 0) return cudaSuccess;
    ci_kernel<<<(unsigned) b, 256, 0, st>>>(M, ld, idx, n, stat);
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
