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
#include "gk_kernels.cuh"
#include "gk_plan.hpp"

namespace gk {

namespace {

constexpr int kThreads = 256;
constexpr int T = kPanel;                 // 32
constexpr int PS1 = kCB + 1;              // phase-1 transpose pitch (doubles): conflict-free both ways
constexpr int PS2 = T + 1;                // phase-2 transpose pitch

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_count(const int *p, int target) {
    while (ld_acquire(p) < target) __nanosleep(32);
}
__device__ __forceinline__ void signal_count(int *p) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(p) : "memory");
}

__device__ __forceinline__ double *row_ptr(const RowTable &t, int row, long long ld) {
    double *base = t.base[0];
    int r0 = 0;
#pragma unroll
    for (int q = 1; q < kMaxShards; q++)
        if (q < t.n && row >= t.row0[q]) { base = t.base[q]; r0 = t.row0[q]; }
    return base + (long long) (row - r0) * ld;
}

__global__ void __launch_bounds__(kThreads, 4) phi_step_kernel(const StepDev s) {
    __shared__ double S[T * PS1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int g = 0;
    int ready_panel = -1;        // phase-1 of this panel is known complete
    int freed_panel = -1;        // phase-2 of this panel is known complete (its ring slot is free)
    for (long long q = blockIdx.x; q < s.total_items; q += gridDim.x) {
        while (q >= s.grp_off[g + 1]) ++g;
        const int info = s.grp_info[g];
        const int I = info & 0x3FFFFFFF;
        const int local = (int) (q - s.grp_off[g]);
        double *buf = s.ring + (long long) ((I - s.panel_begin) % kRing) * s.ring_stride;
        if ((info >> 30) == 0) {
            // ---------------- phase 1: N^T[c][i] for c in [c0, c0 + 64) -------------------------
            const int c0 = local * kCB;
            double a[4][2], b[4][2];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int i = I * T + warp + 8 * t;
                const int F = s.pF[i], M = s.pM[i];
                const double *rf = F >= 0 ? row_ptr(s.prev, F, s.ld_prev) : nullptr;
                const double *rm = M >= 0 ? row_ptr(s.prev, M, s.ld_prev) : nullptr;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int c = c0 + lane + 32 * h;
                    const bool ok = c < s.Kpad_prev;
                    a[t][h] = (rf && ok) ? __ldg(rf + c) : 0.0;
                    b[t][h] = (rm && ok) ? __ldg(rm + c) : 0.0;
                }
            }
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int h = 0; h < 2; h++) S[(warp + 8 * t) * PS1 + lane + 32 * h] = 0.5 * (a[t][h] + b[t][h]);
            // the ring slot must have been drained by phase 2 of the panel that used it before
            const int old = I - kRing;
            if (old >= s.panel_begin && old > freed_panel) {
                if (tid == 0) wait_count(s.done2 + old, s.full_rows ? s.np : old + 1);
                freed_panel = old;
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < kCB / 8; t++) {
                const int cl = warp + 8 * t, c = c0 + cl;
                if (c < s.Kpad_prev) __stcg(buf + (long long) c * T + lane, S[lane * PS1 + cl]);
            }
            __syncthreads();
            if (tid == 0) signal_count(s.done1 + I);
        } else {
            // ---------------- phase 2: tile (J, I) and its mirror --------------------------------
            const int J = local;
            if (I != ready_panel) {
                if (tid == 0) wait_count(s.done1 + I, s.ncb);
                ready_panel = I;
                __syncthreads();
            }
            double v[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int j = J * T + warp + 8 * t;
                const int F = s.pF[j], M = s.pM[j];
                const double x = F >= 0 ? __ldcg(buf + (long long) F * T + lane) : 0.0;
                const double y = M >= 0 ? __ldcg(buf + (long long) M * T + lane) : 0.0;
                v[t] = 0.5 * (x + y);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) S[(warp + 8 * t) * PS2 + lane] = v[t];        // S[j][i]
            if (!s.full_rows && J != I) {
                // direct: row j of tile J, columns of panel I (the mirror of what this panel owns)
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int j = J * T + warp + 8 * t;
                    row_ptr(s.out, j, s.ld_new)[I * T + lane] = v[t];
                }
            }
            __syncthreads();
            if (J != I || s.full_rows) {
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int il = warp + 8 * t;
                    row_ptr(s.out, I * T + il, s.ld_new)[J * T + lane] = S[lane * PS2 + il];    // B[i][j] = S[j][i]
                }
            } else {
                // diagonal tile: element (j, i) is valid where j <= i (in RANKED mode that is the
                // reference's grouping); the other triangle is its mirror
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int r = warp + 8 * t;
                    row_ptr(s.out, I * T + r, s.ld_new)[I * T + lane] = r <= lane ? S[r * PS2 + lane] : S[lane * PS2 + r];
                }
            }
            __syncthreads();
            if (tid == 0) signal_count(s.done2 + I);
        }
    }
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

// Multi-GPU expansion: copy the class rows this rank's probands need from their owners (peer
// reads over NVLink, 16-byte vectorised, one CTA per row) into a compact local matrix.
__global__ void __launch_bounds__(256) fetch_rows_kernel(const RowTable t, long long ld, const int32_t *rows, double *dst) {
    const double2 *src = reinterpret_cast<const double2 *>(row_ptr(t, rows[blockIdx.x], ld));
    double2 *out = reinterpret_cast<double2 *>(dst + (long long) blockIdx.x * ld);
    for (long long c = threadIdx.x; c < ld / 2; c += 256) out[c] = src[c];
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
    const double *row = e.G + (long long) (e.compact ? a : c) * e.ld;
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

static int g_step_grid = 0;
int step_grid() { return g_step_grid; }

cudaError_t configure_kernels() {
    cudaError_t e;
    int dev = 0, sms = 0, occ = 0;
    if ((e = cudaGetDevice(&dev))) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev))) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phi_step_kernel, kThreads, 0))) return e;
    const char *env = getenv("GK_CTAS_PER_SM");
    if (env && atoi(env) > 0 && atoi(env) < occ) occ = atoi(env);
    g_step_grid = sms * (occ > 0 ? occ : 1);     // persistent: one wave, every CTA resident (the waits rely on it)
    return cudaSuccess;
}

cudaError_t launch_step(const StepDev &s, cudaStream_t st) {
    if (s.total_items <= 0) return cudaSuccess;
    long long grid = g_step_grid > 0 ? g_step_grid : 148;
    if (grid > s.total_items) grid = s.total_items;
    phi_step_kernel<<<(unsigned) grid, kThreads, 0, st>>>(s);
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

cudaError_t launch_fetch_rows(const RowTable &t, int64_t ld, const int32_t *rows, int32_t nrows, double *dst, cudaStream_t st) {
    if (nrows <= 0) return cudaSuccess;
    fetch_rows_kernel<<<(unsigned) nrows, 256, 0, st>>>(t, ld, rows, dst);
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
