// Hand-written sm_100a kernels of the dense-kinship path (no Triton, no library kernels).
//
// The step kernel replaces the reference's hot loop nest
//   compute_kinship_between_probands (lib/compute.cpp:571-588)  [n'(n'-1)/2 calls of]
//   compute_kinship                  (lib/compute.cpp:491-550)
//   compute_kinship_with_oneself     (lib/compute.cpp:553-568)
// for one pair of adjacent vertex cuts.  It is an HBM-bound FP64 gather-average: per output
// cell 4 gathered reads, 3 DADD + 1 DMUL, one 8-byte store.  Tensor cores are not used.
//
// One CTA = one unordered pair (I >= J) of T-class tiles of the NEW matrix:
//   1. stage  E[rows(I)][cols(J)]  of the PREVIOUS matrix in shared memory with cp.async
//      (LDGSTS): rows(I)/cols(J) = the tile's contiguous parent WINDOW (16-byte coalesced row
//      segments) + its LIST of scattered parent rows.  The window-rows x list-cols block is
//      fetched through the symmetric entry (list row, window column) so it is coalesced too;
//      only list x list is an 8-byte gather.
//   2. patch "same individual" entries with the self kinship (rule :522-528),
//   3. out[a][b] = 1/4 * ((p + q) + (r + s)) from shared memory (RANKED: grouping chosen per
//      pair by rank, reproducing the reference's rounding on pedigrees deeper than 26),
//   4. store tile (I,J) and, transposed through shared memory, its mirror (J,I): both coalesced.
// Tile pairs are enumerated I-major so CTAs resident together share rows(I): the scattered
// list rows are then served from the 126 MB L2 instead of HBM.
#include "gk_kernels.cuh"

namespace gk {

namespace {

constexpr int kThreads = 256;
constexpr int kMiss = 0xFFFF;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

constexpr int kConf = 12;      // == gk::kConfMax (gk_plan.hpp)

// E(x, y) of the previous cut for one (row parent, column parent) pair, applying
// "same individual -> self kinship" (lib/compute.cpp:522-528) when the CTA is on the slow path.
template <bool CHECK>
__device__ __forceinline__ double fetch_e(const double *S, int pitch, int ra, int cb, int ia, int ib,
                                          const double *dprev, int cls_a) {
    if (ra == kMiss || cb == kMiss) return 0.0;
    if (CHECK && ia == ib) return dprev[cls_a];
    return S[(size_t) ra * pitch + cb];
}

template <int T, bool RANKED>
__global__ void __launch_bounds__(kThreads) phi_step_kernel(const StepDev s) {
    constexpr int W = T, LMAX = 2 * T, DW = 4 + LMAX + kConf;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    // ---- which tile pair (I-major: CTAs resident together share rows(I) -> L2 reuse) ----------
    const long long bid = blockIdx.x;
    int I = (int) ((sqrt(8.0 * (double) bid + 1.0) - 1.0) * 0.5);
    while ((long long) I * (I + 1) / 2 > bid) --I;
    while ((long long) (I + 1) * (I + 2) / 2 <= bid) ++I;
    const int J = (int) (bid - (long long) I * (I + 1) / 2);
    const int tid = threadIdx.x;
    const int pitch = s.pitch;

    // ---- shared memory carve-up ----------------------------------------------------------------
    double *S = reinterpret_cast<double *>(smem_raw);                    // nr_max x pitch; reused as O (T x (T+1))
    int32_t *dI = reinterpret_cast<int32_t *>(S + (size_t) pitch * pitch); // DW: descriptor of tile I
    int32_t *dJ = dI + DW;                                                // DW
    int32_t *laI = dJ + DW;                                               // T
    int32_t *laJ = laI + T;                                               // T
    int32_t *piI = laJ + T;                                               // 2T parent individuals
    int32_t *piJ = piI + 2 * T;                                           // 2T
    int32_t *rkI = piJ + 2 * T;                                           // T (RANKED)
    int32_t *rkJ = rkI + T;                                               // T (RANKED)

    // ---- phase 0: ONE round trip for every index the CTA needs (fixed-size descriptors) --------
    for (int i = tid; i < 2 * DW; i += kThreads) {
        const int t = i < DW ? I : J, w = i < DW ? i : i - DW;
        dI[i] = s.desc[(size_t) t * DW + w];
    }
    for (int i = tid; i < 2 * T; i += kThreads) {
        const int t = i < T ? I : J, k = i < T ? i : i - T;
        const long long c = (long long) t * T + k;
        const bool ok = c < s.K;
        laI[i] = ok ? s.la[c] : (kMiss | (kMiss << 16));
        piI[2 * i] = ok ? s.pind[2 * c] : -1;
        piI[2 * i + 1] = ok ? s.pind[2 * c + 1] : -1;
        if (RANKED) rkI[i] = ok ? s.cls_rank[c] : 0;
    }
    __syncthreads();
    const int c0I = dI[0], nlI = dI[1], c0J = dJ[0], nlJ = dJ[1];
    const int32_t *listI = dI + 4, *listJ = dJ + 4;
    const int nrI = W + nlI;
    // slow path <=> the two tiles share a parent individual (or it is a diagonal tile)
    bool slow = (I == J) || dI[2] < 0 || dJ[2] < 0;
    if (!slow)
        for (int k = 0; k < dI[2]; k++) slow |= (dI[4 + LMAX + k] == J);

    // ---- phase 1: stage E[rows(I)][cols(J)] with cp.async (second and last global round trip) --
    const double *__restrict__ Gp = s.Gprev;
    const long long ldp = s.ld_prev;
    const int lastrow = (int) (s.K_prev - 1);
    {   // (a) all rows x window columns of J: 16-byte chunks, half a warp (T=32) per 256-byte segment
        constexpr int CH = W / 2;
        for (int r = tid / CH; r < nrI; r += kThreads / CH) {
            const int ch = tid % CH;
            int row = r < W ? c0I + r : listI[r - W];
            row = row > lastrow ? lastrow : row;
            cp_async16(S + (size_t) r * pitch + 2 * ch, Gp + (long long) row * ldp + c0J + 2 * ch);
        }
    }
    {   // (b) window rows of I x list columns of J through the symmetric entry (list row, window col)
        for (int j = tid / W; j < nlJ; j += kThreads / W) {
            const int r = tid % W;
            cp_async8(S + (size_t) r * pitch + W + j, Gp + (long long) listJ[j] * ldp + c0I + r);
        }
    }
    {   // (c) list rows of I x list columns of J: 8-byte gathers (L2-resident across the I sweep)
        const int lane = tid & 31, warp = tid >> 5;
        for (int j0 = 0; j0 < nlJ; j0 += 32) {
            const int j = j0 + lane;
            if (j < nlJ) {
                const long long col = listJ[j];
                for (int r = warp; r < nlI; r += kThreads / 32)
                    cp_async8(S + (size_t) (W + r) * pitch + W + j, Gp + (long long) listI[r] * ldp + col);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 2: gather-average from shared memory -----------------------------------------
    constexpr int CPT = T * T / kThreads;      // cells per thread
    double v[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int cell = tid + k * kThreads;
        const int a = cell / T, b = cell % T;
        const int la = laI[a], lb = laJ[b];
        const int a0 = la & 0xFFFF, a1 = (la >> 16) & 0xFFFF;
        const int b0 = lb & 0xFFFF, b1 = (lb >> 16) & 0xFFFF;
        double p, q, r, t;
        if (!slow) {
            p = fetch_e<false>(S, pitch, a0, b0, 0, 0, nullptr, 0);
            q = fetch_e<false>(S, pitch, a1, b0, 0, 0, nullptr, 0);
            r = fetch_e<false>(S, pitch, a0, b1, 0, 0, nullptr, 0);
            t = fetch_e<false>(S, pitch, a1, b1, 0, 0, nullptr, 0);
        } else {
            const int ia0 = piI[2 * a], ia1 = piI[2 * a + 1], ib0 = piJ[2 * b], ib1 = piJ[2 * b + 1];
            const int ca0 = a0 == kMiss ? 0 : (a0 < W ? c0I + a0 : listI[a0 - W]);
            const int ca1 = a1 == kMiss ? 0 : (a1 < W ? c0I + a1 : listI[a1 - W]);
            p = fetch_e<true>(S, pitch, a0, b0, ia0, ib0, s.dprev, ca0);
            q = fetch_e<true>(S, pitch, a1, b0, ia1, ib0, s.dprev, ca1);
            r = fetch_e<true>(S, pitch, a0, b1, ia0, ib1, s.dprev, ca0);
            t = fetch_e<true>(S, pitch, a1, b1, ia1, ib1, s.dprev, ca1);
        }
        if (RANKED) {
            // lib/compute.cpp:529-548: the higher-ranked individual is expanded first, so the
            // inner sums run over the LOWER-ranked one's parents.
            v[k] = (rkI[a] < rkJ[b]) ? 0.25 * (__dadd_rn(p, q) + __dadd_rn(r, t))
                                     : 0.25 * (__dadd_rn(p, r) + __dadd_rn(q, t));
        } else {
            v[k] = 0.25 * ((p + q) + (r + t));
        }
    }
    if (I == J) {
        // self kinship of the new classes (lib/compute.cpp:553-568): a kept individual copies its
        // value (bitwise what the reference recomputes); a new one gets 1/2 + 1/2 phi(father, mother)
        for (int a = tid; a < T; a += kThreads) {
            const long long ga = (long long) I * T + a;
            if (ga < s.K) {
                const int la = laI[a];
                const int a0 = la & 0xFFFF, a1 = (la >> 16) & 0xFFFF;
                const int i0 = piI[2 * a], i1 = piI[2 * a + 1];
                double d = 0.5;
                if (a0 != kMiss && i0 == i1) d = s.dprev[a0 < W ? c0I + a0 : listI[a0 - W]];
                else if (a0 != kMiss && a1 != kMiss) d = 0.5 + 0.5 * S[(size_t) a0 * pitch + a1];
                s.dnew[ga] = d;
            }
        }
    }
    // ---- phase 3: store tile (I,J) from registers; mirror (J,I) transposed through smem --------
    const long long ldn = s.ld_new;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
        const int cell = tid + k * kThreads;
        const int a = cell / T, b = cell % T;
        const long long ga = (long long) I * T + a, gb = (long long) J * T + b;
        if (ga < s.K && gb < s.K) s.Gnew[ga * ldn + gb] = v[k];
    }
    if (I != J) {
        __syncthreads();                       // everyone is done reading S
        double *O = S;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int cell = tid + k * kThreads;
            O[(cell / T) * (T + 1) + cell % T] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int cell = tid + k * kThreads;
            const int b = cell / T, a = cell % T;          // lanes run along a -> coalesced mirror rows
            const long long ga = (long long) I * T + a, gb = (long long) J * T + b;
            if (ga < s.K && gb < s.K) s.Gnew[gb * ldn + ga] = O[a * (T + 1) + b];
        }
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
// the gen.phiMean reduction (geneakit/compute.py:124-132): every CTA writes a 16 x 256 block
// of the m x m result and one (sum, diagonal sum) partial.
constexpr int kExpRows = 16;
__global__ void __launch_bounds__(256) expand_kernel(const ExpandDev e) {
    const long long j = (long long) blockIdx.x * 256 + threadIdx.x;
    const long long i0 = (long long) blockIdx.y * kExpRows;
    double sum = 0.0, dsum = 0.0;
    if (j < e.m) {
        const int cj = e.single_cut ? 0 : e.cls[j];
        const int uj = e.ind[j];
#pragma unroll 4
        for (int k = 0; k < kExpRows; k++) {
            const long long i = i0 + k;
            if (i >= e.m) break;
            double v;
            if (e.single_cut) {
                v = (i == j) ? 0.5 : 0.0;
            } else {
                const int ci = e.cls[i];
                v = (e.ind[i] == uj) ? e.d[ci] : e.G[(long long) ci * e.ld + cj];
            }
            e.out[i * e.m + j] = v;
            sum += v;
            if (i == j) dsum += v;
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
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) { a += ws[w]; b += wd[w]; }
        const long long blk = (long long) blockIdx.y * gridDim.x + blockIdx.x;
        e.partials[2 * blk] = a;
        e.partials[2 * blk + 1] = b;
    }
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

template <int T, bool RANKED>
cudaError_t launch_step_t(const StepDev &s, size_t smem, cudaStream_t st) {
    const long long pairs = (long long) s.nt * (s.nt + 1) / 2;
    phi_step_kernel<T, RANKED><<<(unsigned) pairs, kThreads, smem, st>>>(s);
    return cudaGetLastError();
}

}  // namespace

size_t step_smem_bytes(int tile, int nr_max, bool ranked, int *pitch_out) {
    (void) ranked;
    const int W = tile, LMAX = 2 * tile, DW = 4 + LMAX + kConf;
    int pitch = (nr_max + 1) & ~1;
    if (pitch < W + 2) pitch = W + 2;          // also large enough for the T x (T+1) transpose buffer
    while (pitch % 4 != 2) pitch += 2;         // even (16-byte rows) and 2 mod 4 (transposed stores: 2-way at worst)
    if (pitch_out) *pitch_out = pitch;
    size_t bytes = (size_t) pitch * pitch * sizeof(double);
    bytes += (size_t) (2 * DW + 2 * tile + 4 * tile + 2 * tile) * sizeof(int32_t);
    return bytes;
}

cudaError_t configure_kernels() {
    const int maxsmem = 227 * 1024;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(phi_step_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem))) return e;
    if ((e = cudaFuncSetAttribute(phi_step_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsmem))) return e;
    return cudaSuccess;
}

cudaError_t launch_step(const StepDev &s, int tile, bool ranked, size_t smem, cudaStream_t st) {
    if (s.nt <= 0) return cudaSuccess;
    if (tile == 64) return ranked ? launch_step_t<64, true>(s, smem, st) : launch_step_t<64, false>(s, smem, st);
    return ranked ? launch_step_t<32, true>(s, smem, st) : launch_step_t<32, false>(s, smem, st);
}

cudaError_t launch_init(double *G0, double *d0, int64_t K0, int64_t ld0, cudaStream_t st) {
    long long total = (long long) K0 * ld0;
    int blocks = (int) ((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    init_kernel<<<blocks, 256, 0, st>>>(G0, d0, K0, ld0);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandDev &e, int *nblocks_out, cudaStream_t st) {
    dim3 grid((unsigned) ((e.m + 255) / 256), (unsigned) ((e.m + kExpRows - 1) / kExpRows));
    if (nblocks_out) *nblocks_out = (int) (grid.x * grid.y);
    expand_kernel<<<grid, 256, 0, st>>>(e);
    return cudaGetLastError();
}

cudaError_t launch_mean_finish(const double *partials, int nblocks, double *sums, cudaStream_t st) {
    mean_finish_kernel<<<1, 256, 0, st>>>(partials, nblocks, sums);
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
