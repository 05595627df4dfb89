// Microbenchmarks behind the round-2 step kernel (B200, sm_100a):
//   mode "check B"   : does cp.async.bulk.tensor.2d ... tile::gather4 deliver rows {r0..r3} x W columns when the
//                      tensor map's box is {W, B} (B = 1 or 4)?  Prints the first mismatches.
//   mode "bench"     : throughput of (a) gather4 of W-column rows, (b) 1-D bulk copies, from an L2-resident
//                      "ring" (10 MB) and from a matrix far larger than L2; (c) LDG.128 streaming of an
//                      L2-resident / DRAM-resident buffer (the L2 cap); (d) plain streaming stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_bench gather4_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeTiled) fn;
}
// 2-D f64 tensor [rows][cols] (cols contiguous, pitch ld doubles), box {W, B}
static CUtensorMap make_map(double *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t W, uint32_t B) {
    static EncodeTiled enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[2] = { cols, rows }, strides[1] = { ld * 8 };
    cuuint32_t box[2] = { W, B }, es[2] = { 1, 1 };
    // no FLOAT64 enum before 12.x? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 exists since 12.0
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (W=%u B=%u)\n", (int) r, W, B); exit(3); }
    return m;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned par) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void gather4(unsigned dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- correctness: one gather4 of rows r[0..3] at column c0, W columns; dump shared memory ----
__global__ void check_kernel(const __grid_constant__ CUtensorMap map, int c0, int r0, int r1, int r2, int r3, int W, double *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + 4 * W * 8);
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = -1.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(smem_u32(bar), 4 * W * 8);
        gather4(smem_u32(sm), &map, c0, r0, r1, r2, r3, smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    for (int i = threadIdx.x; i < 4 * W; i += blockDim.x) out[i] = reinterpret_cast<double *>(sm)[i];
}

// ---- throughput: one warp per CTA drives the TMA engine; stage = R rows x W columns; NS stages in flight ----
template <int MODE>   // 0: gather4 (R/4 instructions per stage)   1: 1-D bulk copies (R per stage)
__global__ void __launch_bounds__(32, 1) tma_kernel(const __grid_constant__ CUtensorMap map, const double *A, long long ld, int W, int R, int NS,
                                                    const int *rows, int nrows, int ncolblocks, long long items, double *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int stage_bytes = R * W * 8;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + (size_t) NS * stage_bytes);
    const int lane = threadIdx.x;
    if (lane == 0) { for (int i = 0; i < NS; i++) mbar_init(smem_u32(bars + i), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncwarp();
    const long long per = items / gridDim.x;
    double acc = 0;
    for (long long k = 0; k < per + NS; k++) {
        if (k >= NS) {
            const int st = (int) ((k - NS) % NS);
            mbar_wait(smem_u32(bars + st), (unsigned) (((k - NS) / NS) & 1));
            acc += reinterpret_cast<double *>(sm + (size_t) st * stage_bytes)[lane];
            __syncwarp();
        }
        if (k < per) {
            const int st = (int) (k % NS);
            const long long it = blockIdx.x * per + k;
            const int cb = (int) (it % ncolblocks);
            if (lane == 0) mbar_expect(smem_u32(bars + st), stage_bytes);
            __syncwarp();
            if (MODE == 0) {
                for (int g = lane; g < R / 4; g += 32) {
                    const int b = (int) ((it * 7 + g * 52) % (nrows - 4));
                    gather4(smem_u32(sm + (size_t) st * stage_bytes + (size_t) g * 4 * W * 8), &map, cb * W,
                            rows[b], rows[b + 1], rows[b + 2], rows[b + 3], smem_u32(bars + st));
                }
            } else {
                for (int r = lane; r < R; r += 32) {
                    const int row = rows[(it * 7 + r * 13) % nrows];
                    bulk(smem_u32(sm + (size_t) st * stage_bytes + (size_t) r * W * 8), A + (long long) row * ld + (long long) cb * W, W * 8, smem_u32(bars + st));
                }
            }
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}

// ---- LDG.128 streaming of a buffer (n16 16-byte words), repeated `reps` times: L2-resident or DRAM ----
__global__ void __launch_bounds__(1024, 1) stream_read(const double2 *p, long long n16, int reps, double *sink) {
    double a = 0;
    for (int r = 0; r < reps; r++)
        for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n16; i += (long long) gridDim.x * blockDim.x) {
            double2 v;
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + i));
            a += v.x + v.y;
        }
    if (a == 12345.678) sink[0] = a;
}
__global__ void __launch_bounds__(1024, 1) stream_write(double2 *p, long long n16, int reps) {
    for (int r = 0; r < reps; r++)
        for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n16; i += (long long) gridDim.x * blockDim.x)
            p[i] = make_double2(1.0, 2.0);
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); return ms; }

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "bench";
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double *sink; CK(cudaMalloc(&sink, 8));
    if (!strcmp(mode, "check")) {
        const int B = argc > 2 ? atoi(argv[2]) : 1, W = argc > 3 ? atoi(argv[3]) : 32;
        const long long rows = 4096, ld = 1024;
        std::vector<double> h((size_t) rows * ld);
        for (long long r = 0; r < rows; r++) for (long long c = 0; c < ld; c++) h[r * ld + c] = (double) (r * 65536 + c);
        double *A, *out; CK(cudaMalloc(&A, h.size() * 8)); CK(cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&out, 4 * W * 8));
        CUtensorMap m = make_map(A, rows, ld, ld, W, B);
        const int c0 = 64, r[4] = { 1000, 7, 3333, 42 };
        CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * W * 8 + 64));
        check_kernel<<<1, 128, 4 * W * 8 + 64>>>(m, c0, r[0], r[1], r[2], r[3], W, out);
        CK(cudaDeviceSynchronize());
        std::vector<double> o(4 * W); CK(cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int k = 0; k < 4; k++) for (int j = 0; j < W; j++) {
            const double want = (double) ((long long) r[k] * 65536 + c0 + j);
            if (o[k * W + j] != want && bad++ < 8) printf("  mismatch k=%d j=%d got %.0f (row %lld col %lld) want %.0f\n", k, j, o[k * W + j], (long long) o[k * W + j] / 65536, (long long) o[k * W + j] % 65536, want);
        }
        printf("gather4 check box={%d,%d}: %s (%d mismatches of %d)\n", W, B, bad ? "WRONG" : "OK", bad, 4 * W);
        return bad ? 1 : 0;
    }
    const int B = argc > 2 ? atoi(argv[2]) : 1;
    // big matrix (DRAM) and small ring (L2)
    const long long ld = 40000, nr = 24000;
    double *A; CK(cudaMalloc(&A, ld * nr * 8)); CK(cudaMemset(A, 0, ld * nr * 8));
    const long long ring_rows = 40000;
    double *ring; CK(cudaMalloc(&ring, ring_rows * 32 * 8)); CK(cudaMemset(ring, 0, ring_rows * 32 * 8));
    std::vector<int> h(65536); srand(1);
    int *rowsA, *rowsR;
    for (auto &x : h) x = rand() % nr;
    CK(cudaMalloc(&rowsA, h.size() * 4)); CK(cudaMemcpy(rowsA, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    for (auto &x : h) x = rand() % ring_rows;
    CK(cudaMalloc(&rowsR, h.size() * 4)); CK(cudaMemcpy(rowsR, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const int nrows = (int) h.size();

    auto run = [&](int MODE, bool l2, int W, int R, int NS) {
        double *base = l2 ? ring : A;
        const long long tld = l2 ? 32 : ld, trows = l2 ? ring_rows : nr;
        if (l2 && W != 32) return;
        CUtensorMap m = make_map(base, trows, tld, tld, W, B);
        const size_t smem = (size_t) NS * R * W * 8 + NS * 8;
        const long long items = 148LL * (l2 ? 20000 : 3000) * 40 / R * 32 / W;
        const int ncb = (int) (tld / W);
        auto k = MODE == 0 ? tma_kernel<0> : tma_kernel<1>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k<<<148, 32, smem>>>(m, base, tld, W, R, NS, l2 ? rowsR : rowsA, nrows, ncb, items / 10, sink);
        CK(cudaEventRecord(e0));
        k<<<148, 32, smem>>>(m, base, tld, W, R, NS, l2 ? rowsR : rowsA, nrows, ncb, items, sink);
        CK(cudaEventRecord(e1));
        const float ms = time_ms(e0, e1);
        CK(cudaGetLastError());
        const double ops = (double) items / 148 * (MODE == 0 ? R / 4 : R);
        printf("%-8s %-4s W=%3d (%4d B rows) R=%2d NS=%d (%3zu KB smem): %7.3f ms  %6.0f GB/s  %5.1f cycles/instr/SM\n", MODE == 0 ? "gather4" : "bulk1d",
               l2 ? "L2" : "DRAM", W, W * 8, R, NS, smem / 1024, ms, (double) items * R * W * 8 / ms / 1e6, ms * 1e-3 * 1.9e9 / ops);
        fflush(stdout);
    };
    for (int l2 = 1; l2 >= 0; l2--) {
        run(0, l2, 32, 40, 8); run(0, l2, 32, 40, 16); run(1, l2, 32, 40, 8); run(1, l2, 32, 40, 16);
    }
    run(0, false, 64, 40, 8); run(0, false, 128, 40, 4); run(0, false, 256, 40, 2); run(0, false, 256, 20, 5);
    run(1, false, 64, 40, 8); run(1, false, 128, 40, 4); run(1, false, 256, 40, 2); run(1, false, 256, 20, 5);

    // L2 cap: LDG.128 over an L2-resident buffer (32 MB) vs a 4 GB buffer; streaming stores
    for (long long mb : { 32LL, 64LL, 4096LL }) {
        const long long n16 = mb * 1024 * 1024 / 16;
        const int reps = mb <= 64 ? 200 : 4;
        stream_read<<<148, 1024>>>(reinterpret_cast<double2 *>(A), n16, 2, sink);
        CK(cudaEventRecord(e0));
        stream_read<<<148, 1024>>>(reinterpret_cast<double2 *>(A), n16, reps, sink);
        CK(cudaEventRecord(e1));
        float ms = time_ms(e0, e1);
        printf("LDG.128.cg stream read  %5lld MB x %3d: %7.3f ms  %6.0f GB/s\n", mb, reps, ms, (double) n16 * 16 * reps / ms / 1e6);
    }
    for (long long mb : { 32LL, 4096LL }) {
        const long long n16 = mb * 1024 * 1024 / 16;
        const int reps = mb <= 64 ? 200 : 4;
        CK(cudaEventRecord(e0));
        stream_write<<<148, 1024>>>(reinterpret_cast<double2 *>(A), n16, reps);
        CK(cudaEventRecord(e1));
        float ms = time_ms(e0, e1);
        printf("STG.128 stream write    %5lld MB x %3d: %7.3f ms  %6.0f GB/s\n", mb, reps, ms, (double) n16 * 16 * reps / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
