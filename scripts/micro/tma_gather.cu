// Microbenchmark: how fast can one SM gather row segments of a big row-major matrix into shared
// memory?  (a) 1-D TMA bulk copies of SZ bytes, depth D in flight, issued by one warp;
// (b) LDGSTS 16 B/lane from W warps.  Rows are random, segments walk along the row.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned par) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one warp per CTA drives the TMA: stage = group of R row copies of SZ bytes; NS stages in flight
template <int SZ, int R, int NS>
__global__ void __launch_bounds__(32, 1) tma_kernel(const double *A, long long ld, const int *rows, int nrows, long long items, double *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + (size_t) NS * R * SZ);
    const int lane = threadIdx.x;
    if (lane == 0) { for (int i = 0; i < NS; i++) mbar_init(smem_u32(bars + i), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncwarp();
    const long long per = items / gridDim.x;
    const long long segs = ld * 8 / SZ;
    double acc = 0;
    for (long long k = 0; k < per + NS; k++) {
        if (k >= NS) {   // consume item k - NS
            const int st = (int) ((k - NS) % NS);
            mbar_wait(smem_u32(bars + st), (unsigned) (((k - NS) / NS) & 1));
            acc += reinterpret_cast<double *>(sm + (size_t) st * R * SZ)[lane];
            __syncwarp();
        }
        if (k < per) {
            const int st = (int) (k % NS);
            const long long it = blockIdx.x * per + k;
            const long long seg = it % segs;
            if (lane == 0) mbar_expect(smem_u32(bars + st), R * SZ);
            __syncwarp();
            for (int r = lane; r < R; r += 32) {
                const int row = rows[(it * 7 + r * 13) % nrows];
                bulk(smem_u32(sm + (size_t) st * R * SZ + (size_t) r * SZ), A + (long long) row * ld + seg * (SZ / 8), SZ, smem_u32(bars + st));
            }
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}
// LDGSTS: W warps per CTA, each stages 64 row segments of 256 B (2 rows per instruction), waits, repeats
template <int RW>
__global__ void __launch_bounds__(1024, 1) ldgsts_kernel(const double *A, long long ld, const int *rows, int nrows, long long items, double *sink, int cg) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    unsigned char *S = sm + (size_t) warp * RW * 272;
    const long long per = items / ((long long) gridDim.x * W);
    const long long segs = ld / 32;
    const int half = lane >> 4, l16 = lane & 15;
    double acc = 0;
    for (long long k = 0; k < per; k++) {
        const long long it = ((long long) blockIdx.x * W + warp) * per + k;
        const long long seg = it % segs;
        #pragma unroll 8
        for (int r = 0; r < RW / 2; r++) {
            const int row = rows[(it * 7 + (r * 2 + half) * 13) % nrows];
            const double *src = A + (long long) row * ld + seg * 32 + l16 * 2;
            const unsigned dst = smem_u32(S + (size_t) (r * 2 + half) * 272 + l16 * 16);
            if (cg) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        acc += reinterpret_cast<double *>(S)[lane];
        __syncwarp();
    }
    if (acc == 12345.678) sink[0] = acc;
}
// plain LDG.128 into registers: W warps, each lane 16 loads of 16 B in flight (2 rows per instruction)
__global__ void __launch_bounds__(1024, 1) ldg_kernel(const double *A, long long ld, const int *rows, int nrows, long long items, double *sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const long long per = items / ((long long) gridDim.x * W);
    const long long segs = ld / 32;
    const int half = lane >> 4, l16 = lane & 15;
    double acc = 0;
    for (long long k = 0; k < per; k++) {
        const long long it = ((long long) blockIdx.x * W + warp) * per + k;
        const long long seg = it % segs;
        for (int b = 0; b < 2; b++) {
            double2 v[16];
            #pragma unroll
            for (int r = 0; r < 16; r++) {
                const int row = rows[(it * 7 + ((b * 16 + r) * 2 + half) * 13) % nrows];
                v[r] = __ldg(reinterpret_cast<const double2 *>(A + (long long) row * ld + seg * 32 + l16 * 2));
            }
            #pragma unroll
            for (int r = 0; r < 16; r++) acc += v[r].x + v[r].y;
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}
template <int SZ, int R, int NS>
void run_tma(const double *A, long long ld, const int *rows, int nrows, double *sink) {
    const size_t smem = (size_t) NS * R * SZ + NS * 8;
    cudaFuncSetAttribute(tma_kernel<SZ, R, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    const long long items = 148LL * 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tma_kernel<SZ, R, NS><<<148, 32, smem>>>(A, ld, rows, nrows, items / 10, sink);
    cudaEventRecord(e0);
    tma_kernel<SZ, R, NS><<<148, 32, smem>>>(A, ld, rows, nrows, items, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("TMA bulk  SZ=%5d R=%3d NS=%2d (%.0f KB in flight/SM): %.3f ms, %.0f GB/s, %.1f cycles/op/SM %s\n", SZ, R, NS, NS * R * SZ / 1024.0, ms,
           (double) items * R * SZ / ms / 1e6, ms * 1e-3 * 1.9e9 / ((double) items / 148 * R), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const long long ld = 40000, nr = 24000;
    double *A, *sink; int *rows;
    cudaMalloc(&A, ld * nr * 8); cudaMemset(A, 0, ld * nr * 8); cudaMalloc(&sink, 8);
    std::vector<int> h(65536); srand(1); for (auto &x : h) x = rand() % nr;
    cudaMalloc(&rows, h.size() * 4); cudaMemcpy(rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int nrows = (int) h.size();
    run_tma<256, 64, 12>(A, ld, rows, nrows, sink);
    run_tma<512, 64, 6>(A, ld, rows, nrows, sink);
    run_tma<1024, 32, 6>(A, ld, rows, nrows, sink);
    run_tma<2048, 16, 6>(A, ld, rows, nrows, sink);
    run_tma<4096, 8, 6>(A, ld, rows, nrows, sink);
    run_tma<8192, 4, 6>(A, ld, rows, nrows, sink);
    run_tma<2048, 16, 3>(A, ld, rows, nrows, sink);
    run_tma<2048, 32, 3>(A, ld, rows, nrows, sink);
    // LDGSTS: what it sustains depends on the rows per item (stage size), the warps per SM and the L1 the
    // shared-memory footprint leaves (B200: 64 rows x 12 warps = 204 KB: 2.0 TB/s; 32 x 16 = 136 KB: 5.5; 16 x 32: 6.5)
    auto run_l = [&](auto kern, int RW, int W, int cg) {
        const size_t smem = (size_t) W * RW * 272;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        const long long items = 148LL * W * 1500 * 64 / RW;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        kern<<<148, W * 32, smem>>>(A, ld, rows, nrows, items / 10, sink, cg);
        cudaEventRecord(e0);
        kern<<<148, W * 32, smem>>>(A, ld, rows, nrows, items, sink, cg);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("LDGSTS.%s rows/item=%2d W=%2d warps/SM (%3zu KB smem): %.3f ms, %.0f GB/s  %s\n", cg ? "cg" : "ca", RW, W, smem / 1024, ms,
               (double) items * RW * 256 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int W : { 8, 12 }) run_l(ldgsts_kernel<64>, 64, W, 1);
    for (int W : { 8, 12, 16, 24 }) run_l(ldgsts_kernel<32>, 32, W, 1);
    for (int W : { 12, 16, 24, 32 }) run_l(ldgsts_kernel<16>, 16, W, 1);
    for (int W : { 8, 16, 32 }) {
        const long long items = 148LL * W * 1500;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        ldg_kernel<<<148, W * 32>>>(A, ld, rows, nrows, items / 10, sink);
        cudaEventRecord(e0);
        ldg_kernel<<<148, W * 32>>>(A, ld, rows, nrows, items, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("LDG.128 W=%2d warps/SM (16 in flight/lane): %.3f ms, %.0f GB/s  %s\n", W, ms, (double) items * 16384 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
