// Device-side interface of the kinship kernels (sm_100a).  See gk_kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gk {

// Everything one recurrence step needs, by value.
struct StepDev {
    const double *Gprev;     // K_prev x ld_prev class matrix of cut g-1 (row-major, symmetric)
    const double *dprev;     // self kinship of each class of cut g-1
    double *Gnew;            // K x ld class matrix of cut g
    double *dnew;            // self kinship of each class of cut g
    int64_t ld_prev, ld_new;
    int64_t K_prev, K;
    const int32_t *desc;     // nt tile descriptors of kDescWords int32 (layout: gk_plan.hpp)
    int64_t npairs;          // nt*(nt+1)/2 unordered tile pairs
    int32_t nt;
    int32_t prefetch;        // 1: stream the list rows of upcoming tile rows into L2
};

struct ExpandDev {
    const double *G;         // K_last x ld last class matrix
    const double *d;         // self kinship per class
    int64_t ld;
    const int32_t *cls;      // m: class row of caller entry j
    const int32_t *ind;      // m: individual of caller entry j
    const int32_t *cact;     // nact: classes that have at least one caller entry among the local rows
    const int32_t *coff;     // nact + 1: CSR into cmem
    const int32_t *cmem;     // local caller entries (row indices i, row0 <= i < row1) grouped by class
    double *out;             // (row1 - row0) x m, row-major, caller order
    int64_t m, row0, row1;
    int32_t nact, nchunks;
    double *partials;        // nact*nchunks*2 : per-CTA (sum all, sum diag)
};

size_t step_smem_bytes();
int step_prefetch_default();
cudaError_t launch_step(const StepDev &s, bool ranked, cudaStream_t st);
cudaError_t launch_init(double *G0, double *d0, int64_t K0, int64_t ld0, cudaStream_t st);
constexpr int kExpandCols = 1024;     // result columns per CTA of the expansion kernel
cudaError_t launch_expand(const ExpandDev &e, cudaStream_t st);
cudaError_t launch_eye(double *out, int64_t m, int64_t row0, int64_t row1, double *partials, cudaStream_t st);
// two-stage deterministic reduction of the per-CTA partials: scratch holds 2*256 doubles
cudaError_t launch_mean_finish(const double *partials, int64_t nblocks, double *scratch, double *sums /*2*/, cudaStream_t st);
cudaError_t configure_kernels();   // cudaFuncSetAttribute for large dynamic shared memory

// gen.gc propagation: out row = 0.5*w0*row(p0) + 0.5*w1*row(p1), then seed own ancestor columns.
struct GcStepDev {
    const double *Vprev; double *Vnew;
    int64_t ld;              // columns (ancestors, padded)
    int64_t rows;            // members of the new cut
    const int32_t *p0, *p1;  // parent row in Vprev or -1 (kept: p0 = own old row, p1 = -2)
    const int32_t *seed_off; // rows + 1 : CSR of ancestor columns equal to this individual
    const int32_t *seed_col;
};
cudaError_t launch_gc_step(const GcStepDev &g, cudaStream_t st);
cudaError_t launch_gc_gather(const double *V, int64_t ld, const int32_t *row_of /*m, -1 = zero row*/,
                             double *out, int64_t m, int64_t a, cudaStream_t st);

}  // namespace gk
