// Device-side interface of the kinship kernels (sm_100a).  See gk_kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gk_plan.hpp"

namespace gk {

constexpr int kMaxShards = 8;     // GPUs of one box
constexpr int kRingMax = 40;      // L2-resident panel buffers (transposed N panels), at most
constexpr int kStepWarps = 12;     // warps per CTA of the step kernel (one CTA per SM), each with a private 17 KB stage
constexpr int kStepP2Warps = 7;      // multi-GPU step kernel (phi_step_tma_kernel): phase-2 warps (LDGSTS)
constexpr int kStepP1Consumers = 6;  // phase-1 consumer warps
constexpr int kStepP1Stages = 3;     // phase-1 TMA stages (33 KB each)
constexpr int kP1Classes = 8;        // phase-1 item: 8 classes of a panel ...
constexpr int kP1Cols = 256;         // ... x 256 columns of the previous matrix (2 KB row segments)
constexpr int kCB = 32;           // columns of the previous matrix per phase-1 work item
// third step kernel (phi_step_g4_kernel, one GPU): phase 1 through TMA bulk copies of 4 KB row segments, phase 2 straight from L2
constexpr int kG4Stages = 2;         // TMA stages (a FIFO the producer fills and all consumer warps drain together)
constexpr int kG4Consumers = 4;      // phase-1 consumer warps
constexpr int kG4P2Warps = 10;       // phase-2 warps (ring rows straight from L2 into registers)
constexpr int kG4StageBytes = kUnitStageBytes;  //   // 20 rows x 512 columns (pitch 514), or up to 39 rows x 256 columns (pitch 258)
constexpr int kG4RowsMax = kUnitRowsMax; //      // distinct parent rows of a UNIT of 16 classes (<= 32) + an all-zero row, padded

// Where the rows of a (possibly row-sharded) class matrix live: row r is owned by the shard q with
// row0[q] <= r < row0[q+1] and sits at base[q] + (r - row0[q]) * ld.  Peer shards are CUDA-IPC
// mappings of the other ranks' buffers (NVLink); a single GPU has n = 1.
struct RowTable {
    double *base[kMaxShards];
    int32_t row0[kMaxShards + 1];
    int32_t n;
};

// Everything one recurrence step needs, by value.
struct StepDev {
    RowTable prev;            // class matrix of cut g-1: Kpad_prev rows x ld_prev
    RowTable outt;            // every rank's rows of the class matrix of cut g (mirror stores of a ranked multi-GPU job)
    double *out_base;         // own rows of the class matrix of cut g (Kpad columns): global row r at out_base + (r - out_row0) * ld_new
    int64_t out_row0;
    int64_t ld_prev, ld_new;
    int32_t Kpad_prev;
    int32_t np;               // panels / tiles of the new cut (global)
    const int32_t *pF, *pM;   // Kpad: parent rows of every new class in the previous matrix (-1 none)
    const int32_t *first_use; // Kpad_prev: lowest tile of this cut that reads row c of a transposed panel (np: none)
    const int32_t *prow, *pcnt, *slots;   // per panel: its distinct parent rows (64 slots), their count; per class: father / mother slot
    const int32_t *prow8, *pcnt8, *slots8; // the same per sub-panel of 8 classes (16 slots): phase-1 items of phi_step_tma_kernel
    const double *zero_row;   // >= 512 zeros (the parent row of a class with an unknown parent)
    double *ring;             // ring_n x ring_stride doubles: N^T panels, [Kpad_prev][32]
    int64_t ring_stride;
    int32_t ring_n;
    int32_t *done1, *done2;   // np counters each: finished phase-1 / phase-2 items of a panel
    // phi_step_tma_kernel only: two item streams instead of one queue
    const int64_t *grp_off2;  // own panels + 1: first phase-2 item of each own panel
    int64_t total1, total2;   // phase-1 / phase-2 items of this launch
    int32_t n1pp;             // phase-1 items per panel = 4 * ncb256
    int32_t ncb256;           // 256-column blocks of the previous matrix
    const int64_t *grp_off;   // ngroups + 1: first work item of each group
    const int32_t *grp_info;  // ngroups: panel | (phase - 1) << 30
    int64_t total_items;
    int32_t ngroups;
    int32_t ncb;              // phase-1 items per panel
    int32_t panel_begin, panel_end;   // panels this launch (rank) owns
    int32_t full_rows;        // 0: tiles J <= I, mirror stored too; 1: every tile J, own rows only
    int32_t out_policy;       // L2 eviction priority of the stores to the new matrix: 0 evict_first, 1 normal, 2 evict_last
    int32_t debug_nowait;     // timing experiments only (GK_DEBUG_NOWAIT=1): skip the dependency waits, results are garbage
    // phi_step_g4_kernel only
    const int32_t *prow4;     // 2 np x kG4RowsMax: per UNIT (16 classes) its distinct parent rows, then the zero row (index Kpad_prev)
    const int32_t *pcnt4;     // 2 np: staged positions | rows copied << 8 | (W == 512) << 16
    const int32_t *slots4;    // Kpad: staged row of the father | staged row of the mother << 8 (the zero row for an unknown parent)
    const int64_t *item_off;  // 2 np + 1: first phase-1 item of each unit (a unit has ceil(Kpad_prev / W) items, W = 512 or 256 columns)
    int64_t total_g4;         // phase-1 items of this launch
    int32_t chunk2;              // tiles per phase-2 queue grab (GK_CHUNK2, default 2)
    unsigned long long *queue2;  // phase-2 work queue of this launch: next chunk of tiles (zeroed with the counters)
    unsigned long long *prof; // optional (GK_PROF=1): 16 cycle counters summed over the warps of each role, see gk_phi_debug_prof
};

struct DiagDev {
    RowTable prev, out;
    int64_t ld_prev, ld_new;
    const double *dprev;      // self kinship per class of cut g-1 (replicated)
    double *dnew;             // self kinship per class of cut g (this launch writes [c_begin, c_end))
    double *dnew_peer[kMaxShards];   // one-process multi-GPU jobs: the other ranks' replicas of that vector (peer stores, 8 B per class)
    int32_t n_peer;
    const int32_t *pF, *pM, *flags;
    int32_t c_begin, c_end;   // classes whose rows this launch owns
    const int32_t *pg_i, *pg_j, *pg_off, *pt_cls, *pt_mult;
    int32_t ngroups;
};

struct ExpandDev {
    const double *G;         // K_last x ld last class matrix (expand_sym_kernel of a sharded job: the fetched rows TRANSPOSED, Kpad x ld)
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
    int32_t a_begin;         // first active class of this launch
    int32_t compact;         // 1: G holds only the active classes' rows, row a <-> cact[a] (multi-GPU)
    double *partials;        // nact*nchunks*2 : per-CTA (sum all, sum diag)
};

int step_grid();             // CTAs of the persistent step kernel (SMs x resident CTAs)
cudaError_t launch_step(const StepDev &s, cudaStream_t st);
cudaError_t launch_step_tma(const StepDev &s, cudaStream_t st);
cudaError_t launch_step_g4(const StepDev &s, cudaStream_t st);
cudaError_t launch_diag(const DiagDev &d, cudaStream_t st);
cudaError_t launch_init(double *G0, double *d0, const int32_t *flags, int64_t K0, int64_t Kpad0, cudaStream_t st);
cudaError_t launch_init_rows(double *G0, double *d0, const int32_t *flags, int64_t K0, int64_t Kpad0,
                             int64_t row_begin, int64_t rows, int64_t dlen, cudaStream_t st);
// ranked multi-GPU jobs: this rank's tiles (J own, I >= panel_end) = transposes of the tiles in the rows of I's owners
cudaError_t launch_pull_mirror(const RowTable &t, int64_t ld, int32_t pb, int32_t pe, int32_t np, double *own, int64_t own_row0, cudaStream_t st);
cudaError_t launch_fetch_rows(const RowTable &t, int64_t ld, const int32_t *rows, int32_t first, int32_t count /* positions in the list */, double *dst, cudaStream_t st);
constexpr int kExpandCols = 2048;     // result columns per CTA of the expansion kernel
cudaError_t launch_expand(const ExpandDev &e, int32_t a_begin, int32_t a_count /* range of the active classes */, cudaStream_t st);
// one GPU, every class active (cact = identity): 8 classes x 1024 columns per CTA, 64-byte segments of the symmetric matrix
cudaError_t launch_expand_sym(const ExpandDev &e, int32_t a_begin /* multiple of 8 */, int32_t a_count, cudaStream_t st);
// sharded jobs: GT[c][a] = G[rows[a]][c], the active classes' rows fetched from their owners and transposed (G of expand_sym_kernel, ld = ldt)
cudaError_t launch_fetch_rows_t(const RowTable &t, int64_t ld, const int32_t *rows, int32_t a_begin, int32_t a_count, int32_t nact,
                                int32_t column_panels, double *GT, int64_t ldt, cudaStream_t st);
int64_t expand_sym_blocks(int32_t nact, int64_t m);     // its grid = its (sum, diagonal sum) partials
cudaError_t launch_eye(double *out, int64_t m, int64_t row0, int64_t row1, double *partials, cudaStream_t st);
// two-stage deterministic reduction of the per-CTA partials: scratch holds 2*256 doubles
cudaError_t launch_mean_finish(const double *partials, int64_t nblocks, double *scratch, double *sums /*2*/, cudaStream_t st);
cudaError_t configure_kernels();

// device-side consumers of the m x m result: pair selection (gen.phiOver, sparse CSR) and gen.phiCI's bootstrap statistic
struct SelectDev {
    const double *M;          // rows [row0, row1) of the result, m columns
    int64_t m, row0, row1;
    int32_t mode;             // 0: j < i and value > thr (phiOver); 1: j <= i and |2 value| > thr (sparse CSR, float values)
    double thr;
    int64_t *row_count;       // pass 1 out
    const int64_t *row_off;   // pass 2 in: exclusive prefix of row_count
    int32_t *oi, *oj;
    double *ov;
    float *ovf;
};
cudaError_t launch_select(const SelectDev &q, int fill, cudaStream_t st);
cudaError_t launch_ci(const double *M, int64_t ld, const int32_t *idx /* b x n */, int32_t b, int32_t n, double *stat /* b */, cudaStream_t st);

// gen.gc propagation: out row = 0.5*w0*row(p0) + 0.5*w1*row(p1), then seed own ancestor columns.
struct GcStepDev {
    const double *Vprev; double *Vnew;
    int64_t ld;              // columns (ancestors, padded)
    int64_t rows;            // members of the new cut
    const int32_t *p0, *p1;  // parent row in Vprev or -1 (kept: p0 = own old row, p1 = -2)
    const int32_t *seed_off; // rows + 1 : CSR of ancestor columns equal to this individual
    const int32_t *seed_col;
    int32_t col0, ncols;     // this launch propagates the ancestor columns [col0, col0 + ncols) (column blocks of a wide ancestor list)
};
cudaError_t launch_gc_step(const GcStepDev &g, cudaStream_t st);
cudaError_t launch_gc_gather(const double *V, int64_t ld, const int32_t *row_of /*m, -1 = zero row*/,
                             double *out, int64_t m, int64_t a, cudaStream_t st);

}  // namespace gk
