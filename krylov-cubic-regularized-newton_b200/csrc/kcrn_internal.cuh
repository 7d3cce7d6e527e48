// Internal declarations shared by the translation units of libkcrn.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "kcrn.h"

namespace kcrn {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
#define KCRN_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            kcrn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return KCRN_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)
#define KCRN_TRY(call)                 \
    do {                               \
        int r_ = (call);               \
        if (r_ != KCRN_OK) return r_;  \
    } while (0)
#define KCRN_LAUNCH_CHECK() KCRN_CUDA(cudaGetLastError())

// Kernel launch with (pdl = true) or without the programmatic-dependent-launch attribute.
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ----------------------------------------------------------------------------- process switches
// Every KCRN_* environment switch is read in ONE place, read_switches() in kcrn_api.cu, once per
// handle at kcrn_create (kcrn_plan_tiles, which has no handle, reads them per call); nothing else calls getenv.
struct Switches {
    bool verbose = false;          // KCRN_VERBOSE
    bool no_pdl = false;           // KCRN_NO_PDL: no programmatic dependent launch inside the captured graphs
    bool no_defer = false;         // KCRN_NO_DEFER: grid-wide sums of the fused recurrence finished by a last-CTA tail
    bool emulate_sharded = false;  // KCRN_EMULATE_SHARDED: a one-GPU handle takes the row-sharded code path
    bool no_p2p = false;           // KCRN_NO_P2P: NCCL allreduce even where the NVLink peer exchange fits
    bool no_gather_order = false;  // KCRN_NO_GATHER_ORDER: sweeps read the entries as stored
    bool natural_csc = false;      // KCRN_NATURAL_CSC: CSC copy in the scipy column order (no length classes)
    int gather_plain = -1, gather_lanczos = -1;   // KCRN_GATHER=p[,l]: forced gather variants (-1: measure)
    int short_seg = 0, tile_items = 0;            // KCRN_SHORT_SEG, KCRN_TILE_ITEMS: tiling knobs (0: default)
    bool persist = false, persist_trace = false;  // KCRN_PERSIST, KCRN_PERSIST_TRACE: one-kernel recurrence (opt-in)
    int persist_flags = 0;                        // KCRN_PERSIST_FLAGS: its measured experiments
    int dshard = -1;                              // KCRN_DSHARD=1 / 0: d-sharded re-orthogonalisation on / off (-1: by size)
    bool no_reorth_reg = false;                   // KCRN_NO_REORTH_REG: re-orthogonalisation stages 1, 2 by the general kernel (A/B)
    bool force_reorth_reg = false;                // KCRN_REORTH_REG: ... by the register kernel whatever d is (tests)
};
Switches read_switches();

// ----------------------------------------------------------------------------- geometry
constexpr int kTileNnz = 2048;     // non-zeros per CTA tile at the default 8 per thread (11 or 12 per thread for small matrices, see build_tiles)
constexpr int kTileSegs = 1024;    // at most this many segments (rows / columns) per stream tile
constexpr int kShortSeg = 1024;    // longest segment that shares a stream tile with others
constexpr int kSlotBits = 12;      // gather-ordered words: low bits = position inside the tile (tiles hold <= 3072 entries)
constexpr int kSingleTag = 15;     // descriptor tag of a tile that holds one medium segment
constexpr int kSweepWarps = 8;     // warps per CTA of the sweep kernels
constexpr int kSweepThreads = kSweepWarps * 32;
constexpr int kVecThreads = 256;   // CTA size of the d-length vector kernels
constexpr int kMaxM = 512;         // largest Lanczos subspace the single-CTA solver accepts
constexpr int kNumSm = 148;        // B200
constexpr int kMaxPeers = 16;      // ranks of one NVLink domain that can share exchange buffers

// One sparse matrix stored by segments: rows (CSR) or columns (CSC).
//   tiles[i] = {seg0, k0, z, nk | tag << 16}  (one CTA per tile):
//     z > 0, tag = g < 6 : "stream" tile -- z whole short segments seg0..seg0+z whose non-zeros are
//             [k0, k0+nk); every group of 2^g threads reduces one segment
//     z = 1, tag = 15    : one medium segment, reduced by the whole CTA
//     z < 0 : one 2048-non-zero piece [k0, k0+nk) of the long segment seg0; -z-1 indexes longtab
//   longtab[l] = {first partial slot, number of pieces, ptr[seg], seg}
struct SegMat {
    int64_t nseg = 0, nnz = 0;
    int32_t *ptr = nullptr, *idx = nullptr;
    double *val = nullptr;
    int4 *tiles = nullptr;
    int ntiles = 0;
    int items = 8;               // non-zeros per thread of the tiles: 8, 11 or 12 (tile = items x 256)
    int4 *longtab = nullptr;
    int nlong = 0, npartial = 0;
    double *partial = nullptr;   // [npartial] piece sums of long segments
    double *longdot = nullptr;   // [nlong]    v.w contribution of each long segment (Lanczos)
    unsigned *ticket = nullptr;  // [nlong]    self-resetting arrival counters
    int grid = 0;                // CTAs of the sweep kernels over this matrix (fixed -> deterministic sums)
    int64_t gather_len = 0;      // length of the vector this matrix gathers from (d for CSR, n for CSC)
    // Gather-ordered copy used by the sweeps (build_tiles): inside every stream tile the non-zeros are sorted by
    // the index they gather with, word = index << kSlotBits | position the entry has in the tile's natural order.
    // The lanes of a warp then ask L1TEX for far fewer distinct 128-byte lines per load (rcv1 shape: 31.5 -> 15.7
    // rows, 23.0 -> 11.6 columns); the products go back to their natural positions in shared memory, so every sum
    // keeps its order and its bits.  Null when gather_len does not fit the word (sweeps then use idx / val).
    uint32_t *gidx = nullptr;    // [nnz]
    double *gval = nullptr;      // [nnz] values in the same order
    // which gather variant of the sweep kernel (kcrn_sweep.cu: GatherVariant) serves this matrix, measured at
    // set-up: one choice for the Lanczos passes, one for all others
    int gather_plain = 0, gather_lanczos = 0;
    // CSC copy only: its columns are stored in a length-class order (kcrn_transpose.cu)
    int32_t *segid = nullptr;    // [nseg] original column of the s-th stored segment (null: natural order)
    int32_t *pos = nullptr;      // [nseg] where column c is stored
    std::vector<int32_t> cptr_host;   // canonical column pointer (natural order), for kcrn_csc_view
};

// ----------------------------------------------------------------------------- persistent recurrence (kcrn_persist.cu)
// Small problems (every pass over A is at most one tile per SM): the whole Lanczos recurrence runs as ONE kernel of
// one 1024-thread CTA per SM; kernel boundaries become grid barriers, the (constant) tile streams arrive by bulk
// copy (cp.async.bulk + mbarrier) while the previous pass finishes, the recurrence state lives in registers.
constexpr int kPThreads = 1024;
constexpr int kPSlotBits = 14;                       // word = gather index << 14 | position inside the tile
constexpr int kPCap = 12288;                         // entries per tile (capacity; tiles are balanced: about nnz / SMs)
constexpr int kPSegs = 4096;                         // segments per tile
constexpr int kPSplit = 1024;                        // segments longer than this may be cut at a tile boundary
constexpr int kPBig = 1024;                          // local segments longer than this are summed by the whole CTA
constexpr int kPMaxIndex = 1 << (32 - kPSlotBits);   // gather length the word can address

struct PTile {             // one tile = the entries [e0, e0 + nk) of a matrix in stored order
    int e0, nk;
    int seg0, nseg;        // stored segments seg0 .. seg0 + nseg - 1 intersect it (first / last possibly as pieces)
    int lg;                // log2 of the number of threads that share one short segment
    int head_long, head_piece;   // >= 0: local segment 0 is piece `head_piece` of cut segment `head_long` (longtab)
    int tail_long, tail_piece;   // the same for local segment nseg - 1
    int pad[3];            // [0], [1]: slot of the head / tail piece among the matrix's piece sums
};
static_assert(sizeof(PTile) == 48, "tile descriptor = three 16-byte words");

struct PersistMat {
    PTile *tiles = nullptr;
    int ntiles = 0;
    uint32_t *word = nullptr;    // [ntiles][kPCap] words of the tile in gather order
    double *tval = nullptr;      // [ntiles][kPCap] values in the same order
    int4 *longtab = nullptr;     // cut segments: {first piece slot, pieces, stored segment, 0}
    int nlong = 0, npartial = 0; // cut segments, pieces
};

// Device-resident state of the outer loop (kcrn_krylov_run): what Cubic_Krylov_LS carries from one iteration to
// the next (cubic.py:248-262) and the loop counters.  The first 10 doubles are filled by the host before a run.
struct RunCtl {
    double value, reg_coef, r0, beta, tol, f_opt;   // running value f(x), M estimate, r0, backtracking factor, tolerance, best value
    double max_iters_d, max_tries_d, pad0, pad1;
    double M;                                        // M of the trial in flight
    unsigned long long t0;                           // %globaltimer at the start of the run
    long long solver_it;
    int iter, max_iters, tries, max_tries, status, copy_best, improved, pad2;
};

// Device-resident control block of the Lanczos recurrence (written by single threads on the device,
// read by every kernel of the captured graph; no host involvement).
struct LanczosCtl {
    int col;        // index of the current Lanczos vector
    int redo;       // 1: the next operator application re-uses V[col] as is (after a breakdown)
    int final_;     // 1: the next operator application only produces alphas[store_idx]
    int done;       // 1: recurrence finished, remaining graph nodes are no-ops
    int m_eff;      // number of alphas the reference would return
    int nvalid;     // number of basis columns written
    int store_idx;  // where the final alpha goes
    int m;
    int broke;
    int hvps;       // operator applications consumed
    double last_beta;
    // the two scalars every kernel of a step needs, kept in the same 64-byte line as the indices above so that one
    // round trip to L2 fetches everything (bet[col] / alphas[col] would be a second, dependent one)
    double cur_beta;    // = bet[col]: the norm that normalises column col (|g| for col 0)
    double cur_alpha;   // = alphas[col] once the row sweep of the step has produced it
};
static_assert(sizeof(LanczosCtl) == 64, "control block = one 64-byte line");

#ifdef __CUDACC__
// Bookkeeping of cubic.py:98-109 once beta = |w'| of column `col` is known (executed by ONE thread).
__device__ inline void lanczos_advance(LanczosCtl *c, double *bet, double beta) {
    const int col = c->col, m = c->m;
    c->last_beta = beta;
    if (fabs(beta) < 1e-6) {               // cubic.py:98 -- absolute test
        c->broke = 1;
        if (col < m - 2) { c->m_eff = col + 1; c->store_idx = col; }   // cubic.py:105-108
        else { c->m_eff = m; c->store_idx = m - 1; }
        c->redo = 1;                       // cubic.py:109 re-applies the operator to the same v
        c->final_ = 1;
    } else {
        if (bet) bet[col + 1] = beta;
        c->cur_beta = beta;
        c->col = col + 1;
        c->nvalid = col + 2;
        if (col + 1 == m - 1) { c->final_ = 1; c->store_idx = m - 1; }
    }
}
#endif

}  // namespace kcrn

struct kcrn_problem {
    int64_t n = 0, d = 0, nnz = 0, n_global = 0;
    double l2 = 0.0;
    cudaStream_t stream = nullptr;
    kcrn::SegMat csr, csc;
    double *b = nullptr;
    // caches at the current x
    bool have_x = false;
    double *x_cur = nullptr;                       // [d]
    double *x_base = nullptr;                      // [d] iterate the trial points of the running step start from
    double *step_params = nullptr;                 // [4] {M, r0, eps, it_max} of the next sub-problem solve (device)
    bool have_base = false;
    // device-resident outer loop (kcrn_krylov_run)
    kcrn::RunCtl *run_ctl = nullptr;
    double *run_recs = nullptr;     // [run_cap][16] one record per outer iteration
    double *run_xs = nullptr;       // [run_cap][d] accepted iterates, when asked for
    double *run_x_opt = nullptr;    // [d] best point seen (loss.py:70-72)
    double *run_host = nullptr;     // pinned: parameters in, records out
    int run_cap = 0, run_xs_cap = 0;
    int64_t run_delta[3][8] = {};   // counters of: once per run, once per outer iteration, once per trial
    // batched values of stored iterates (kcrn_values_batch)
    double *batch_xt = nullptr, *batch_scal = nullptr, *batch_partial = nullptr;
    unsigned *batch_ticket = nullptr;
    int batch_grid = 0;
    double *Ax = nullptr, *D = nullptr, *resid = nullptr, *ell = nullptr;  // [n]
    // per-tile sums of the loss terms left by the forward sweep (+ the shares of long rows), added in tile order by
    // value_record_kernel; invalid after kcrn_set_Ax / kcrn_update_Ax (those derive the terms without a sweep)
    double *fwd_part = nullptr, *fwd_longdot = nullptr;
    bool fwd_parts_valid = false;
    // config 4: dense panels of the selected columns for kcrn_partial_hessian (kept across SSCN steps)
    double *panel_x = nullptr, *panel_y = nullptr;
    size_t panel_bytes = 0;
    double *splitk_part = nullptr;     // per-slice accumulators of the split-K block Hessian
    size_t splitk_bytes = 0;
    // sweep temporaries
    double *t = nullptr;        // [n]   D .* (A v)
    double *ypart = nullptr;    // [d]   local A' t before the allreduce (row-sharded runs)
    // Lanczos state
    int m_cap = 0;
    double *V = nullptr;        // [m_cap * d]
    double *w = nullptr, *wp = nullptr, *g = nullptr;  // [d]
    double *alphas = nullptr, *bet = nullptr, *coef = nullptr, *coef2 = nullptr;  // [kMaxM + 1]
    kcrn::LanczosCtl *ctl = nullptr;               // [2]: deferred mode ping-pongs between them, everything else uses [0]
    double *lan_rows_part = nullptr, *lan_cols_part = nullptr;   // per-CTA pieces of alpha (2 x csr.grid) / beta^2 (csc.grid)
    double *post_part = nullptr;                                 // row-sharded fused loop: per-CTA pieces of beta^2 left by the epilogue kernel
    int32_t *info = nullptr;    // [8]
    int cur_m = 0, cur_reorth = 0;
    // sub-problem result {s[kMaxM], r, model_decrease, iterations, status}
    double *sub = nullptr;
    double *scal = nullptr;     // [16] device scalars: value, |x|^2, ...
    double *rec_host = nullptr; // pinned [16]: step record [12] + trial parameters [4]
    cudaEvent_t ev_trial = nullptr;   // behind the graph of a trial: what kcrn_krylov_trial waits for
    // reductions
    double *red_partials = nullptr;
    unsigned *red_ticket = nullptr;   // [8]
    int vec_grid = 0, vec_elems = 1;
    // device staging for the host-buffer entry points
    double *x_stage = nullptr, *v_stage = nullptr, *o_stage = nullptr;  // [d] device
    // captured Lanczos graphs, keyed by (m, reorth)
    struct GraphEntry { int kind, m, reorth; cudaGraphExec_t exec; int64_t delta[8]; };   // kind 0: Lanczos, 1 / 2: first / repeated trial
    std::vector<GraphEntry> graphs;
    int64_t graph_delta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int step_col = 0;           // column the step-wise Lanczos API is working on
    // NCCL
    void *comm = nullptr;
    int rank = 0, world = 1;
    // NVLink peer exchange of the fused Lanczos step (kcrn_p2p_export / kcrn_p2p_attach)
    uint4 *p2p_recv = nullptr;          // [2][world][d + 2] slots {value.lo, step, value.hi, step} the peers push into
    unsigned *p2p_epoch = nullptr;      // exchange steps done so far (device counter; picks the buffer half)
    uint4 *peer_recv[kcrn::kMaxPeers] = {};
    bool p2p_ready = false;
    int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // KCRN_EMULATE_SHARDED=1 (read at kcrn_create): a single-GPU handle takes the code path of a row-sharded one --
    // raw transpose sweep, (identity) sum over the ranks, separate epilogue kernels -- so that this path is covered by
    // the one-GPU test suite
    bool emulate_sharded = false;
    bool sharded() const { return world > 1 || emulate_sharded; }
    // d-sharded re-orthogonalisation: for row-sharded handles whose basis is large enough that 1/world of its traffic
    // outweighs four NCCL collectives per step (KCRN_DSHARD=1 / 0 forces it on / off).  Measured at 8 GPUs
    // (profiles/r2s_dshard_*.json): a small NCCL allreduce costs ~44 us there, so at d = 1 M, m = 100 the two ways tie
    // (39.1 vs 36.3 ms of re-orthogonalisation per recurrence) and at the news20 shape the sharded one loses; the
    // default threshold sits above both
    int cur_m_enqueue = 0;         // m of the recurrence being enqueued
    bool dshard_reorth() const {
        if (!sharded()) return false;
        if (sw.dshard >= 0) return sw.dshard != 0;
        return world > 1 && (double)d * (double)cur_m_enqueue > 2.0e8;
    }
    // persistent recurrence (kcrn_persist.cu); persist_ok: both matrices fit one tile per SM
    kcrn::PersistMat prow, pcol;
    bool persist_ok = false;
    int persist_grid = 0;
    double *gp = nullptr, *wpp = nullptr, *Vp = nullptr;   // gradient, w', basis in the stored column order of the CSC copy
    uint4 *pslots = nullptr;       // exchange slots {value.lo, epoch, value.hi, epoch}: [2][2][grid] per CTA, [2][piece_cap] per piece
    int piece_cap = 0;
    unsigned *pbar = nullptr;      // [0] exchanges done so far (epoch of the slots)
    long long *ptrace = nullptr;   // KCRN_PERSIST_TRACE=1: SM clock at the marks of the last launch
    kcrn::Switches sw;             // environment switches as they were at kcrn_create
    unsigned long long *sweep_trace = nullptr;   // -DKCRN_SWEEP_TRACE builds only: %globaltimer marks of the Lanczos passes
    bool lanczos_tuned = false;    // the gather variants of the Lanczos passes have been measured (first use)
    bool pdl = false;              // launch with programmatic dependent launch (set while the Lanczos graph is built)
    bool allow_pdl = true;         // KCRN_NO_PDL=1 switches it off
};

namespace kcrn {

// ----------------------------------------------------------------------------- kcrn_transpose.cu
int build_csc(const SegMat &csr, int64_t n_cols, SegMat *csc, cudaStream_t s, bool natural_order);
int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t len, cudaStream_t s);
int canonical_csc_to_host(const SegMat &csc, int32_t *indptr_host, int32_t *indices_host, double *data_host,
                          cudaStream_t s);

// ----------------------------------------------------------------------------- kcrn_sweep.cu
int build_tiles(SegMat *A, int64_t gather_len, cudaStream_t s, const Switches &sw);
void plan_tiles(const std::vector<int32_t> &ptr, int64_t nseg, int64_t nnz, std::vector<int4> &tiles,
                std::vector<int4> &longtab, int &npartial, int &items, const Switches &sw);   // host only
void free_segmat(SegMat *A);
int launch_forward(kcrn_problem *h, const double *x);                 // A x and everything derived
int launch_rows_weighted(kcrn_problem *h, const double *v);           // t = D .* (A v)
int launch_rows_lanczos(kcrn_problem *h, double *alpha_raw, int slot, int defer, int init_m = 0);   // init_m > 0 (slot 0): starts the recurrence itself          // same, source picked by ctl; sharded runs:
                                                                      // raw alpha sum -> alpha_raw[0]
int launch_cols_raw(kcrn_problem *h, const double *src, double *out); // out = A' src
int launch_cols_scaled(kcrn_problem *h, const double *src, const double *u, double *out, double *sq_part = nullptr);  // A' src / n + l2 u
int launch_cols_lanczos(kcrn_problem *h, int update_ctl, int slot, int defer);
bool lanczos_defer_fits(const kcrn_problem *h);      // small grids only: see kcrn_sweep.cu             // w' and beta fused
int launch_derive_from_Ax(kcrn_problem *h);
int tune_plain_sweeps(kcrn_problem *h);               // once per handle, at the end of kcrn_create
int gather_candidates_last();                            // last variant the tuners try

// ----------------------------------------------------------------------------- kcrn_lanczos.cu
int lanczos_alloc(kcrn_problem *h, int m);
int launch_lanczos_init(kcrn_problem *h, int m);      // ctl, alphas, bet reset; bet[0] = |g|
int launch_lanczos_post(kcrn_problem *h, const double *hv_or_raw, int raw);  // epilogue after allreduce / generic feed
int launch_lanczos_update(kcrn_problem *h, int reorth, int fused, int max_cols = 0);
int launch_lanczos_post_sharded(kcrn_problem *h, const double *y_allreduced, int reorth, int slot = -1);   // slot >= 0: deferred sums
int values_batch(kcrn_problem *h, const double *X, int k, double *out);
int launch_run_init(kcrn_problem *h);
int launch_run_begin(kcrn_problem *h, unsigned long long trial_loop);
int launch_run_decide(kcrn_problem *h, unsigned long long trial_loop);
int launch_run_commit(kcrn_problem *h, unsigned long long outer_loop);
int launch_lanczos_post_peer(kcrn_problem *h, const double *y_partial, int reorth, int slot = -1);
int launch_normalise_current(kcrn_problem *h);        // V[col] = src / bet[col] outside the fused sweep
int launch_sum(kcrn_problem *h, const double *a, int64_t len, double scale, double *out);  // deterministic sum
int launch_value_record(kcrn_problem *h, double *out, int with_record);   // value from the forward sweep's tile sums (+ record)
int launch_axpy_basis(const double *x, const double *V, const double *s, int m, const int32_t *info,
                      int64_t d, double *out, cudaStream_t st, double *partials = nullptr,
                      unsigned *ticket = nullptr, double *step_norm = nullptr);

// ----------------------------------------------------------------------------- kcrn_persist.cu
int persist_build(kcrn_problem *h);                   // end of kcrn_create: plans the tiles when the problem is small enough
int persist_alloc(kcrn_problem *h, int m);            // permuted basis for m columns
void persist_free(kcrn_problem *h);
int launch_persist_lanczos(kcrn_problem *h, int m);   // the whole m-step recurrence, one kernel
bool plan_persist_tiles(const std::vector<int32_t> &ptr, int64_t nseg, int64_t nnz, int grid, std::vector<PTile> &tiles,
                        std::vector<int4> &longtab, int &npartial);   // host only

// ----------------------------------------------------------------------------- kcrn_cubic.cu
int launch_cubic(const double *alphas, const double *betas, int m, const int32_t *info,
                 const double *gnorm_dev, double gnorm, double M, double r0, double eps, int it_max,
                 double *s_out, double *scal_out, cudaStream_t st, const double *params_dev);

int launch_cubic_dense(const double *H, const double *g, int m, double M, double r0, double eps, int it_max, double *out,
                       cudaStream_t st);

// ----------------------------------------------------------------------------- kcrn_nccl.cu
int nccl_unique_id(void *out128);
int nccl_init(kcrn_problem *h, const void *id128, int rank, int world);
int nccl_allreduce_sum(kcrn_problem *h, double *buf, int64_t count);
void nccl_destroy(kcrn_problem *h);

}  // namespace kcrn
