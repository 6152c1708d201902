// libcldrd.so — shard lifecycle, search orchestration and the CUDA entry points of the C ABI
// (include/cldrd.h).  Replaces, for CL-DRD's retrieval path, what the reference gets from faiss:
//   index_cpu_to_gpu(...)            retriever/retrieval_utils.py:159-163
//   index.search(x, k)               retriever/retrieval_utils.py:135,143
//   IndexShards merge                retriever/retrieval_utils.py:176-182
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common_host.h"
#include "scan_simt.cuh"
#include "scan_tc.cuh"
#include "scan_tc2.cuh"
#include "select.cuh"
#include "node.cuh"

using namespace cldrd;

namespace {

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(CLDRD_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

constexpr int kQueryBatch = 8192;   // queries processed per pass over the index
constexpr int kDensePiece = 8192;   // rows per dense piece (first chunk and fallback)
constexpr int kSurvCap = 8192;      // survivors one select CTA can take in per chunk (shared memory)
constexpr size_t kSurvTotal = size_t(kQueryBatch) * 32768;  // survivor-buffer entries (all queries), 2 GiB
constexpr int kMaxQStride = 65536;  // survivor slice per query when the batch is small
constexpr int kMaxGroups = 511;     // segments per query slice (+1 pool slot = select kernel's scan width)
constexpr size_t kSegCntInts = size_t(kQueryBatch) * (kMaxGroups + 1);
constexpr int64_t kSeedMinRows = 1 << 20;  // below this the progressive scheme is already cheap
constexpr int kMaxRunLen = 16;      // row tiles per work unit: short runs keep the CTAs in flight
                                    // inside a window of index rows that stays hot in L2
constexpr int kMaxRunLenByGroup = 64;  // ... stretched up to this if that makes #groups <= #CTAs: then
                                    // every unit owns a fresh survivor segment and the survivors
                                    // spread exactly evenly (with a few groups per CTA the
                                    // (query, CTA) cells would be loaded very unevenly)

// NVTX range for the timeline tools (Nsight Systems / ncu --nvtx); a no-op when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2D row-major [rows][d] tensor, box = [box_rows][128 bytes of K], 128B swizzle, zero OOB fill.
int make_tensor_map(CUtensorMap* tm, int scan, const void* base, int64_t rows, int d, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(CLDRD_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    CUtensorMapDataType dt;
    int esz;
    if (scan == CLDRD_SCAN_TC_TF32) {
        dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        esz = 4;
    } else if (scan == CLDRD_SCAN_TC_F16) {
        dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        esz = 2;
    } else {
        dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        esz = 2;
    }
    cuuint64_t gdim[2] = {cuuint64_t(d), cuuint64_t(rows)};
    cuuint64_t gstride[1] = {cuuint64_t(d) * esz};
    cuuint32_t box[2] = {cuuint32_t(TC_KB_BYTES / esz), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CLDRD_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return CLDRD_OK;
}

bool is_tc(int scan) { return scan == CLDRD_SCAN_TC_TF32 || scan == CLDRD_SCAN_TC_F16 || scan == CLDRD_SCAN_TC_BF16; }
int lp_kind_of(int scan) { return scan == CLDRD_SCAN_TC_F16 ? 1 : scan == CLDRD_SCAN_TC_BF16 ? 2 : 0; }

}  // namespace

struct cldrd_shard {
    int device = 0;
    int64_t row0 = 0;
    int64_t nrows = 0;
    int d = 0;
    int scan = 0;
    int scan_eff = 0;        // scan actually used (SIMT when TMA cannot describe the shape)
    bool finalized = false;

    float* xb = nullptr;     // fp32 rows (owned unless adopted)
    bool xb_owned = false;
    void* xlp = nullptr;     // fp16 / bf16 copy
    int64_t* ids = nullptr;  // external ids (device), optional
    float bmax_norm = 0.f, bmax_abs = 0.f;
    int vec4 = 0;
    int num_sms = 0;
    CUtensorMap tmB;
    CUtensorMap tmBh;        // half row tile (128 rows) for the 2-CTA scan
    bool use_tc2 = true;     // cta_group::2 scan kernel where it pays (CLDRD_TC2=0: always the 1-CTA kernel)

    // workspace for one query batch
    int ws_keep_cap = 0;
    size_t ws_d = 0;
    float* w_thr = nullptr;
    float* w_band = nullptr;
    float* w_seed = nullptr;      // seed threshold per query (-inf = unseeded)
    float* w_seed_in = nullptr;   // seed computed from this shard's own sample
    float* w_topj = nullptr;      // [Q][CLDRD_SEED_J] sample scores
    int* w_list_len = nullptr;
    int* w_seg_cnt = nullptr;
    unsigned long long* w_wait = nullptr;   // [4] issuer wait-cycle breakdown of the profiled scans
    int* w_unit_ctr = nullptr;   // [kUnitCtrSlots] work-unit counters, one per scan launch of a pass
    int* w_fail = nullptr;
    int* w_fail_index = nullptr;
    uint64_t* w_list = nullptr;
    uint64_t* w_surv = nullptr;
    float* w_dense = nullptr;
    void* w_qlp = nullptr;
    float* w_qfail = nullptr;     // compacted queries for the dense fallback
    unsigned long long* w_stats = nullptr;   // device [ST_COUNT]
    unsigned long long* h_stats = nullptr;   // pinned   [ST_COUNT]
    int* h_fail = nullptr;                   // pinned   [kQueryBatch]

    // pinned staging for cldrd_search_host
    float* h_q = nullptr;
    size_t h_q_bytes = 0;
    float* h_D = nullptr;
    int64_t* h_I = nullptr;
    size_t h_out_elems = 0;
    float* d_q = nullptr;
    size_t d_q_bytes = 0;
    float* d_D = nullptr;
    int64_t* d_I = nullptr;
    size_t d_out_elems = 0;

    int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    float* w_levels = nullptr;    // [kQueryBatch][CLDRD_SEED_J] levels of the node-wide search (best first)

    // scatter mode of the current re-score (node-wide search): lists go to the merging rank's key planes
    struct {
        int world = 0, rank = 0;
        int64_t slice = 0;
        int key_stride = 0;
        bool raise_fail = false;
        PeerPtrs keys;
        PeerPtrs len;
        PeerPtrs qfail;
    } sc;

    // optional per-kernel timing of the scan launches (bench roofline): event pairs on the
    // launching stream, summed after the search's final synchronisation
    bool profile = false;
    std::vector<cudaEvent_t> ev;   // 2 per scan launch
    size_t ev_used = 0;
    double scan_ms = 0.0;
    int64_t scan_launches = 0;
    std::vector<int64_t> ev_rows;      // rows of each timed scan launch
    std::vector<float> ev_ms;          // its device time

    // tuning overrides (environment: CLDRD_RUN_LEN, CLDRD_GROWTH), 0 = automatic
    int tune_run_len = 0;
    double tune_growth = 0.0;
    bool no_seed = false;   // CLDRD_NO_SEED=1: always use the progressive scheme
    bool pipeline = false;  // CLDRD_PIPELINE=1: re-score of the first half batch overlaps the scan of the second
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_half = nullptr, ev_aux_done = nullptr;
    int tune_seed_chunks = 0;    // CLDRD_SEED_CHUNKS: force the number of launches of a seeded pass
    float tune_seed_bias = 0.f;  // CLDRD_SEED_BIAS: added to every seed (tests force seed misses with it)
};

namespace {

void free_workspace(cldrd_shard* s) {
    cudaFree(s->w_thr);
    cudaFree(s->w_band);
    cudaFree(s->w_seed);
    cudaFree(s->w_seed_in);
    cudaFree(s->w_topj);
    s->w_seed = s->w_seed_in = s->w_topj = nullptr;
    cudaFree(s->w_list_len);
    cudaFree(s->w_seg_cnt);
    cudaFree(s->w_unit_ctr);
    cudaFree(s->w_wait);
    s->w_wait = nullptr;
    s->w_unit_ctr = nullptr;
    cudaFree(s->w_fail);
    cudaFree(s->w_fail_index);
    cudaFree(s->w_levels);
    s->w_levels = nullptr;
    cudaFree(s->w_list);
    cudaFree(s->w_surv);
    cudaFree(s->w_dense);
    cudaFree(s->w_qlp);
    cudaFree(s->w_qfail);
    cudaFree(s->w_stats);
    if (s->h_stats) cudaFreeHost(s->h_stats);
    if (s->h_fail) cudaFreeHost(s->h_fail);
    s->w_thr = s->w_band = nullptr;
    s->w_list_len = s->w_seg_cnt = s->w_fail = s->w_fail_index = nullptr;
    s->w_list = s->w_surv = nullptr;
    s->w_dense = nullptr;
    s->w_qlp = nullptr;
    s->w_qfail = nullptr;
    s->w_stats = nullptr;
    s->h_stats = nullptr;
    s->h_fail = nullptr;
    s->ws_keep_cap = 0;
}

int ensure_workspace(cldrd_shard* s) {
    if (s->w_thr) return CLDRD_OK;
    const int keep_cap = (s->scan_eff == CLDRD_SCAN_TC_BF16) ? 8192 : 4096;
    const size_t Q = kQueryBatch;
    CU_TRY(cudaMalloc(&s->w_thr, Q * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_band, Q * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_seed, Q * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_seed_in, Q * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_topj, Q * CLDRD_SEED_J * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_list_len, Q * sizeof(int)));
    CU_TRY(cudaMalloc(&s->w_seg_cnt, kSegCntInts * sizeof(int)));
    CU_TRY(cudaMalloc(&s->w_unit_ctr, sizeof(int)));
    CU_TRY(cudaMalloc(&s->w_wait, 4 * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(s->w_wait, 0, 4 * sizeof(unsigned long long)));
    CU_TRY(cudaMalloc(&s->w_fail, Q * sizeof(int)));
    CU_TRY(cudaMalloc(&s->w_fail_index, Q * sizeof(int)));
    CU_TRY(cudaMalloc(&s->w_levels, Q * CLDRD_SEED_J * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_list, Q * keep_cap * sizeof(uint64_t)));
    CU_TRY(cudaMalloc(&s->w_surv, kSurvTotal * sizeof(uint64_t)));
    CU_TRY(cudaMalloc(&s->w_dense, Q * kDensePiece * sizeof(float)));
    CU_TRY(cudaMalloc(&s->w_qfail, Q * size_t(s->d) * sizeof(float)));
    if (lp_kind_of(s->scan_eff)) CU_TRY(cudaMalloc(&s->w_qlp, Q * size_t(s->d) * 2));
    CU_TRY(cudaMalloc(&s->w_stats, ST_COUNT * sizeof(unsigned long long)));
    CU_TRY(cudaHostAlloc(&s->h_stats, ST_COUNT * sizeof(unsigned long long), cudaHostAllocDefault));
    CU_TRY(cudaHostAlloc(&s->h_fail, Q * sizeof(int), cudaHostAllocDefault));
    if (!s->aux_stream) {
        CU_TRY(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&s->ev_half, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s->ev_aux_done, cudaEventDisableTiming));
    }
    s->ws_keep_cap = keep_cap;
    return CLDRD_OK;
}

// CUDA loads a kernel's code on its first launch, and that load may synchronise the whole context.  The node-wide
// search keeps flag barriers spinning on one stream while another stream of the same process still has kernels to
// launch (several shards driven by one host thread): a first-time load behind a spinning barrier would wait for it
// forever.  So every kernel of the library is loaded up front, once per device.
template <typename F>
void preload_one(F* fn) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, fn) != cudaSuccess) cudaGetLastError();
}
void preload_kernels(int device) {
    static std::mutex mu;
    static std::map<int, bool> done;
    std::lock_guard<std::mutex> lk(mu);
    if (done[device]) return;
    done[device] = true;
#define PRELOAD_TC(KIND)                          \
    preload_one(scan_tc_kernel<KIND, TC_DENSE>);  \
    preload_one(scan_tc_kernel<KIND, TC_MAXES>);  \
    preload_one(scan_tc_kernel<KIND, TC_FILTER>); \
    preload_one(scan_tc2_kernel<KIND, TC_DENSE>); \
    preload_one(scan_tc2_kernel<KIND, TC_MAXES>); \
    preload_one(scan_tc2_kernel<KIND, TC_FILTER>)
    PRELOAD_TC(0);
    PRELOAD_TC(1);
    PRELOAD_TC(2);
#undef PRELOAD_TC
    preload_one(scan_simt_kernel<true, true>);
    preload_one(scan_simt_kernel<true, false>);
    preload_one(scan_simt_kernel<false, true>);
    preload_one(scan_simt_kernel<false, false>);
    preload_one(select_merge_kernel);
    preload_one(rescore_sort_kernel);
    preload_one(merge_kernel);
    preload_one(query_prep_kernel);
    preload_one(index_prep_kernel);
    preload_one(sample_topj_kernel);
    preload_one(seed_from_samples_kernel);
    preload_one(verify_seed_kernel);
    preload_one(apply_seed_kernel);
    preload_one(gather_failed_kernel);
    preload_one(node_barrier_kernel);
    preload_one(levels_seed_kernel);
    preload_one(count_levels_peers_kernel);
    preload_one(merge_keys_kernel);
    preload_one(node_tail_kernel);
    preload_one(node_publish_kernel);
    preload_one(node_spread_kernel);
}

// error-bound coefficients: eps = coef * |q| * max|b| + abs_coef * (|q| + max|b|)   
void eps_coefs(int scan, int d, float* coef, float* abs_coef) {
    const double u24 = std::ldexp(1.0, -24);
    const double rescore = (d + 16) * u24;          // fp32 re-score vs the real dot product
    const double accum = 2.2 * d * std::ldexp(1.0, -23);  // tensor-core accumulation slack
    double c = 0.0, a = 0.0;
    switch (scan) {
        case CLDRD_SCAN_SIMT_F32: c = (d + 16) * u24; break;
        case CLDRD_SCAN_TC_TF32: c = std::ldexp(1.0, -9) * 1.002 + accum; break;   // operands truncated to 10+1 bits
        case CLDRD_SCAN_TC_F16:
            c = std::ldexp(1.0, -10) * 1.002 + accum;
            a = std::ldexp(1.0, -24) * std::sqrt(double(d));                    // fp16 subnormal inputs
            break;
        case CLDRD_SCAN_TC_BF16: c = std::ldexp(1.0, -7) * 1.01 + accum; break;
    }
    *coef = float((c + rescore) * 1.01);
    *abs_coef = float(a);
}

struct BatchCtx {
    cldrd_shard* s;
    cudaStream_t st;
    const float* q;    // fp32 queries of this batch (device)
    int nq;
    int k;
    CUtensorMap tmA;
    bool have_tmA = false;
    int64_t launches = 0;
    int64_t chunks = 0;
    // survivor-buffer plan of the current chunk (scan writes it, select reads it)
    int q_stride = 0, seg_cap = 0, pool_cap = 0, groups = 1, run_len = 1, grid = 0, seg_by_group = 0;
    double seed_rank = 0.0;   // expected rank (in the whole index) of the seed threshold
    double expected_surv = 0.0;  // survivors per query this shard expects from a one-launch seeded pass (0 = unknown)
    bool tc2 = false;         // this launch uses the CTA-pair kernel
    // Workspace view.  A batch normally owns the whole workspace (qoff = 0, region = 0, 1 region);
    // a pipelined batch is cut in two halves that use disjoint query ranges of every per-query
    // array (qoff) and disjoint halves of the survivor buffer / segment counters (region).
    int qoff = 0;
    int region = 0, regions = 1;
};

inline uint64_t* ws_surv(const BatchCtx& c) { return c.s->w_surv + size_t(c.region) * (kSurvTotal / size_t(c.regions)); }
inline int* ws_seg_cnt(const BatchCtx& c) { return c.s->w_seg_cnt + size_t(c.region) * (kSegCntInts / size_t(c.regions)); }

// Work-unit plan for one chunk.  The grid is one persistent CTA per SM (fewer when there is less
// work); a unit is `run_len` consecutive row tiles of one query tile; the query's survivor slice
// is cut into one segment per CTA.
void plan_chunk(BatchCtx& c, int nrows) {
    cldrd_shard* s = c.s;
    const int nq_pad = ((c.nq + TC_BM - 1) / TC_BM) * TC_BM;
    c.q_stride = int(std::min<size_t>((kSurvTotal / size_t(c.regions)) / size_t(nq_pad), size_t(kMaxQStride)));
    if (!is_tc(s->scan_eff)) {
        c.groups = 1;
        c.run_len = 1;
        c.pool_cap = c.q_stride / 4;
        c.seg_cap = c.q_stride - c.pool_cap;
        c.grid = 0;
        c.seg_by_group = 0;
        return;
    }
    const int num_m = nq_pad / TC_BM;
    const int num_n = std::max(1, (nrows + TC_BN - 1) / TC_BN);
    long long tiles = (long long)num_m * num_n;
    // CTA pairs compute 256-query tiles: they pay when the query tiles pair up without much padding
    // (a lone 128-query tile would double the tensor work of the HBM-bound small-batch regime)
    c.tc2 = s->use_tc2 && (num_m >= 8 || (num_m >= 2 && num_m % 2 == 0));
    if (c.tc2) {
        // CTA pairs: a work unit covers two query tiles; the grid is an even number of CTAs
        tiles = (long long)((num_m + 1) / 2) * num_n;
        const int clusters = int(std::min<long long>(tiles, std::min(s->num_sms, kMaxGroups) / 2));
        c.grid = 2 * clusters;
        const long long per_cluster = tiles / clusters;
        c.run_len = int(std::max<long long>(1, std::min<long long>(kMaxRunLen, per_cluster / 16)));
        if (s->tune_run_len > 0 && per_cluster >= 16 * kMaxRunLen) c.run_len = s->tune_run_len;
        if ((num_n + c.run_len - 1) / c.run_len > clusters && (num_n + clusters - 1) / clusters <= kMaxRunLenByGroup)
            c.run_len = (num_n + clusters - 1) / clusters;
        const int num_groups2 = (num_n + c.run_len - 1) / c.run_len;
        c.seg_by_group = num_groups2 <= clusters ? 1 : 0;
        c.groups = c.seg_by_group ? num_groups2 : clusters;   // one survivor segment per CTA pair
        c.pool_cap = c.q_stride / 4;
        c.seg_cap = std::max(1, (c.q_stride - c.pool_cap) / c.groups);
        return;
    }
    c.grid = int(std::min<long long>(tiles, std::min(s->num_sms, kMaxGroups)));
    const long long per_cta = tiles / c.grid;
    c.run_len = int(std::max<long long>(1, std::min<long long>(kMaxRunLen, per_cta / 16)));
    if (s->tune_run_len > 0 && per_cta >= 16 * kMaxRunLen) c.run_len = s->tune_run_len;
    if ((num_n + c.run_len - 1) / c.run_len > c.grid && (num_n + c.grid - 1) / c.grid <= kMaxRunLenByGroup)
        c.run_len = (num_n + c.grid - 1) / c.grid;
    const int num_groups = (num_n + c.run_len - 1) / c.run_len;
    c.seg_by_group = num_groups <= c.grid ? 1 : 0;
    c.groups = c.seg_by_group ? num_groups : c.grid;   // survivor segments per query
    c.pool_cap = c.q_stride / 4;
    c.seg_cap = std::max(1, (c.q_stride - c.pool_cap) / c.groups);
}

cudaEvent_t next_event(cldrd_shard* s) {
    if (s->ev_used == s->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s->ev.push_back(e);
    }
    return s->ev[s->ev_used++];
}

int launch_scan(BatchCtx& c, int mode, int64_t row_begin, int nrows, int tile_stride = 1) {
    NvtxRange nvtx(mode == TC_FILTER ? "cldrd::scan(filter)" : mode == TC_MAXES ? "cldrd::scan(sample)" : "cldrd::scan(dense)");
    const bool dense = mode != TC_FILTER;
    cldrd_shard* s = c.s;
    if (s->profile) {
        cudaEventRecord(next_event(s), c.st);
        s->ev_rows.push_back(nrows);
    }
    ScanParams p{};
    p.xb = s->xb;
    p.q = c.q;
    p.d = s->d;
    p.nq = c.nq;
    p.row_begin = row_begin;
    p.nrows = nrows;
    p.tile_stride = tile_stride;
    plan_chunk(c, nrows);
    p.thr = s->w_thr + c.qoff;
    p.surv = ws_surv(c);
    p.seg_cnt = ws_seg_cnt(c);
    p.q_stride = c.q_stride;
    p.seg_cap = c.seg_cap;
    p.pool_cap = c.pool_cap;
    p.groups = c.groups;
    p.run_len = c.run_len;
    p.seg_by_group = c.seg_by_group;
    p.unit_ctr = s->w_unit_ctr;
    p.wait_cycles = (s->profile && mode == TC_FILTER) ? s->w_wait : nullptr;
    p.dense = s->w_dense + size_t(c.qoff) * kDensePiece;
    p.dense_ld = kDensePiece;
    p.stats = s->w_stats;
    if (is_tc(s->scan_eff)) {
        const int esz = (s->scan_eff == CLDRD_SCAN_TC_TF32) ? 4 : 2;
        p.kb_elems = TC_KB_BYTES / esz;
        p.num_kb = (s->d * esz + TC_KB_BYTES - 1) / TC_KB_BYTES;
        const int grid = c.grid;
        if (grid <= 0) {
            if (s->profile) cudaEventRecord(next_event(s), c.st);
            return CLDRD_OK;
        }
        CU_TRY(cudaMemsetAsync(s->w_unit_ctr, 0, sizeof(int), c.st));
        if (mode == TC_FILTER)   // survivor cursors of this launch's segment layout start at zero
            CU_TRY(cudaMemsetAsync(ws_seg_cnt(c), 0, size_t(c.nq) * (c.groups + 1) * sizeof(int), c.st));
#define LAUNCH_TC(KIND)                                                                                   \
    do {                                                                                                  \
        if (mode == TC_DENSE)                                                                             \
            scan_tc_kernel<KIND, TC_DENSE><<<grid, TC_THREADS, TC_SMEM_BYTES, c.st>>>(c.tmA, s->tmB, p);  \
        else if (mode == TC_MAXES)                                                                        \
            scan_tc_kernel<KIND, TC_MAXES><<<grid, TC_THREADS, TC_SMEM_BYTES, c.st>>>(c.tmA, s->tmB, p);  \
        else                                                                                              \
            scan_tc_kernel<KIND, TC_FILTER><<<grid, TC_THREADS, TC_SMEM_BYTES, c.st>>>(c.tmA, s->tmB, p); \
    } while (0)
#define LAUNCH_TC2(KIND)                                                                                      \
    do {                                                                                                      \
        if (mode == TC_DENSE)                                                                                 \
            scan_tc2_kernel<KIND, TC_DENSE><<<grid, TC_THREADS, TC2_SMEM_BYTES, c.st>>>(c.tmA, s->tmBh, p);   \
        else if (mode == TC_MAXES)                                                                            \
            scan_tc2_kernel<KIND, TC_MAXES><<<grid, TC_THREADS, TC2_SMEM_BYTES, c.st>>>(c.tmA, s->tmBh, p);   \
        else                                                                                                  \
            scan_tc2_kernel<KIND, TC_FILTER><<<grid, TC_THREADS, TC2_SMEM_BYTES, c.st>>>(c.tmA, s->tmBh, p);  \
    } while (0)
        if (c.tc2) {
            if (s->scan_eff == CLDRD_SCAN_TC_F16) LAUNCH_TC2(0);
            else if (s->scan_eff == CLDRD_SCAN_TC_BF16) LAUNCH_TC2(1);
            else LAUNCH_TC2(2);
        } else if (s->scan_eff == CLDRD_SCAN_TC_F16) LAUNCH_TC(0);
        else if (s->scan_eff == CLDRD_SCAN_TC_BF16) LAUNCH_TC(1);
        else LAUNCH_TC(2);
#undef LAUNCH_TC
#undef LAUNCH_TC2
    } else {
        dim3 grid((nrows + SIMT_BN - 1) / SIMT_BN, (c.nq + SIMT_BM - 1) / SIMT_BM);
        if (grid.x == 0 || grid.y == 0) {
            if (s->profile) cudaEventRecord(next_event(s), c.st);
            return CLDRD_OK;
        }
        const bool vec = s->vec4 && (reinterpret_cast<uintptr_t>(c.q) % 16 == 0);
        if (!dense) CU_TRY(cudaMemsetAsync(ws_seg_cnt(c), 0, size_t(c.nq) * (c.groups + 1) * sizeof(int), c.st));
        if (dense) {
            if (vec) scan_simt_kernel<true, true><<<grid, 256, 0, c.st>>>(p);
            else scan_simt_kernel<true, false><<<grid, 256, 0, c.st>>>(p);
        } else {
            if (vec) scan_simt_kernel<false, true><<<grid, 256, 0, c.st>>>(p);
            else scan_simt_kernel<false, false><<<grid, 256, 0, c.st>>>(p);
        }
    }
    CU_TRY(cudaGetLastError());
    if (s->profile) cudaEventRecord(next_event(s), c.st);
    c.launches++;
    c.chunks++;
    return CLDRD_OK;
}

size_t select_smem(const cldrd_shard* s) { return size_t(s->ws_keep_cap + kSurvCap) * 8 + size_t(s->d) * 4 + 16; }

// k_sel: how many best rows the list must keep (the search's k, or the sample's j)
int launch_select(BatchCtx& c, bool dense, int64_t row_begin, int nrows, int k_sel, int tile_stride = 1,
                  bool one_launch_pass = false) {
    cldrd_shard* s = c.s;
    // Shared memory decides how many select CTAs share an SM (2 at the full 8192-survivor capacity), and the kernel is a
    // chain of barrier-separated passes: latency, not work.  A one-launch seeded pass knows how many survivors to expect
    // (the seed's rank, split over the shards), so it takes 5x that (at least 2048) and lets 4 CTAs share the SM; a query
    // that still overflows is flagged and searched again like any other overflow.
    int surv_cap = kSurvCap;
    if (!dense && one_launch_pass && c.expected_surv > 0.0)
        surv_cap = int(std::min<double>(kSurvCap, std::max(2048.0, 5.0 * c.expected_surv)));
    SelectParams p{};
    p.list = s->w_list + size_t(c.qoff) * s->ws_keep_cap;
    p.list_len = s->w_list_len + c.qoff;
    p.keep_cap = s->ws_keep_cap;
    p.surv = ws_surv(c);
    p.seg_cnt = ws_seg_cnt(c);
    p.q_stride = c.q_stride;
    p.seg_cap = c.seg_cap;
    p.pool_cap = c.pool_cap;
    p.groups = c.groups;
    p.surv_cap = surv_cap;
    p.dense = dense ? s->w_dense + size_t(c.qoff) * kDensePiece : nullptr;
    p.dense_ld = kDensePiece;
    p.dense_n = nrows;
    p.dense_row0 = uint32_t(row_begin);
    p.dense_tile_stride = tile_stride;
    p.thr = s->w_thr + c.qoff;
    p.seed = s->w_seed + c.qoff;
    p.band = s->w_band + c.qoff;
    p.k = k_sel;
    p.fail = s->w_fail + c.qoff;
    p.stats = s->w_stats;
    p.xb = s->xb;
    p.q = c.q;
    p.d = s->d;
    p.vec4 = s->vec4 && (reinterpret_cast<uintptr_t>(c.q) % 16 == 0);
    select_merge_kernel<<<c.nq, 512, size_t(s->ws_keep_cap + surv_cap) * 8 + size_t(s->d) * 4 + 16, c.st>>>(p);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return CLDRD_OK;
}

int launch_prep(BatchCtx& c) {
    cldrd_shard* s = c.s;
    QueryPrepParams p{};
    p.q = c.q;
    p.nq = c.nq;
    p.d = s->d;
    p.lp_kind = lp_kind_of(s->scan_eff);
    char* qlp = s->w_qlp ? static_cast<char*>(s->w_qlp) + size_t(c.qoff) * s->d * 2 : nullptr;
    p.q_lp = qlp;
    eps_coefs(s->scan_eff, s->d, &p.coef, &p.abs_coef);
    p.bmax_norm = s->bmax_norm;
    p.band = s->w_band + c.qoff;
    p.seed = s->w_seed + c.qoff;
    p.thr = s->w_thr + c.qoff;
    p.list_len = s->w_list_len + c.qoff;
    p.fail = s->w_fail + c.qoff;
    p.stats = s->w_stats;
    const int threads = 256;
    const int blocks = (c.nq * 32 + threads - 1) / threads;
    query_prep_kernel<<<blocks, threads, 0, c.st>>>(p);
    CU_TRY(cudaGetLastError());
    c.launches++;
    if (is_tc(s->scan_eff)) {
        const void* base = lp_kind_of(s->scan_eff) ? static_cast<const void*>(qlp) : static_cast<const void*>(c.q);
        int rc = make_tensor_map(&c.tmA, s->scan_eff, base, c.nq, s->d, TC_BM);
        if (rc) return rc;
        c.have_tmA = true;
    }
    return CLDRD_OK;
}

// out_scores / out_ids: the batch's output base (this view's qoff is applied here, unless the
// rows are scattered through out_index).  n_pad_small > 0: launch with shared memory for lists of
// at most that many entries (so that the CTAs fit next to a running scan CTA); longer lists are
// flagged as failed and redone by the fallback.
struct CountedCut {
    const int* planes = nullptr;   // [parts][stride] ints: counts the shards stored into this rank's block
    size_t stride = 0;
    int parts = 0;
    const float* levels = nullptr;
};

int launch_rescore(BatchCtx& c, float* out_scores, int64_t* out_ids, bool translate, const int* out_index,
                   int* fail_flags, cudaStream_t st, int n_pad_small = 0, const CountedCut* cc = nullptr) {
    cldrd_shard* s = c.s;
    RescoreParams p{};
    p.xb = s->xb;
    p.q = c.q;
    p.d = s->d;
    p.vec4 = s->vec4 && (reinterpret_cast<uintptr_t>(c.q) % 16 == 0);
    p.list = s->w_list + size_t(c.qoff) * s->ws_keep_cap;
    p.list_len = s->w_list_len + c.qoff;
    p.keep_cap = s->ws_keep_cap;
    p.n_pad = n_pad_small > 0 ? std::min(n_pad_small, s->ws_keep_cap) : s->ws_keep_cap;
    p.k = c.k;
    p.row0 = s->row0;
    p.ids = (translate && s->ids) ? s->ids : nullptr;
    const size_t ooff = out_index ? 0 : size_t(c.qoff) * c.k;
    p.out_scores = out_scores ? out_scores + ooff : nullptr;   // NULL in scatter mode
    p.out_ids = out_ids ? out_ids + ooff : nullptr;
    p.out_index = out_index;
    p.fail = fail_flags ? fail_flags + c.qoff : nullptr;
    p.fail_set = fail_flags ? fail_flags + c.qoff : nullptr;
    p.stats = s->w_stats;
    p.cut = nullptr;
    if (cc) {
        p.cnt_planes = cc->planes;
        p.cnt_plane_stride = cc->stride;
        p.cnt_parts = cc->parts;
        p.levels = cc->levels;
        p.lv_j = CLDRD_SEED_J;
        p.band = s->w_band + c.qoff;
    }
    p.sc_world = s->sc.world;
    if (s->sc.world > 0) {
        p.sc_rank = s->sc.rank;
        p.sc_slice = s->sc.slice;
        p.q_base = out_index ? 0 : c.qoff;
        p.sc_key_stride = s->sc.key_stride;
        p.sc_raise_fail = s->sc.raise_fail ? 1 : 0;
        p.sc_keys = s->sc.keys;
        p.sc_len = s->sc.len;
        p.sc_qfail = s->sc.qfail;
    }
    const size_t smem = size_t(p.n_pad) * 8 + size_t(s->d) * 4 + 16;
    // few queries (the reference's batch=128 loop): one CTA per query leaves most SMs with a single
    // CTA, so give it twice the warps to keep the row gather's loads in flight; the per-row result
    // does not depend on which warp computes it
    const int threads = c.nq <= 2 * s->num_sms ? 1024 : 512;
    rescore_sort_kernel<<<c.nq, threads, smem, st>>>(p);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return CLDRD_OK;
}

int read_stats(cldrd_shard* s, cudaStream_t st) {
    CU_TRY(cudaMemcpyAsync(s->h_stats, s->w_stats, ST_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (s->h_stats[ST_KERNEL_ERR])
        return fail(CLDRD_ECUDA, "scan kernel watchdog fired (code %llu)", s->h_stats[ST_KERNEL_ERR]);
    return CLDRD_OK;
}

// ---- seeded thresholds ---------------------------------------------------------

// Sample geometry of a shard: every `stride`-th 256-row tile, `tiles` tiles in all.  The sample
// fraction f = min(J / (3k), 1/64) makes the global J-th best sample score sit near rank J / f
// (about 3k) of the whole index.
struct SamplePlan {
    int tiles = 0;
    int stride = 1;
};
SamplePlan sample_plan(const cldrd_shard* s, int k) {
    SamplePlan sp;
    const int64_t full_tiles = s->nrows / TC_BN;   // only whole tiles are sampled
    if (full_tiles <= 0) return sp;
    // target rank of the seed: comfortably beyond k plus the rows inside the error band, which is
    // wider for the coarser scan formats
    const double rank_factor = s->scan_eff == CLDRD_SCAN_TC_BF16 ? 8.0 : s->scan_eff == CLDRD_SCAN_TC_TF32 ? 4.5 : 3.5;
    const double f = std::min(double(CLDRD_SEED_J) / (rank_factor * k), 1.0 / 64.0);
    int64_t tiles = int64_t(std::ceil(f * double(s->nrows) / TC_BN));
    // one launch: the tensor-core scan keeps 8 group maxima per tile in the dense buffer, the
    // SIMT scan keeps raw scores (256 per tile)
    const int64_t cap = is_tc(s->scan_eff) ? kDensePiece / (TC_BN / 32) : kDensePiece / TC_BN;
    tiles = std::max<int64_t>(1, std::min<int64_t>(tiles, std::min<int64_t>(full_tiles, cap)));
    sp.tiles = int(tiles);
    sp.stride = int(std::max<int64_t>(1, full_tiles / tiles));
    return sp;
}

// Scan the shard's sample in one launch and leave each query's best CLDRD_SEED_J sample values
// (best first, -inf padded) at element offset out_off of every buffer in outs, [nq][CLDRD_SEED_J].  Uses the batch
// workspace.
int run_sample(BatchCtx& c, const PeerPtrs& outs, int nouts, size_t out_off) {
    cldrd_shard* s = c.s;
    const SamplePlan sp = sample_plan(s, c.k);
    int cols = 0;
    if (sp.tiles > 0) {
        const bool tc = is_tc(s->scan_eff);
        int rc = launch_scan(c, tc ? TC_MAXES : TC_DENSE, 0, sp.tiles * TC_BN, sp.stride);
        if (rc) return rc;
        cols = tc ? sp.tiles * (TC_BN / 32) : sp.tiles * TC_BN;
    }
    static_assert(CLDRD_SEED_J <= 64, "sample_topj_kernel sorts at most 64 values");
    sample_topj_kernel<<<c.nq, 256, size_t(std::max(cols, 1)) * 8, c.st>>>(s->w_dense, kDensePiece, cols, CLDRD_SEED_J, outs, nouts,
                                                                           out_off);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return CLDRD_OK;
}

int run_sample(BatchCtx& c, float* out_topj) {
    PeerPtrs one{};
    one.p[0] = out_topj;
    return run_sample(c, one, 1, 0);
}

int apply_seed(BatchCtx& c, const float* seed_in) {
    cldrd_shard* s = c.s;
    apply_seed_kernel<<<(c.nq + 255) / 256, 256, 0, c.st>>>(seed_in, c.nq, s->tune_seed_bias, s->w_seed, s->w_thr, s->w_list_len);
    CU_TRY(cudaGetLastError());
    c.launches++;
    return CLDRD_OK;
}

// One pass of `c.nq` (<= kQueryBatch) queries over the whole shard.
//   PASS_SEEDED      thresholds start at the seed already applied to the workspace: a few big chunks
//   PASS_PROGRESSIVE first piece dense, then geometrically growing filtered chunks
//   PASS_DENSE       every piece dense: cannot overflow (last-resort fallback)
enum PassKind { PASS_SEEDED, PASS_PROGRESSIVE, PASS_DENSE };

int run_chunks(BatchCtx& c, PassKind kind) {
    cldrd_shard* s = c.s;
    const int64_t N = s->nrows;
    int rc;
    if (kind == PASS_SEEDED) {
        // The seed already sits near the final threshold, so the whole shard is ONE launch: no
        // inter-chunk dependency, no select between chunks, one tail.  (Chunks only come back for
        // shards beyond the int32 row range of a launch.)
        const int64_t target = (int64_t(1) << 31) - TC_BN;
        int nchunks = int(std::max<int64_t>(1, (N + target - 1) / target));
        // ... or when the seed's expected rank exceeds what a select CTA takes in per launch
        // (large k, or the small sample of the SIMT scan): then thresholds tighten between chunks
        nchunks = std::max(nchunks, int(std::ceil(c.seed_rank / (0.5 * kSurvCap))));
        if (s->tune_seed_chunks > 0) nchunks = s->tune_seed_chunks;
        int64_t done = 0;
        for (int i = 0; i < nchunks; ++i) {
            int64_t m = (N * (i + 1)) / nchunks - done;
            if (i + 1 < nchunks) m = (m / TC_BN) * TC_BN;
            if (m <= 0) continue;
            if ((rc = launch_scan(c, TC_FILTER, done, int(m)))) return rc;
            if ((rc = launch_select(c, false, done, int(m), c.k, 1, nchunks == 1))) return rc;
            done += m;
        }
        return CLDRD_OK;
    }
    int64_t done = 0;
    const int first = int(std::min<int64_t>(N, kDensePiece));
    if (first > 0) {
        if ((rc = launch_scan(c, TC_DENSE, 0, first))) return rc;
        if ((rc = launch_select(c, true, 0, first, c.k))) return rc;
        done = first;
    }
    if (done >= N) return CLDRD_OK;
    if (kind == PASS_DENSE) {
        while (done < N) {
            const int m = int(std::min<int64_t>(N - done, kDensePiece));
            if ((rc = launch_scan(c, TC_DENSE, done, m))) return rc;
            if ((rc = launch_select(c, true, done, m, c.k))) return rc;
            done += m;
        }
        return CLDRD_OK;
    }
    // Size the growing chunks from the candidate-list length after the first piece: a chunk of
    // m rows after `done` rows is expected to push about len * m / done survivors per query;
    // keep that at half of what a select CTA can take in (the counts are sums of many
    // near-independent hits, so the spread around the mean is small; order-correlated data
    // that breaks this goes to the dense fallback).
    if ((rc = read_stats(s, c.st))) return rc;
    const double k_eff = std::max<double>(c.k, double(s->h_stats[ST_MAX_LIST])) * 1.1;
    double growth = double(kSurvCap) / (2.0 * k_eff);
    if (s->tune_growth > 0.0) growth = s->tune_growth;
    while (done < N) {
        int64_t m = int64_t(double(done) * growth);
        m = std::max<int64_t>(m, TC_BN);
        m = (m / TC_BN) * TC_BN;
        m = std::min<int64_t>(m, int64_t(1) << 30);
        if (m > N - done) m = N - done;
        if ((rc = launch_scan(c, TC_FILTER, done, int(m)))) return rc;
        if ((rc = launch_select(c, false, done, int(m), c.k))) return rc;
        done += m;
    }
    return CLDRD_OK;
}

struct SearchTotals {
    int64_t launches = 0, chunks = 0, fallback_queries = 0, seed_misses = 0;
};

// Queries flagged in the workspace's fail[] are searched again with a safer pass kind, results
// scattered into their rows of the output.  Level 1: unseeded progressive.  Level 2: dense.
int run_fallbacks(cldrd_shard* s, const float* q_dev, int nq, int k, bool translate, float* out_scores,
                  int64_t* out_ids, cudaStream_t st, bool allow_progressive, SearchTotals* tot) {
    std::vector<int> pending;
    CU_TRY(cudaMemcpyAsync(s->h_fail, s->w_fail, size_t(nq) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < nq; ++i)
        if (s->h_fail[i]) pending.push_back(i);
    tot->fallback_queries += int64_t(pending.size());
    for (int level = allow_progressive ? 1 : 2; level <= 2 && !pending.empty(); ++level) {
        const int nf = int(pending.size());
        CU_TRY(cudaMemcpyAsync(s->w_fail_index, pending.data(), size_t(nf) * sizeof(int), cudaMemcpyHostToDevice, st));
        gather_failed_kernel<<<nf, 128, 0, st>>>(q_dev, s->d, s->w_fail_index, nf, s->w_qfail);
        CU_TRY(cudaGetLastError());
        BatchCtx f{};
        f.s = s;
        f.st = st;
        f.q = s->w_qfail;
        f.nq = nf;
        f.k = k;
        int rc = launch_prep(f);
        if (rc) return rc;
        if ((rc = run_chunks(f, level == 1 ? PASS_PROGRESSIVE : PASS_DENSE))) return rc;
        if ((rc = launch_rescore(f, out_scores, out_ids, translate, s->w_fail_index, s->w_fail, st))) return rc;
        if ((rc = read_stats(s, st))) return rc;
        tot->launches += f.launches + 1;
        tot->chunks += f.chunks;
        CU_TRY(cudaMemcpyAsync(s->h_fail, s->w_fail, size_t(nf) * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        std::vector<int> next;
        for (int i = 0; i < nf; ++i)
            if (s->h_fail[i]) next.push_back(pending[i]);
        pending.swap(next);
    }
    if (!pending.empty()) return fail(CLDRD_ECUDA, "internal: %zu queries failed the dense fallback", pending.size());
    return CLDRD_OK;
}

// Search one batch.  seed_mode: 0 = none (progressive), 1 = automatic (sample this shard, verify
// here: single-shard search), 2 = external seed_ext (no verification here: the caller verifies on
// the merged result of all shards).
int search_batch(cldrd_shard* s, const float* q_dev, int nq, int k, bool translate, int seed_mode,
                 const float* seed_ext, float* out_scores, int64_t* out_ids, float* eps_out, cudaStream_t st,
                 SearchTotals* tot) {
    NvtxRange nvtx("cldrd::search_batch");
    BatchCtx c{};
    c.s = s;
    c.st = st;
    c.q = q_dev;
    c.nq = nq;
    c.k = k;
    int rc = launch_prep(c);
    if (rc) return rc;
    if (seed_mode) {
        // expected rank of the seed = J / sample fraction (all shards sample the same fraction)
        const SamplePlan sp = sample_plan(s, k);
        const double frac = sp.tiles > 0 ? double(sp.tiles) * TC_BN / double(std::max<int64_t>(s->nrows, 1)) : 1.0;
        c.seed_rank = double(CLDRD_SEED_J) / frac;
        if (seed_mode == 1) c.expected_surv = c.seed_rank;   // own sample: the seed's rank in this shard
    }
    if (seed_mode == 1) {
        if ((rc = run_sample(c, s->w_topj))) return rc;
        seed_from_samples_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(s->w_topj, 1, nq, CLDRD_SEED_J, CLDRD_SEED_J,
                                                                        s->w_seed_in);
        CU_TRY(cudaGetLastError());
        c.launches++;
        if ((rc = apply_seed(c, s->w_seed_in))) return rc;
    } else if (seed_mode == 2) {
        if ((rc = apply_seed(c, seed_ext))) return rc;
    }
    // Pipelined halves (seeded tensor-core scans of big batches): the re-score of the first half is
    // HBM-bound, the scan of the second half is tensor-bound, so they run concurrently -- the
    // re-score on an auxiliary stream with a shared-memory footprint small enough for its CTAs to
    // sit next to the persistent scan CTAs.  The halves use disjoint query ranges of every
    // per-query array and disjoint halves of the survivor buffer.
    const int half = ((nq / 2 + 2 * TC_BM - 1) / (2 * TC_BM)) * (2 * TC_BM);
    const bool pipelined = s->pipeline && seed_mode != 0 && is_tc(s->scan_eff) && nq >= 8 * TC_BM && half < nq &&
                           s->aux_stream != nullptr;
    if (!pipelined) {
        if ((rc = run_chunks(c, seed_mode ? PASS_SEEDED : PASS_PROGRESSIVE))) return rc;
        if ((rc = launch_rescore(c, out_scores, out_ids, translate, nullptr, s->w_fail, st))) return rc;
        if (seed_mode == 1) {
            verify_seed_kernel<<<(nq + 255) / 256, 256, 0, st>>>(out_scores, nq, k, s->w_seed, s->w_band, s->w_fail, s->w_stats);
            CU_TRY(cudaGetLastError());
            c.launches++;
        }
    } else {
        BatchCtx h[2] = {c, c};
        h[0].nq = half;
        h[1].nq = nq - half;
        h[1].q = q_dev + size_t(half) * s->d;
        h[1].qoff = half;
        for (int i = 0; i < 2; ++i) {
            h[i].region = i;
            h[i].regions = 2;
            h[i].launches = h[i].chunks = 0;
            const void* base = lp_kind_of(s->scan_eff)
                                   ? static_cast<const void*>(static_cast<char*>(s->w_qlp) + size_t(h[i].qoff) * s->d * 2)
                                   : static_cast<const void*>(h[i].q);
            if ((rc = make_tensor_map(&h[i].tmA, s->scan_eff, base, h[i].nq, s->d, TC_BM))) return rc;
        }
        auto verify = [&](BatchCtx& x, cudaStream_t vs) -> int {
            if (seed_mode != 1) return CLDRD_OK;
            verify_seed_kernel<<<(x.nq + 255) / 256, 256, 0, vs>>>(out_scores + size_t(x.qoff) * k, x.nq, k, s->w_seed + x.qoff,
                                                                   s->w_band + x.qoff, s->w_fail + x.qoff, s->w_stats);
            CU_TRY(cudaGetLastError());
            x.launches++;
            return CLDRD_OK;
        };
        // half 0: scan + select on the main stream, re-score (+verify) on the auxiliary stream
        if ((rc = run_chunks(h[0], PASS_SEEDED))) return rc;
        CU_TRY(cudaEventRecord(s->ev_half, st));
        CU_TRY(cudaStreamWaitEvent(s->aux_stream, s->ev_half, 0));
        if ((rc = launch_rescore(h[0], out_scores, out_ids, translate, nullptr, s->w_fail, s->aux_stream, 2048))) return rc;
        if ((rc = verify(h[0], s->aux_stream))) return rc;
        CU_TRY(cudaEventRecord(s->ev_aux_done, s->aux_stream));
        // half 1: everything on the main stream, concurrent with half 0's re-score
        if ((rc = run_chunks(h[1], PASS_SEEDED))) return rc;
        if ((rc = launch_rescore(h[1], out_scores, out_ids, translate, nullptr, s->w_fail, st))) return rc;
        if ((rc = verify(h[1], st))) return rc;
        CU_TRY(cudaStreamWaitEvent(st, s->ev_aux_done, 0));
        c.launches += h[0].launches + h[1].launches;
        c.chunks += h[0].chunks + h[1].chunks;
    }
    if (eps_out) {   // eps = band / 2, for the caller's verification of an external seed
        CU_TRY(cudaMemcpyAsync(eps_out, s->w_band, size_t(nq) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // any query to redo?  (one small D2H; also surfaces watchdog / range errors)
    if ((rc = read_stats(s, st))) return rc;
    tot->launches += c.launches;
    tot->chunks += c.chunks;
    if (s->h_stats[ST_RANGE_ERR])
        return fail(CLDRD_EINVAL, "query values exceed the fp16 range; use the bf16 or tf32 scan");
    if (s->h_stats[ST_FAILED] == 0) return CLDRD_OK;
    // seeded: first retry unseeded-progressive; unseeded: straight to dense
    return run_fallbacks(s, q_dev, nq, k, translate, out_scores, out_ids, st, seed_mode != 0, tot);
}

}  // namespace

extern "C" {

int cldrd_shard_create(cldrd_shard** out, int device, int64_t row0, int64_t nrows, int32_t d, int32_t scan) {
    if (!out || nrows < 0 || d <= 0 || row0 < 0 || scan < 0 || scan > 3)
        return fail(CLDRD_EINVAL, "shard_create: bad argument (nrows=%lld d=%d scan=%d)", (long long)nrows, d, scan);
    if (row0 + nrows >= (int64_t(1) << 32) - 1)
        return fail(CLDRD_EINVAL, "shard_create: rows beyond 2^32-2 are not addressable by the candidate keys");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(CLDRD_ECUDA, "no CUDA device: libcldrd has no CPU search path");
    if (device < 0 || device >= ndev) return fail(CLDRD_EINVAL, "shard_create: device %d of %d", device, ndev);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CLDRD_ECUDA, "device %d is sm_%d%d; libcldrd is built for sm_100a only", device, prop.major, prop.minor);
    auto* s = new cldrd_shard();
    s->device = device;
    s->row0 = row0;
    s->nrows = nrows;
    s->d = d;
    s->scan = scan;
    s->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("CLDRD_RUN_LEN")) s->tune_run_len = atoi(e);
    if (const char* e = getenv("CLDRD_GROWTH")) s->tune_growth = atof(e);
    if (const char* e = getenv("CLDRD_NO_SEED")) s->no_seed = atoi(e) != 0;
    if (const char* e = getenv("CLDRD_SEED_BIAS")) s->tune_seed_bias = float(atof(e));
    if (const char* e = getenv("CLDRD_SEED_CHUNKS")) s->tune_seed_chunks = atoi(e);
    if (const char* e = getenv("CLDRD_TC2")) s->use_tc2 = atoi(e) != 0;
    if (const char* e = getenv("CLDRD_PIPELINE")) s->pipeline = atoi(e) != 0;
    *out = s;
    return CLDRD_OK;
}

void cldrd_shard_destroy(cldrd_shard* s) {
    if (!s) return;
    DeviceGuard g(s->device);
    free_workspace(s);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
    if (s->ev_half) cudaEventDestroy(s->ev_half);
    if (s->ev_aux_done) cudaEventDestroy(s->ev_aux_done);
    if (s->xb_owned) cudaFree(s->xb);
    cudaFree(s->xlp);
    cudaFree(s->ids);
    if (s->h_q) cudaFreeHost(s->h_q);
    if (s->h_D) cudaFreeHost(s->h_D);
    if (s->h_I) cudaFreeHost(s->h_I);
    cudaFree(s->d_q);
    cudaFree(s->d_D);
    cudaFree(s->d_I);
    delete s;
}

static int ensure_rows(cldrd_shard* s) {
    if (s->xb) return CLDRD_OK;
    CU_TRY(cudaMalloc(&s->xb, std::max<size_t>(size_t(s->nrows) * s->d * sizeof(float), 16)));
    s->xb_owned = true;
    return CLDRD_OK;
}

int cldrd_shard_upload(cldrd_shard* s, const float* rows_host, int64_t row_off, int64_t n) {
    if (!s || (!rows_host && n) || row_off < 0 || n < 0 || row_off + n > s->nrows)
        return fail(CLDRD_EINVAL, "shard_upload: bad range");
    if (s->xb && !s->xb_owned) return fail(CLDRD_ESTATE, "shard_upload: rows were adopted");
    DeviceGuard g(s->device);
    int rc = ensure_rows(s);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(s->xb + size_t(row_off) * s->d, rows_host, size_t(n) * s->d * sizeof(float), cudaMemcpyHostToDevice));
    s->finalized = false;
    return CLDRD_OK;
}

int cldrd_shard_adopt(cldrd_shard* s, const float* rows_dev) {
    if (!s || (!rows_dev && s->nrows)) return fail(CLDRD_EINVAL, "shard_adopt: NULL");
    if (s->xb && s->xb_owned) return fail(CLDRD_ESTATE, "shard_adopt: rows already uploaded");
    if (reinterpret_cast<uintptr_t>(rows_dev) % 16) return fail(CLDRD_EINVAL, "shard_adopt: buffer must be 16-byte aligned");
    s->xb = const_cast<float*>(rows_dev);
    s->xb_owned = false;
    s->finalized = false;
    return CLDRD_OK;
}

int cldrd_shard_load_file(cldrd_shard* s, const char* path) {
    if (!s || !path) return fail(CLDRD_EINVAL, "shard_load_file: NULL");
    IndexFileInfo info;
    int rc = probe_index_file(path, &info);
    if (rc) return rc;
    if (info.d != s->d) return fail(CLDRD_EINVAL, "shard_load_file: file d=%d, shard d=%d", info.d, s->d);
    if (s->row0 + s->nrows > info.ntotal)
        return fail(CLDRD_EINVAL, "shard_load_file: rows [%lld,+%lld) outside ntotal=%lld", (long long)s->row0,
                    (long long)s->nrows, (long long)info.ntotal);
    if (s->xb && !s->xb_owned) return fail(CLDRD_ESTATE, "shard_load_file: rows were adopted");
    DeviceGuard g(s->device);
    NvtxRange nvtx("cldrd::shard_load_file");
    if ((rc = ensure_rows(s))) return rc;
    // pread -> page-locked ring -> cudaMemcpyAsync: the payload starts at an odd byte offset (82), so it is staged
    // rather than mapped.  One reader is bound by a single core's page-cache copy (a few GB/s) while the PCIe link
    // takes > 50 GB/s, so several readers each stream their own interleaved pieces through their own two buffers
    // and their own stream.  Everything a reader owns is released on every path out of it.
    const size_t row_bytes = size_t(s->d) * 4;
    const size_t piece_rows = std::max<size_t>(1, (size_t(32) << 20) / row_bytes);
    const int64_t npieces = (s->nrows + int64_t(piece_rows) - 1) / int64_t(piece_rows);
    int readers = 4;
    if (const char* e = getenv("CLDRD_LOAD_THREADS")) readers = std::max(1, atoi(e));
    readers = int(std::max<int64_t>(1, std::min<int64_t>(readers, npieces)));
    std::vector<int> rcs(size_t(readers), CLDRD_OK);
    std::vector<std::string> errs{size_t(readers)};
    const int device = s->device;
    auto reader = [&](int t) {
        auto bail = [&](int code, const char* what, const char* detail) {
            rcs[size_t(t)] = code;
            errs[size_t(t)] = std::string(what) + (detail ? detail : "");
        };
        if (cudaSetDevice(device) != cudaSuccess) return bail(CLDRD_ECUDA, "cudaSetDevice failed", nullptr);
        char* pin[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        int fd = -1;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaHostAlloc(&pin[i], piece_rows * row_bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            bail(CLDRD_ECUDA, "staging buffers: ", cudaGetErrorString(e));
        } else if ((fd = open(path, O_RDONLY)) < 0) {
            bail(CLDRD_EIO, "cannot open ", path);
        }
        int b = 0;
        for (int64_t p = t; p < npieces && !rcs[size_t(t)]; p += readers, b ^= 1) {
            const int64_t r = p * int64_t(piece_rows);
            const size_t n = size_t(std::min<int64_t>(int64_t(piece_rows), s->nrows - r));
            cudaEventSynchronize(ev[b]);     // the copy that last used this buffer is done
            const size_t want = n * row_bytes;
            size_t got = 0;
            const off_t off = off_t(info.data_off) + off_t(s->row0 + r) * off_t(row_bytes);
            while (got < want) {
                ssize_t x = pread(fd, pin[b] + got, want - got, off + off_t(got));
                if (x <= 0) {
                    bail(CLDRD_EIO, "short read from ", path);
                    break;
                }
                got += size_t(x);
            }
            if (rcs[size_t(t)]) break;
            if (cudaMemcpyAsync(reinterpret_cast<char*>(s->xb) + size_t(r) * row_bytes, pin[b], want, cudaMemcpyHostToDevice, st) !=
                cudaSuccess) {
                cudaGetLastError();
                bail(CLDRD_ECUDA, "H2D copy failed", nullptr);
                break;
            }
            cudaEventRecord(ev[b], st);
        }
        if (fd >= 0) close(fd);
        if (st) cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i) {
            if (pin[i]) cudaFreeHost(pin[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (st) cudaStreamDestroy(st);
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < readers; ++t) th.emplace_back(reader, t);
        reader(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < readers; ++t)
        if (rcs[size_t(t)]) return fail(rcs[size_t(t)], "shard_load_file: %s", errs[size_t(t)].c_str());
    if (info.has_ids) {
        std::vector<int64_t> ids(size_t(s->nrows));
        if ((rc = cldrd_index_read_ids(path, s->row0, s->nrows, ids.data()))) return rc;
        if ((rc = cldrd_shard_set_ids(s, ids.data()))) return rc;
    }
    s->finalized = false;
    return CLDRD_OK;
}

int cldrd_shard_set_ids(cldrd_shard* s, const int64_t* ids_host) {
    if (!s) return fail(CLDRD_EINVAL, "shard_set_ids: NULL");
    DeviceGuard g(s->device);
    cudaFree(s->ids);
    s->ids = nullptr;
    if (!ids_host || s->nrows == 0) return CLDRD_OK;
    CU_TRY(cudaMalloc(&s->ids, size_t(s->nrows) * sizeof(int64_t)));
    CU_TRY(cudaMemcpy(s->ids, ids_host, size_t(s->nrows) * sizeof(int64_t), cudaMemcpyHostToDevice));
    return CLDRD_OK;
}

int cldrd_shard_finalize(cldrd_shard* s, void* cuda_stream) {
    if (!s) return fail(CLDRD_EINVAL, "shard_finalize: NULL");
    if (s->nrows > 0 && !s->xb) return fail(CLDRD_ESTATE, "shard_finalize: no rows uploaded / loaded / adopted");
    DeviceGuard g(s->device);
    NvtxRange nvtx("cldrd::shard_finalize");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // can TMA describe this shape?  (row pitch multiple of 16 bytes, 16-byte aligned base)
    s->scan_eff = s->scan;
    s->vec4 = (s->d % 4 == 0) && (reinterpret_cast<uintptr_t>(s->xb) % 16 == 0);
    if (is_tc(s->scan)) {
        const int esz = (s->scan == CLDRD_SCAN_TC_TF32) ? 4 : 2;
        if ((s->d * esz) % 16 != 0 || !s->vec4) s->scan_eff = CLDRD_SCAN_SIMT_F32;
    }
    free_workspace(s);
    cudaFree(s->xlp);
    s->xlp = nullptr;
    const int lp = lp_kind_of(s->scan_eff);
    if (lp && s->nrows) CU_TRY(cudaMalloc(&s->xlp, size_t(s->nrows) * s->d * 2));
    unsigned int* d_max = nullptr;
    CU_TRY(cudaMalloc(&d_max, 2 * sizeof(unsigned int)));
    unsigned int h_max[2] = {0, 0};
    {
        cudaError_t e = cudaMemsetAsync(d_max, 0, 2 * sizeof(unsigned int), st);
        if (e == cudaSuccess && s->nrows) {
            const int threads = 256;
            const int blocks = int(std::min<int64_t>((s->nrows * 32 + threads - 1) / threads, int64_t(s->num_sms) * 16));
            index_prep_kernel<<<blocks, threads, 0, st>>>(s->xb, s->nrows, s->d, lp, s->xlp, d_max, d_max + 1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(d_max);
        if (e != cudaSuccess) return fail(CLDRD_ECUDA, "shard_finalize: index preparation pass failed: %s", cudaGetErrorString(e));
    }
    float n2, mx;
    memcpy(&n2, &h_max[0], 4);
    memcpy(&mx, &h_max[1], 4);
    s->bmax_norm = std::sqrt(n2) * 1.0001f;
    s->bmax_abs = mx;
    if (s->scan_eff == CLDRD_SCAN_TC_F16 && !(mx < 65504.f))
        return fail(CLDRD_EINVAL, "index values (max |x| = %g) exceed the fp16 range; use the bf16 or tf32 scan", double(mx));
    if (is_tc(s->scan_eff) && s->nrows) {
        const void* base = lp ? s->xlp : static_cast<const void*>(s->xb);
        int rc = make_tensor_map(&s->tmB, s->scan_eff, base, s->nrows, s->d, TC_BN);
        if (rc) return rc;
        if ((rc = make_tensor_map(&s->tmBh, s->scan_eff, base, s->nrows, s->d, TC_BN / 2))) return rc;
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<0, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<0, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<0, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<1, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<1, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<1, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<2, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<2, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc_kernel<2, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM_BYTES)));
    }
    if (is_tc(s->scan_eff) && s->nrows) {
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<0, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<0, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<0, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<1, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<1, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<1, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<2, TC_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<2, TC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
        CU_TRY(cudaFuncSetAttribute(scan_tc2_kernel<2, TC_MAXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC2_SMEM_BYTES)));
    }
    preload_kernels(s->device);
    int rc = ensure_workspace(s);
    if (rc) return rc;
    CU_TRY(cudaFuncSetAttribute(select_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(select_smem(s))));
    CU_TRY(cudaFuncSetAttribute(sample_topj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kDensePiece) * 8));
    CU_TRY(cudaFuncSetAttribute(rescore_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                int(size_t(s->ws_keep_cap) * 8 + size_t(s->d) * 4 + 16)));

    s->finalized = true;
    return CLDRD_OK;
}

int64_t cldrd_shard_nrows(const cldrd_shard* s) { return s ? s->nrows : 0; }
int32_t cldrd_shard_dim(const cldrd_shard* s) { return s ? s->d : 0; }
int32_t cldrd_shard_scan(const cldrd_shard* s) { return s ? s->scan_eff : -1; }
int64_t cldrd_shard_scan_bytes(const cldrd_shard* s) {
    if (!s) return 0;
    const int esz = lp_kind_of(s->scan_eff) ? 2 : 4;
    return s->nrows * int64_t(s->d) * esz;
}

static int search_common(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k, int32_t translate_ids,
                         int seed_mode, const float* seed_dev, float* out_scores_dev, int64_t* out_ids_dev,
                         float* eps_out_dev, void* cuda_stream) {
    if (!s || nq < 0 || (nq && (!q_dev || !out_scores_dev || !out_ids_dev)))
        return fail(CLDRD_EINVAL, "search: NULL argument");
    if (k < 1 || k > CLDRD_MAX_K) return fail(CLDRD_EINVAL, "search: k=%d outside [1,%d]", k, CLDRD_MAX_K);
    if (!s->finalized) return fail(CLDRD_ESTATE, "search: shard not finalized");
    if (reinterpret_cast<uintptr_t>(q_dev) % 16 && is_tc(s->scan_eff))
        return fail(CLDRD_EINVAL, "search: query buffer must be 16-byte aligned");
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    SearchTotals totals;
    unsigned long long tot[ST_COUNT] = {0};
    s->ev_used = 0;
    s->scan_ms = 0.0;
    s->scan_launches = 0;
    s->ev_rows.clear();
    s->ev_ms.clear();
    if (seed_mode == 1 && (s->nrows < kSeedMinRows || s->no_seed)) seed_mode = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += kQueryBatch) {
        const int nb = int(std::min<int64_t>(kQueryBatch, nq - q0));
        CU_TRY(cudaMemsetAsync(s->w_stats, 0, ST_COUNT * sizeof(unsigned long long), st));
        int rc = search_batch(s, q_dev + size_t(q0) * s->d, nb, k, translate_ids != 0, seed_mode,
                              seed_dev ? seed_dev + q0 : nullptr, out_scores_dev ? out_scores_dev + size_t(q0) * k : nullptr,
                              out_ids_dev ? out_ids_dev + size_t(q0) * k : nullptr, eps_out_dev ? eps_out_dev + q0 : nullptr,
                              st, &totals);
        if (rc) return rc;
        for (int i = 0; i < ST_COUNT; ++i) {
            if (i == ST_MAX_LIST) tot[i] = std::max(tot[i], s->h_stats[i]);
            else tot[i] += s->h_stats[i];
        }
    }
    if (s->profile) {  // every batch ended with a stream synchronisation: the events are complete
        for (size_t i = 0; i + 1 < s->ev_used; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]) == cudaSuccess) s->scan_ms += ms;
            s->ev_ms.push_back(ms);
            s->scan_launches++;
        }
    }
    s->stats[0] = totals.launches;
    s->stats[1] = totals.chunks;
    s->stats[2] = totals.fallback_queries;
    s->stats[3] = int64_t(tot[ST_RESCORED]);
    s->stats[4] = int64_t(tot[ST_SURVIVORS]);
    s->stats[5] = int64_t(tot[ST_MAX_LIST]);
    s->stats[6] = int64_t(tot[ST_TILES]);
    s->stats[7] = int64_t(tot[ST_EXACT_COMPACT]);
    return CLDRD_OK;
}

int cldrd_search_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k, int32_t translate_ids,
                     float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    return search_common(s, q_dev, nq, k, translate_ids, 1, nullptr, out_scores_dev, out_ids_dev, nullptr, cuda_stream);
}

int cldrd_search_dev_seeded(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k, int32_t translate_ids,
                            const float* seed_dev, float* out_scores_dev, int64_t* out_ids_dev, float* eps_out_dev,
                            void* cuda_stream) {
    return search_common(s, q_dev, nq, k, translate_ids, seed_dev ? 2 : 0, seed_dev, out_scores_dev, out_ids_dev,
                         eps_out_dev, cuda_stream);
}

// cudaMalloc may carve a small request out of a larger driver allocation; the IPC handle then names
// the whole allocation and the peer's mapping starts at ITS base.  The handle we hand out therefore
// carries the offset of our pointer inside that allocation (bytes 64..71).
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetAddressRangeFn get_address_range_fn() {
    static GetAddressRangeFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<GetAddressRangeFn>(p);
    return fn;
}
static std::mutex g_peer_mu;
static std::map<void*, void*> g_peer_base;   // pointer handed out by cldrd_peer_open -> mapping base

static_assert(kQueryBatch == CLDRD_QUERY_BATCH, "header and engine disagree on the query batch");

int cldrd_peer_alloc(int device, int64_t nbytes, void** out_ptr, void* out_handle) {
    if (!out_ptr || !out_handle || nbytes < 1) return fail(CLDRD_EINVAL, "peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(uint64_t) == CLDRD_PEER_HANDLE_BYTES, "IPC handle size");
    DeviceGuard g(device);
    *out_ptr = nullptr;
    // whole 2 MiB pages: smaller requests are carved out of a pool block that IPC would export as a whole
    const size_t page = size_t(2) << 20;
    const size_t alloc_bytes = (size_t(nbytes) + page - 1) / page * page;
    cudaError_t e = cudaMalloc(out_ptr, alloc_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CLDRD_ENOMEM : CLDRD_ECUDA, "peer_alloc: cudaMalloc(%lld): %s",
                    (long long)nbytes, cudaGetErrorString(e));
    }
    // slice rows past the last query are merged too and nobody writes them: all-ones = row -1 = padding
    cudaMemset(*out_ptr, 0xFF, alloc_bytes);
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *out_ptr);
    GetAddressRangeFn range = get_address_range_fn();
    CUdeviceptr base = 0;
    size_t size = 0;
    const bool ranged = range && range(&base, &size, reinterpret_cast<CUdeviceptr>(*out_ptr)) == CUDA_SUCCESS;
    if (e != cudaSuccess || !ranged) {
        cudaGetLastError();
        cudaFree(*out_ptr);
        *out_ptr = nullptr;
        return fail(CLDRD_ECUDA, "peer_alloc: %s", e != cudaSuccess ? cudaGetErrorString(e) : "cuMemGetAddressRange failed");
    }
    const uint64_t offset = uint64_t(reinterpret_cast<CUdeviceptr>(*out_ptr) - base);
    memcpy(out_handle, &h, sizeof(h));
    memcpy(static_cast<char*>(out_handle) + sizeof(h), &offset, sizeof(offset));
    return CLDRD_OK;
}

int cldrd_peer_free(int device, void* ptr) {
    if (!ptr) return CLDRD_OK;
    DeviceGuard g(device);
    CU_TRY(cudaFree(ptr));
    return CLDRD_OK;
}

int cldrd_peer_open(int device, const void* handle, void** out_ptr) {
    if (!handle || !out_ptr) return fail(CLDRD_EINVAL, "peer_open: bad argument");
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    uint64_t offset = 0;
    memcpy(&h, handle, sizeof(h));
    memcpy(&offset, static_cast<const char*>(handle) + sizeof(h), sizeof(offset));
    *out_ptr = nullptr;
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CLDRD_ECUDA, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    *out_ptr = static_cast<char*>(base) + offset;
    std::lock_guard<std::mutex> lk(g_peer_mu);
    g_peer_base[*out_ptr] = base;
    return CLDRD_OK;
}

int cldrd_peer_close(int device, void* ptr) {
    if (!ptr) return CLDRD_OK;
    DeviceGuard g(device);
    void* base = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_peer_mu);
        auto it = g_peer_base.find(ptr);
        if (it == g_peer_base.end()) return fail(CLDRD_EINVAL, "peer_close: not a pointer from cldrd_peer_open");
        base = it->second;
        g_peer_base.erase(it);
    }
    CU_TRY(cudaIpcCloseMemHandle(base));
    return CLDRD_OK;
}

int cldrd_peer_copy(int device, void* dst, const void* src, int64_t nbytes, void* cuda_stream) {
    if (nbytes < 0 || (nbytes && (!dst || !src))) return fail(CLDRD_EINVAL, "peer_copy: bad argument");
    if (nbytes == 0) return CLDRD_OK;
    DeviceGuard g(device);
    CU_TRY(cudaMemcpyAsync(dst, src, size_t(nbytes), cudaMemcpyDefault, static_cast<cudaStream_t>(cuda_stream)));
    return CLDRD_OK;
}

// ---- node-wide sharded search (include/cldrd.h "Sharded search on one node") ----------------------------------

}  // extern "C"

struct cldrd_node {
    int device = 0, world = 1, rank = 0, cap_k = 0, d = 0;
    NodeLayout lay;
    char* block = nullptr;                         // own exchange block
    unsigned char handle[CLDRD_PEER_HANDLE_BYTES];
    char* peer[CLDRD_MAX_PEERS];                   // every rank's block as this process sees it (own included)
    bool opened[CLDRD_MAX_PEERS];                  // mapped by cldrd_peer_open: closed again by cldrd_node_detach
    uint32_t epoch = 0;                            // barriers enqueued so far (the same sequence on every rank)
    unsigned long long timeout_ns = 20000000000ull;
    static constexpr int kRing = 4;                // batches in flight
    BatchStatus* h_status = nullptr;               // page-locked [kRing], written by node_tail_kernel
    BatchStatus* d_status = nullptr;               // the same memory as the device addresses it
    cudaEvent_t done[kRing];
    cudaEvent_t phase[kRing][7];                   // [6]: after the counts barrier (splits phase 2)
    struct Slot {
        int nq = 0, k = 0;
        bool seeded = false;
        int64_t launches = 0, chunks = 0, fallback_queries = 0;
        size_t ev_lo = 0, ev_hi = 0;
    } slot[kRing];
    int64_t seq_begin = 0, seq_end = 0;
    int wait_mode = 0;                             // cldrd_node_set_wait_mode: 0 spin, 1 sleep-poll
    int out_sets = 0;                              // cldrd_node_set_outputs
    float* set_scores[CLDRD_MAX_OUT_SETS];
    long long* set_ids[CLDRD_MAX_OUT_SETS];
    double phase_ms[7] = {0, 0, 0, 0, 0, 0, 0};   // [5]: device idle before this batch, [6]: counts + barrier part of [2]
};

namespace {

int node_barrier(cldrd_shard* s, cldrd_node* n, cudaStream_t st) {
    BarrierParams b{};
    for (int p = 0; p < n->world; ++p) b.peer_flags[p] = reinterpret_cast<uint32_t*>(n->peer[p] + n->lay.flags);
    b.my_flags = reinterpret_cast<const uint32_t*>(n->block + n->lay.flags);
    b.world = n->world;
    b.rank = n->rank;
    b.epoch = ++n->epoch;
    b.timeout_ns = n->timeout_ns;
    b.err = s->w_stats + ST_KERNEL_ERR;
    node_barrier_kernel<<<1, 32, 0, st>>>(b);
    CU_TRY(cudaGetLastError());
    return CLDRD_OK;
}

PeerPtrs node_ptrs(const cldrd_node* n, size_t off) {
    PeerPtrs pp{};
    for (int p = 0; p < n->world; ++p) pp.p[p] = n->peer[p] + off;
    return pp;
}

}  // namespace

extern "C" {

int64_t cldrd_node_block_bytes(int32_t world, int32_t max_k, int32_t d) {
    if (world < 1 || world > CLDRD_MAX_PEERS || max_k < 1 || max_k > CLDRD_MAX_K || d < 0) return 0;
    return int64_t(node_layout(world, max_k, d).total);
}

int cldrd_node_create(cldrd_node** out, int device, int32_t world, int32_t rank, int32_t max_k, int32_t d) {
    if (!out || world < 1 || world > CLDRD_MAX_PEERS || rank < 0 || rank >= world || max_k < 1 || max_k > CLDRD_MAX_K || d < 0 ||
        d % 4 != 0)
        return fail(CLDRD_EINVAL, "node_create: bad argument (world=%d rank=%d max_k=%d d=%d; d a multiple of 4 or 0)", world,
                    rank, max_k, d);
    DeviceGuard g(device);
    preload_kernels(device);
    auto* n = new cldrd_node();
    n->device = device;
    n->world = world;
    n->rank = rank;
    n->cap_k = max_k;
    n->d = d;
    n->lay = node_layout(world, max_k, d);
    for (int p = 0; p < CLDRD_MAX_PEERS; ++p) {
        n->peer[p] = nullptr;
        n->opened[p] = false;
    }
    if (const char* e = getenv("CLDRD_BARRIER_TIMEOUT_MS")) n->timeout_ns = 1000000ull * (unsigned long long)std::max(1, atoi(e));
    {   // the merge kernel's shared memory is sized once, for the largest k this node serves (no attribute changes
        // between the launches of a batch)
        int k_pad = 2;
        while (k_pad < max_k) k_pad <<= 1;
        const size_t smem = std::min<size_t>((size_t(world) * max_k + size_t(k_pad)) * 8, 200 * 1024);
        cudaError_t e = cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) {
            cudaGetLastError();
            delete n;
            return fail(CLDRD_ECUDA, "node_create: %s", cudaGetErrorString(e));
        }
    }
    void* blk = nullptr;
    int rc = cldrd_peer_alloc(device, int64_t(n->lay.total), &blk, n->handle);
    if (rc) {
        delete n;
        return rc;
    }
    n->block = static_cast<char*>(blk);
    n->peer[rank] = n->block;
    // flags 0 = no barrier passed yet, keys 0 = padding
    cudaError_t e = cudaMemset(n->block, 0, n->lay.total);
    if (e == cudaSuccess) e = cudaHostAlloc(&n->h_status, sizeof(BatchStatus) * cldrd_node::kRing, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&n->d_status), n->h_status, 0);
    for (int i = 0; i < cldrd_node::kRing && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&n->done[i], cudaEventDisableTiming);
        for (int j = 0; j < 7 && e == cudaSuccess; ++j) e = cudaEventCreate(&n->phase[i][j]);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(n->block);
        if (n->h_status) cudaFreeHost(n->h_status);
        delete n;   // (events of a half-built ring are reclaimed with the context)
        return fail(CLDRD_ECUDA, "node_create: %s", cudaGetErrorString(e));
    }
    *out = n;
    return CLDRD_OK;
}

int cldrd_node_handle(const cldrd_node* n, void* out_handle) {
    if (!n || !out_handle) return fail(CLDRD_EINVAL, "node_handle: NULL");
    memcpy(out_handle, n->handle, CLDRD_PEER_HANDLE_BYTES);
    return CLDRD_OK;
}

void* cldrd_node_block(const cldrd_node* n) { return n ? n->block : nullptr; }

int cldrd_node_attach(cldrd_node* n, int32_t peer_rank, const void* handle, void* ptr, int32_t peer_device) {
    if (!n || peer_rank < 0 || peer_rank >= n->world || peer_rank == n->rank || (!handle) == (!ptr))
        return fail(CLDRD_EINVAL, "node_attach: bad argument (peer %d; exactly one of handle / pointer)", peer_rank);
    if (n->peer[peer_rank]) return fail(CLDRD_ESTATE, "node_attach: rank %d is attached already", peer_rank);
    DeviceGuard g(n->device);
    if (handle) {
        void* m = nullptr;
        int rc = cldrd_peer_open(n->device, handle, &m);
        if (rc) return rc;
        n->peer[peer_rank] = static_cast<char*>(m);
        n->opened[peer_rank] = true;
        return CLDRD_OK;
    }
    if (peer_device >= 0 && peer_device != n->device) {   // same process, another GPU: direct peer access
        int can = 0;
        CU_TRY(cudaDeviceCanAccessPeer(&can, n->device, peer_device));
        if (!can) return fail(CLDRD_ECUDA, "node_attach: device %d cannot access device %d", n->device, peer_device);
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            return fail(CLDRD_ECUDA, "node_attach: cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
        }
        cudaGetLastError();
    }
    n->peer[peer_rank] = static_cast<char*>(ptr);
    return CLDRD_OK;
}

int cldrd_node_detach(cldrd_node* n) {
    if (!n) return CLDRD_OK;
    DeviceGuard g(n->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < n->world; ++p) {
        if (n->opened[p]) cldrd_peer_close(n->device, n->peer[p]);
        if (p != n->rank) n->peer[p] = nullptr;
        n->opened[p] = false;
    }
    return CLDRD_OK;
}

void cldrd_node_destroy(cldrd_node* n) {
    if (!n) return;
    cldrd_node_detach(n);
    DeviceGuard g(n->device);
    for (int i = 0; i < cldrd_node::kRing; ++i) {
        cudaEventDestroy(n->done[i]);
        for (int j = 0; j < 7; ++j) cudaEventDestroy(n->phase[i][j]);
    }
    if (n->h_status) cudaFreeHost(n->h_status);
    cudaFree(n->block);
    delete n;
}

int cldrd_node_query_ptr(const cldrd_node* n, void** q_dev) {
    if (!n || !q_dev) return fail(CLDRD_EINVAL, "node_query_ptr: NULL");
    if (n->d == 0) return fail(CLDRD_ESTATE, "node_query_ptr: the node was created without a query buffer (d = 0)");
    *q_dev = n->block + n->lay.qx;
    return CLDRD_OK;
}

int cldrd_node_spread_queries(cldrd_shard* s, cldrd_node* n, int64_t row0, int64_t nrows, void* cuda_stream) {
    if (!s || !n || row0 < 0 || nrows < 0 || row0 + nrows > kQueryBatch)
        return fail(CLDRD_EINVAL, "node_spread_queries: rows [%lld, +%lld) outside the batch", (long long)row0, (long long)nrows);
    if (n->d == 0) return fail(CLDRD_ESTATE, "node_spread_queries: the node was created without a query buffer (d = 0)");
    if (!s->finalized) return fail(CLDRD_ESTATE, "node_spread_queries: shard not finalized");
    for (int p = 0; p < n->world; ++p)
        if (!n->peer[p]) return fail(CLDRD_ESTATE, "node_spread_queries: rank %d is not attached", p);
    DeviceGuard g(n->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (nrows > 0 && n->world > 1) {
        const size_t off16 = size_t(row0) * n->d / 4, n16 = size_t(nrows) * n->d / 4;
        const int blocks = int(std::min<size_t>((n16 + 255) / 256, size_t(s->num_sms) * 4));
        node_spread_kernel<<<blocks, 256, 0, st>>>(node_ptrs(n, n->lay.qx), n->world, n->rank, off16, n16);
        CU_TRY(cudaGetLastError());
    }
    // every rank's part has landed everywhere before anybody reads the batch.  (The counters the barrier reports into
    // are reset by the search that follows, so a rank that never arrives shows up there as its first barrier.)
    return node_barrier(s, n, st);
}

int cldrd_node_result_ptrs(const cldrd_node* n, int32_t owner_rank, void** scores, void** ids) {
    if (!n || owner_rank < 0 || owner_rank >= n->world || !n->peer[owner_rank] || !scores || !ids)
        return fail(CLDRD_EINVAL, "node_result_ptrs: bad argument or rank %d not attached", owner_rank);
    *scores = n->peer[owner_rank] + n->lay.res_d;
    *ids = n->peer[owner_rank] + n->lay.res_i;
    return CLDRD_OK;
}

}  // extern "C"

// Everything below is validated BEFORE the first launch: once a batch has started to enqueue, every rank must enqueue the
// same sequence of barriers.  A CUDA failure in the middle (it would be a launch failure: nothing here allocates)
// returns the error and leaves this rank's barrier count behind its peers'; their next barrier then ends in the
// watchdog's CLDRD_ECUDA on every rank instead of a result -- the node has to be rebuilt.
static int node_search_begin_impl(cldrd_shard* s, cldrd_node* n, const float* q_dev, int64_t nq, int32_t k, int32_t seeded,
                                  float* out_scores, int64_t* out_ids, bool use_sets, int32_t out_select, int64_t out_row0,
                                  const int32_t* out_rows_dev, const int64_t* id_map_dev, void* cuda_stream) {
    if (!s || !n || !q_dev || (!use_sets && (!out_scores || !out_ids)) || nq < 1 || nq > kQueryBatch)
        return fail(CLDRD_EINVAL, "node_search_begin: bad argument (1 <= nq <= %d)", kQueryBatch);
    if (use_sets && (n->out_sets < 1 || out_select >= n->out_sets || out_row0 < 0))
        return fail(CLDRD_EINVAL, "node_search_begin_set: set %d of %d registered sets, first row %lld", out_select, n->out_sets,
                    (long long)out_row0);
    if (k < 1 || k > n->cap_k) return fail(CLDRD_EINVAL, "node_search_begin: k=%d outside [1,%d] of this node", k, n->cap_k);
    if (!s->finalized) return fail(CLDRD_ESTATE, "node_search_begin: shard not finalized");
    if (s->device != n->device) return fail(CLDRD_EINVAL, "node_search_begin: shard and node live on different devices");
    for (int p = 0; p < n->world; ++p)
        if (!n->peer[p]) return fail(CLDRD_ESTATE, "node_search_begin: rank %d is not attached", p);
    if (n->seq_begin - n->seq_end >= cldrd_node::kRing)
        return fail(CLDRD_ESTATE, "node_search_begin: %d batches in flight (call cldrd_node_search_end)", cldrd_node::kRing);
    if (reinterpret_cast<uintptr_t>(q_dev) % 16 && is_tc(s->scan_eff))
        return fail(CLDRD_EINVAL, "search: query buffer must be 16-byte aligned");
    int k_pad = 2;
    while (k_pad < k) k_pad <<= 1;
    const size_t merge_smem = (size_t(n->world) * k + size_t(k_pad)) * 8;
    if (merge_smem > 200 * 1024) return fail(CLDRD_EINVAL, "node_search_begin: world*k=%d too large for the merge", n->world * k);
    DeviceGuard g(s->device);
    NvtxRange nvtx("cldrd::node_search_begin");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int slot = int(n->seq_begin % cldrd_node::kRing);
    cldrd_node::Slot& sl = n->slot[slot];
    if (n->seq_begin == n->seq_end) {   // nothing in flight: the scan-timing events start over
        s->ev_used = 0;
        s->ev_rows.clear();
    }
    sl = cldrd_node::Slot();
    sl.nq = int(nq);
    sl.k = k;
    sl.seeded = seeded != 0 && !s->no_seed;
    sl.ev_lo = s->ev_used;
    const bool sd = sl.seeded;
    const int world = n->world, rank = n->rank;
    const size_t plane = size_t(kQueryBatch) * CLDRD_SEED_J;   // elements per plane of the sample / count buffers
    CU_TRY(cudaMemsetAsync(s->w_stats, 0, ST_COUNT * sizeof(unsigned long long), st));
    CU_TRY(cudaEventRecord(n->phase[slot][0], st));
    BatchCtx c{};
    c.s = s;
    c.st = st;
    c.q = q_dev;
    c.nq = int(nq);
    c.k = k;
    int rc = launch_prep(c);
    if (rc) return rc;
    // nobody raises a query of this batch before the first barrier below (seeded) / at all (unseeded)
    CU_TRY(cudaMemsetAsync(n->block + n->lay.qfail, 0, size_t(nq) * sizeof(int), st));
    if (use_sets && out_select >= 0) {   // this rank picks the output set; every rank reads it behind the next barrier
        node_publish_kernel<<<1, 32, 0, st>>>(node_ptrs(n, n->lay.ctrl), n->world, out_select);
        CU_TRY(cudaGetLastError());
    }
    int64_t extra = 0;
    SearchTotals totals;
    if (sd) {
        // 1. sample scores -> plane [rank] of every rank's sample buffer; barrier; levels + seed from the union
        const SamplePlan sp = sample_plan(s, k);
        const double frac = sp.tiles > 0 ? double(sp.tiles) * TC_BN / double(std::max<int64_t>(s->nrows, 1)) : 1.0;
        c.seed_rank = double(CLDRD_SEED_J) / frac;
        c.expected_surv = c.seed_rank / world;               // the seed is the J-th best of the union of all samples
        if ((rc = run_sample(c, node_ptrs(n, n->lay.topj), world, size_t(rank) * plane))) return rc;
        if ((rc = node_barrier(s, n, st))) return rc;
        levels_seed_kernel<<<unsigned(nq), 128, size_t(world) * CLDRD_SEED_J * sizeof(float), st>>>(
            reinterpret_cast<const float*>(n->block + n->lay.topj), plane, world, CLDRD_SEED_J, s->tune_seed_bias, s->w_levels,
            s->w_seed, s->w_thr, s->w_list_len);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(n->phase[slot][1], st));
        // 2. fused scan + filter with that seed, select; candidates above every level -> plane [rank] everywhere
        if ((rc = run_chunks(c, PASS_SEEDED))) return rc;
        CU_TRY(cudaEventRecord(n->phase[slot][2], st));
        count_levels_peers_kernel<<<unsigned(nq), 256, 0, st>>>(s->w_list, s->w_list_len, s->ws_keep_cap, s->w_fail, s->w_levels,
                                                                CLDRD_SEED_J, node_ptrs(n, n->lay.counts), world,
                                                                size_t(rank) * plane);
        CU_TRY(cudaGetLastError());
        if ((rc = node_barrier(s, n, st))) return rc;
        CU_TRY(cudaEventRecord(n->phase[slot][6], st));
        extra += 4;
    } else {
        CU_TRY(cudaEventRecord(n->phase[slot][1], st));
        if ((rc = run_chunks(c, PASS_PROGRESSIVE))) return rc;   // (host-synchronous: sizes its chunks from the first piece)
        CU_TRY(cudaEventRecord(n->phase[slot][2], st));
        CU_TRY(cudaEventRecord(n->phase[slot][6], st));
    }
    // 3. exact re-score of what can still reach the global top-k; every list goes to the rank that merges the query
    const int64_t slice = (nq + world - 1) / world;
    s->sc.world = world;
    s->sc.rank = rank;
    s->sc.slice = slice;
    s->sc.key_stride = n->cap_k;
    s->sc.raise_fail = sd;
    s->sc.keys = node_ptrs(n, n->lay.xkeys);
    s->sc.len = node_ptrs(n, n->lay.xlen);
    s->sc.qfail = node_ptrs(n, n->lay.qfail);
    CountedCut cc;
    cc.planes = reinterpret_cast<const int*>(n->block + n->lay.counts);
    cc.stride = plane;
    cc.parts = world;
    cc.levels = s->w_levels;
    rc = launch_rescore(c, nullptr, nullptr, false, nullptr, s->w_fail, st, 0, sd ? &cc : nullptr);
    if (!rc && !sd) {
        // unseeded batches are host-driven anyway: a query whose survivors overflowed is redone densely right here
        rc = read_stats(s, st);
        if (!rc && s->h_stats[ST_RANGE_ERR])
            rc = fail(CLDRD_EINVAL, "query values exceed the fp16 range; use the bf16 or tf32 scan");
        if (!rc && s->h_stats[ST_FAILED] != 0)
            rc = run_fallbacks(s, q_dev, int(nq), k, false, nullptr, nullptr, st, false, &totals);
    }
    s->sc.world = 0;
    if (rc) return rc;
    CU_TRY(cudaEventRecord(n->phase[slot][3], st));
    if ((rc = node_barrier(s, n, st))) return rc;
    // 4. merge + seed check + ids of this rank's slice, stored where the caller wants the result
    const int64_t lo = int64_t(rank) * slice;
    const int64_t n_mine = std::max<int64_t>(0, std::min<int64_t>(slice, nq - lo));
    if (n_mine > 0) {
        MergeKeysParams m{};
        m.xkeys = reinterpret_cast<const uint64_t*>(n->block + n->lay.xkeys);
        m.xlen = reinterpret_cast<const int*>(n->block + n->lay.xlen);
        m.parts = world;
        m.slice = slice;
        m.key_stride = n->cap_k;
        m.k = k;
        m.k_pad = k_pad;
        m.q_lo = lo;
        m.seed = sd ? s->w_seed : nullptr;
        m.band = s->w_band;
        m.id_map = reinterpret_cast<const long long*>(id_map_dev);
        m.out_scores = out_scores;
        m.out_ids = reinterpret_cast<long long*>(out_ids);
        if (use_sets) {
            m.out_sets = n->out_sets;
            m.ctrl = reinterpret_cast<const int*>(n->block + n->lay.ctrl);
            m.out_row0 = out_row0;
            for (int i = 0; i < n->out_sets; ++i) {
                m.set_scores[i] = n->set_scores[i];
                m.set_ids[i] = n->set_ids[i];
            }
        }
        m.out_rows = out_rows_dev;
        m.qfail = node_ptrs(n, n->lay.qfail);
        m.world = world;
        merge_keys_kernel<<<unsigned(n_mine), 512, merge_smem, st>>>(m);
        CU_TRY(cudaGetLastError());
        extra += 1;
    }
    CU_TRY(cudaEventRecord(n->phase[slot][4], st));
    if ((rc = node_barrier(s, n, st))) return rc;
    // 5. every slice has landed and every raised query is known everywhere: status for the host
    node_tail_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const int*>(n->block + n->lay.qfail), int(nq), s->w_stats,
                                         n->d_status + slot);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(n->phase[slot][5], st));
    CU_TRY(cudaEventRecord(n->done[slot], st));
    extra += 3;
    sl.launches = c.launches + totals.launches + extra;
    sl.chunks = c.chunks + totals.chunks;
    sl.fallback_queries = totals.fallback_queries;
    sl.ev_hi = s->ev_used;
    n->seq_begin++;
    return CLDRD_OK;
}

extern "C" {

int cldrd_node_search_begin(cldrd_shard* s, cldrd_node* n, const float* q_dev, int64_t nq, int32_t k, int32_t seeded,
                            float* out_scores, int64_t* out_ids, const int32_t* out_rows_dev, const int64_t* id_map_dev,
                            void* cuda_stream) {
    return node_search_begin_impl(s, n, q_dev, nq, k, seeded, out_scores, out_ids, false, -1, 0, out_rows_dev, id_map_dev,
                                  cuda_stream);
}

int cldrd_node_search_begin_set(cldrd_shard* s, cldrd_node* n, const float* q_dev, int64_t nq, int32_t k, int32_t seeded,
                                int32_t out_select, int64_t out_row0, const int32_t* out_rows_dev, const int64_t* id_map_dev,
                                void* cuda_stream) {
    return node_search_begin_impl(s, n, q_dev, nq, k, seeded, nullptr, nullptr, true, out_select, out_row0, out_rows_dev,
                                  id_map_dev, cuda_stream);
}

int cldrd_node_set_outputs(cldrd_node* n, int32_t count, void* const* scores, void* const* ids) {
    if (!n || count < 0 || count > CLDRD_MAX_OUT_SETS || (count && (!scores || !ids)))
        return fail(CLDRD_EINVAL, "node_set_outputs: bad argument (count <= %d)", CLDRD_MAX_OUT_SETS);
    if (n->seq_begin != n->seq_end) return fail(CLDRD_ESTATE, "node_set_outputs: batches in flight");
    for (int i = 0; i < count; ++i) {
        if (!scores[i] || !ids[i]) return fail(CLDRD_EINVAL, "node_set_outputs: NULL buffer in set %d", i);
        n->set_scores[i] = static_cast<float*>(scores[i]);
        n->set_ids[i] = static_cast<long long*>(ids[i]);
    }
    n->out_sets = count;
    return CLDRD_OK;
}

int cldrd_node_search_end(cldrd_shard* s, cldrd_node* n, int32_t* nfail_out, int32_t* fail_idx_out, int32_t cap) {
    if (!s || !n) return fail(CLDRD_EINVAL, "node_search_end: NULL");
    if (n->seq_end == n->seq_begin) return fail(CLDRD_ESTATE, "node_search_end: no batch in flight");
    DeviceGuard g(n->device);
    NvtxRange nvtx("cldrd::node_search_end");
    const int slot = int(n->seq_end % cldrd_node::kRing);
    const cldrd_node::Slot& sl = n->slot[slot];
    n->seq_end++;
    if (n->wait_mode == 1) {
        // long searches with host work going on beside them (the run-file writer formats batch i on all cores while
        // batch i+1 is scanned): do not burn a core per rank spinning on the event
        for (;;) {
            cudaError_t q = cudaEventQuery(n->done[slot]);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return fail(CLDRD_ECUDA, "node_search_end: %s", cudaGetErrorString(q));
            usleep(100);
        }
    }
    CU_TRY(cudaEventSynchronize(n->done[slot]));
    const BatchStatus& hs = n->h_status[slot];
    if (hs.stats[ST_KERNEL_ERR] >= kErrBarrierTimeout)
        return fail(CLDRD_ECUDA, "node search: rank %llu did not reach a barrier within %llu ms (peer process gone, or the "
                    "ranks issued different call sequences)", hs.stats[ST_KERNEL_ERR] - kErrBarrierTimeout, n->timeout_ns / 1000000ull);
    if (hs.stats[ST_KERNEL_ERR]) return fail(CLDRD_ECUDA, "scan kernel watchdog fired (code %llu)", hs.stats[ST_KERNEL_ERR]);
    if (hs.stats[ST_RANGE_ERR]) return fail(CLDRD_EINVAL, "query values exceed the fp16 range; use the bf16 or tf32 scan");
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, n->phase[slot][i], n->phase[slot][i + 1]) != cudaSuccess) cudaGetLastError();
        n->phase_ms[i] = ms;
    }
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, n->phase[slot][2], n->phase[slot][6]) != cudaSuccess) cudaGetLastError();
        n->phase_ms[6] = ms;
    }
    n->phase_ms[5] = 0.0;
    if (n->seq_end >= 2) {   // (seq_end was advanced above) the batch before this one: end of its status kernel -> our start
        const int prev = int((n->seq_end - 2) % cldrd_node::kRing);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, n->phase[prev][5], n->phase[slot][0]) == cudaSuccess) n->phase_ms[5] = ms;
        else cudaGetLastError();
    }
    s->scan_ms = 0.0;
    s->scan_launches = 0;
    s->ev_ms.clear();
    if (s->profile) {
        for (size_t i = sl.ev_lo; i + 1 < sl.ev_hi; i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]) == cudaSuccess) s->scan_ms += ms;
            else cudaGetLastError();
            s->ev_ms.push_back(ms);
            s->scan_launches++;
        }
    }
    const int nfail = hs.nfail;
    s->stats[0] = sl.launches;
    s->stats[1] = sl.chunks;
    s->stats[2] = sl.fallback_queries + (sl.seeded ? nfail : 0);
    s->stats[3] = int64_t(hs.stats[ST_RESCORED]);
    s->stats[4] = int64_t(hs.stats[ST_SURVIVORS]);
    s->stats[5] = int64_t(hs.stats[ST_MAX_LIST]);
    s->stats[6] = int64_t(hs.stats[ST_TILES]);
    s->stats[7] = int64_t(hs.stats[ST_EXACT_COMPACT]);
    if (nfail_out) *nfail_out = nfail;
    if (fail_idx_out)
        for (int i = 0; i < nfail && i < cap; ++i) fail_idx_out[i] = hs.idx[i];
    return CLDRD_OK;
}

int cldrd_node_set_wait_mode(cldrd_node* n, int32_t mode) {
    if (!n || mode < 0 || mode > 1) return fail(CLDRD_EINVAL, "node_set_wait_mode: mode 0 (spin) or 1 (sleep-poll)");
    n->wait_mode = mode;
    return CLDRD_OK;
}

int cldrd_node_phase_ms(const cldrd_node* n, double out[7]) {
    if (!n || !out) return fail(CLDRD_EINVAL, "node_phase_ms: NULL");
    for (int i = 0; i < 7; ++i) out[i] = n->phase_ms[i];
    return CLDRD_OK;
}

int cldrd_sample_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int32_t k, float* out_topj_dev, void* cuda_stream) {
    if (!s || !q_dev || !out_topj_dev || nq < 1) return fail(CLDRD_EINVAL, "sample: bad argument");
    if (k < 1 || k > CLDRD_MAX_K) return fail(CLDRD_EINVAL, "sample: k=%d outside [1,%d]", k, CLDRD_MAX_K);
    if (!s->finalized) return fail(CLDRD_ESTATE, "sample: shard not finalized");
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    for (int64_t q0 = 0; q0 < nq; q0 += kQueryBatch) {
        const int nb = int(std::min<int64_t>(kQueryBatch, nq - q0));
        CU_TRY(cudaMemsetAsync(s->w_stats, 0, ST_COUNT * sizeof(unsigned long long), st));
        BatchCtx c{};
        c.s = s;
        c.st = st;
        c.q = q_dev + size_t(q0) * s->d;
        c.nq = nb;
        c.k = k;
        int rc = launch_prep(c);
        if (rc) return rc;
        if ((rc = run_sample(c, out_topj_dev + size_t(q0) * CLDRD_SEED_J))) return rc;
        // no synchronisation here: the caller's all-gather and the seeded search queue up behind
        // this on the same stream; a watchdog code (if any) surfaces at the end of that search
    }
    return CLDRD_OK;
}

int cldrd_seed_from_samples(int device, const float* topj_dev, int32_t parts, int64_t nq, float* seed_out_dev,
                            void* cuda_stream) {
    if (!topj_dev || !seed_out_dev || parts < 1 || nq < 1) return fail(CLDRD_EINVAL, "seed_from_samples: bad argument");
    DeviceGuard g(device);
    seed_from_samples_kernel<<<unsigned((nq * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        topj_dev, parts, int(nq), CLDRD_SEED_J, CLDRD_SEED_J, seed_out_dev);
    CU_TRY(cudaGetLastError());
    return CLDRD_OK;
}

int cldrd_verify_seed(int device, const float* scores_dev, int64_t nq, int32_t k, const float* seed_dev,
                      const float* eps2_dev, int32_t* fail_dev, void* cuda_stream) {
    if (!scores_dev || !seed_dev || !eps2_dev || !fail_dev || nq < 1 || k < 1)
        return fail(CLDRD_EINVAL, "verify_seed: bad argument");
    DeviceGuard g(device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CU_TRY(cudaMemsetAsync(fail_dev, 0, size_t(nq) * sizeof(int32_t), st));
    verify_seed_kernel<<<unsigned((nq + 255) / 256), 256, 0, st>>>(scores_dev, int(nq), k, seed_dev, eps2_dev, fail_dev,
                                                                    nullptr);
    CU_TRY(cudaGetLastError());
    return CLDRD_OK;
}

int cldrd_shard_norm_bound(const cldrd_shard* s, float* out) {
    if (!s || !out) return fail(CLDRD_EINVAL, "norm_bound: NULL");
    *out = s->bmax_norm;
    return CLDRD_OK;
}

int cldrd_shard_set_norm_bound(cldrd_shard* s, float bound) {
    if (!s || !(bound >= 0.f)) return fail(CLDRD_EINVAL, "set_norm_bound: bad argument");
    if (bound < s->bmax_norm) return fail(CLDRD_EINVAL, "set_norm_bound: below this shard's own bound");
    s->bmax_norm = bound;
    return CLDRD_OK;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int cldrd_search_host(cldrd_shard* s, const float* q_host, int64_t nq, int32_t k, float* out_scores_host,
                      int64_t* out_ids_host) {
    if (!s || nq < 0 || (nq && (!q_host || !out_scores_host || !out_ids_host)))
        return fail(CLDRD_EINVAL, "search_host: NULL argument");
    if (k < 1 || k > CLDRD_MAX_K) return fail(CLDRD_EINVAL, "search: k=%d outside [1,%d]", k, CLDRD_MAX_K);
    if (nq == 0) return CLDRD_OK;
    DeviceGuard g(s->device);
    const size_t qbytes = size_t(nq) * s->d * sizeof(float);
    const size_t oelems = size_t(nq) * k;
    // outputs that already live in pinned memory (cldrd_host_alloc) are written by the DMA engine
    // directly; pageable ones go through a pinned staging buffer
    const bool direct_out = is_pinned_host(out_scores_host) && is_pinned_host(out_ids_host);
    // ... and when the GPU can address them (cldrd_host_alloc maps them), the re-score kernel stores the rows there
    // itself: the 12 bytes per hit cross PCIe while the kernel is still gathering, instead of in a copy behind it
    float* map_D = nullptr;
    int64_t* map_I = nullptr;
    static const bool host_direct = [] {
        const char* e = getenv("CLDRD_HOST_DIRECT");
        return !e || atoi(e) != 0;
    }();
    if (direct_out && host_direct) {
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&map_D), out_scores_host, 0) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void**>(&map_I), out_ids_host, 0) != cudaSuccess) {
            cudaGetLastError();
            map_D = nullptr;
            map_I = nullptr;
        }
    }
    const bool stores_direct = map_D && map_I;
    if (qbytes > s->d_q_bytes) {
        cudaFree(s->d_q);
        s->d_q = nullptr;
        s->d_q_bytes = 0;
        CU_TRY(cudaMalloc(&s->d_q, qbytes));
        s->d_q_bytes = qbytes;
    }
    if (!direct_out && oelems > s->h_out_elems) {
        if (s->h_D) cudaFreeHost(s->h_D);
        if (s->h_I) cudaFreeHost(s->h_I);
        s->h_D = nullptr;
        s->h_I = nullptr;
        s->h_out_elems = 0;
        CU_TRY(cudaHostAlloc(&s->h_D, oelems * sizeof(float), cudaHostAllocDefault));
        CU_TRY(cudaHostAlloc(&s->h_I, oelems * sizeof(int64_t), cudaHostAllocDefault));
        s->h_out_elems = oelems;
    }
    if (!stores_direct && oelems > s->d_out_elems) {
        cudaFree(s->d_D);
        cudaFree(s->d_I);
        s->d_D = nullptr;
        s->d_I = nullptr;
        s->d_out_elems = 0;
        CU_TRY(cudaMalloc(&s->d_D, oelems * sizeof(float)));
        CU_TRY(cudaMalloc(&s->d_I, oelems * sizeof(int64_t)));
        s->d_out_elems = oelems;
    }
    cudaStream_t st = cudaStreamPerThread;
    // pageable or pinned: the runtime stages pageable sources itself
    CU_TRY(cudaMemcpyAsync(s->d_q, q_host, qbytes, cudaMemcpyHostToDevice, st));
    if (stores_direct) {
        int rc = cldrd_search_dev(s, s->d_q, nq, k, 1, map_D, map_I, st);   // synchronises the stream before it returns
        return rc;
    }
    int rc = cldrd_search_dev(s, s->d_q, nq, k, 1, s->d_D, s->d_I, st);
    if (rc) return rc;
    float* dst_D = direct_out ? out_scores_host : s->h_D;
    int64_t* dst_I = direct_out ? out_ids_host : s->h_I;
    CU_TRY(cudaMemcpyAsync(dst_D, s->d_D, oelems * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(dst_I, s->d_I, oelems * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (!direct_out) {
        memcpy(out_scores_host, s->h_D, oelems * sizeof(float));
        memcpy(out_ids_host, s->h_I, oelems * sizeof(int64_t));
    }
    return CLDRD_OK;
}

int cldrd_host_alloc(void** out, int64_t nbytes) {
    if (!out || nbytes < 0) return fail(CLDRD_EINVAL, "host_alloc: bad argument");
    *out = nullptr;
    if (nbytes == 0) return CLDRD_OK;
    cudaError_t e = cudaHostAlloc(out, size_t(nbytes), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CLDRD_ENOMEM : CLDRD_ECUDA, "cudaHostAlloc(%lld) failed: %s",
                    (long long)nbytes, cudaGetErrorString(e));
    }
    return CLDRD_OK;
}

void cldrd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int cldrd_host_register(void* p, int64_t nbytes) {
    if (!p || nbytes < 1) return fail(CLDRD_EINVAL, "host_register: bad argument");
    cudaError_t e = cudaHostRegister(p, size_t(nbytes), cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CLDRD_ECUDA, "cudaHostRegister(%lld) failed: %s", (long long)nbytes, cudaGetErrorString(e));
    }
    return CLDRD_OK;
}

int cldrd_host_device_ptr(int device, void* host_ptr, void** out_dev_ptr) {
    if (!host_ptr || !out_dev_ptr) return fail(CLDRD_EINVAL, "host_device_ptr: NULL");
    DeviceGuard g(device);
    *out_dev_ptr = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(out_dev_ptr, host_ptr, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CLDRD_ECUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e));
    }
    return CLDRD_OK;
}

int cldrd_host_unregister(void* p) {
    if (!p) return CLDRD_OK;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CLDRD_ECUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
    }
    return CLDRD_OK;
}

int cldrd_merge_planes(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts, int64_t plane_rows,
                       int64_t nq, int32_t w, int32_t k, const int64_t* id_map_dev, float* out_scores_dev,
                       int64_t* out_ids_dev, void* cuda_stream) {
    if (parts < 1 || nq < 0 || nq > plane_rows || w < 1 || k < 1 || k > CLDRD_MAX_K ||
        (nq && (!scores_dev || !rows_dev || !out_scores_dev || !out_ids_dev)))
        return fail(CLDRD_EINVAL, "merge: bad argument");
    if (nq == 0) return CLDRD_OK;
    int k_pad = 2;
    while (k_pad < k) k_pad <<= 1;
    const size_t smem = (size_t(parts) * w + size_t(k_pad)) * 8;
    if (smem > 200 * 1024) return fail(CLDRD_EINVAL, "merge: parts*w=%d too large", parts * w);
    DeviceGuard g(device);
    CU_TRY(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    merge_kernel<<<unsigned(nq), 512, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
        scores_dev, rows_dev, parts, plane_rows, w, k, k_pad, id_map_dev, out_scores_dev, out_ids_dev);
    CU_TRY(cudaGetLastError());
    return CLDRD_OK;
}

int cldrd_merge_w(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts, int64_t nq, int32_t w,
                  int32_t k, const int64_t* id_map_dev, float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    return cldrd_merge_planes(device, scores_dev, rows_dev, parts, nq, nq, w, k, id_map_dev, out_scores_dev, out_ids_dev,
                              cuda_stream);
}

int cldrd_merge(int device, const float* scores_dev, const int64_t* rows_dev, int32_t parts, int64_t nq, int32_t k,
                const int64_t* id_map_dev, float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    return cldrd_merge_w(device, scores_dev, rows_dev, parts, nq, k, k, id_map_dev, out_scores_dev, out_ids_dev,
                         cuda_stream);
}

int cldrd_shard_set_profiling(cldrd_shard* s, int32_t on) {
    if (!s) return fail(CLDRD_EINVAL, "set_profiling: NULL");
    s->profile = on != 0;
    return CLDRD_OK;
}

int cldrd_shard_last_scan_time(const cldrd_shard* s, double* scan_ms, int64_t* scan_launches) {
    if (!s) return fail(CLDRD_EINVAL, "last_scan_time: NULL");
    if (scan_ms) *scan_ms = s->scan_ms;
    if (scan_launches) *scan_launches = s->scan_launches;
    return CLDRD_OK;
}

int cldrd_shard_last_scan_launches(const cldrd_shard* s, double* ms, int64_t* rows, int32_t cap) {
    if (!s || !ms || !rows) return fail(CLDRD_EINVAL, "last_scan_launches: NULL");
    const int n = int(std::min<size_t>(std::min(s->ev_ms.size(), s->ev_rows.size()), size_t(std::max(cap, 0))));
    for (int i = 0; i < n; ++i) {
        ms[i] = s->ev_ms[i];
        rows[i] = s->ev_rows[i];
    }
    return n;
}

int cldrd_shard_wait_cycles(cldrd_shard* s, uint64_t out[4], int32_t reset) {
    if (!s || !out) return fail(CLDRD_EINVAL, "wait_cycles: NULL");
    DeviceGuard g(s->device);
    if (!s->w_wait) return fail(CLDRD_ESTATE, "wait_cycles: shard not finalized");
    CU_TRY(cudaMemcpy(out, s->w_wait, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) CU_TRY(cudaMemset(s->w_wait, 0, 4 * sizeof(unsigned long long)));
    return CLDRD_OK;
}

int cldrd_shard_last_stats(const cldrd_shard* s, int64_t stats[8]) {
    if (!s || !stats) return fail(CLDRD_EINVAL, "last_stats: NULL");
    memcpy(stats, s->stats, sizeof(s->stats));
    return CLDRD_OK;
}

int cldrd_scan_dense_dev(cldrd_shard* s, const float* q_dev, int64_t nq, int64_t row_begin, int64_t nrows,
                         float* out_dev, void* cuda_stream) {
    if (!s || !q_dev || !out_dev || nq < 1 || nq > kQueryBatch || nrows < 1 || nrows > kDensePiece || row_begin < 0 ||
        row_begin + nrows > s->nrows)
        return fail(CLDRD_EINVAL, "scan_dense: bad argument");
    if (!s->finalized) return fail(CLDRD_ESTATE, "scan_dense: shard not finalized");
    DeviceGuard g(s->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    BatchCtx c{};
    c.s = s;
    c.st = st;
    c.q = q_dev;
    c.nq = int(nq);
    c.k = 1;
    CU_TRY(cudaMemsetAsync(s->w_stats, 0, ST_COUNT * sizeof(unsigned long long), st));
    int rc = launch_prep(c);
    if (rc) return rc;
    if ((rc = launch_scan(c, TC_DENSE, row_begin, int(nrows)))) return rc;
    CU_TRY(cudaMemcpy2DAsync(out_dev, size_t(nrows) * 4, s->w_dense, size_t(kDensePiece) * 4, size_t(nrows) * 4, size_t(nq),
                             cudaMemcpyDeviceToDevice, st));
    return read_stats(s, st);
}

}  // extern "C"
