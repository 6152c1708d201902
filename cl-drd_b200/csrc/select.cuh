// Candidate bookkeeping kernels: per-query streaming select, exact fp32 re-score + final sort,
// multi-shard merge, and the query / index preparation passes.
//
// Streaming invariant.  For query q let eps be the proven bound on
// |scan score - exact fp32 score|.  After every chunk the list holds every row seen so far whose
// scan score is >= v_k - 2*eps, where v_k is the k-th best scan score seen so far, and
// thr[q] = v_k - 2*eps is what the fused filter of the next chunk compares against.  Any row of
// the exact top-k satisfies that inequality, so the final exact re-score of the list returns
// the exact result.
#pragma once
#include <cfloat>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cldrd.h"
#include "device_common.cuh"

namespace cldrd {

struct PeerPtrs {
    void* p[CLDRD_MAX_PEERS];   // the same buffer in every rank's exchange block, as mapped into this process
};

// ------------------------------------------------------------------------------------------
// block-wide radix select of the k-th largest value among keys[0..n) (shared memory)
// BITS = 32: on the score half only.  BITS = 64: on the full key (unique -> exactly k kept).
// Returns the selected value (same in all threads).
// ------------------------------------------------------------------------------------------
template <int BITS>
__device__ uint64_t block_radix_select(const uint64_t* keys, int n, int k, uint32_t* hist /*256*/,
                                       uint64_t* bcast /*2*/) {
    uint64_t prefix = 0, mask = 0;
    int remaining = k;
    constexpr int kLow = (BITS == 32) ? 32 : 0;  // 32: only the score half takes part
    for (int shift = 56; shift >= kLow; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            uint64_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[uint32_t(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0, b = 255;
            for (; b > 0; --b) {
                int h = int(hist[b]);
                if (cum + h >= remaining) break;
                cum += h;
            }
            bcast[0] = uint64_t(b);
            bcast[1] = uint64_t(remaining - cum);
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= uint64_t(255) << shift;
        remaining = int(bcast[1]);
        __syncthreads();
    }
    return prefix;
}

struct SelectParams {
    uint64_t* list;        // [nq][keep_cap]
    int* list_len;         // [nq]
    int keep_cap;
    uint64_t* surv;        // [nq][q_stride]   (filter mode): `groups` segments of seg_cap entries
    int* seg_cnt;          // [nq][groups]     entries written per segment (may exceed seg_cap)
    int q_stride;
    int seg_cap;
    int pool_cap;          // shared overflow pool behind the segments (slot `groups`)
    int groups;            // <= 511
    int surv_cap;          // survivors one CTA can take in (shared-memory space)
    const float* dense;    // [nq][dense_ld]   (dense mode, else nullptr)
    int dense_ld;
    int dense_n;           // valid columns
    uint32_t dense_row0;   // local row of dense column 0
    int dense_tile_stride; // sampled scan: column j came from row dense_row0 + (j/256)*stride*256 + j%256
    float* thr;            // [nq]
    const float* seed;     // [nq]  seed threshold (-inf when unseeded): thr never drops below it
    const float* band;     // [nq]  2*eps
    int k;
    int* fail;             // [nq]
    unsigned long long* stats;
    // exact compaction (tie floods)
    const float* xb;       // fp32 rows of the shard
    const float* q;        // [nq][d]
    int d;
    int vec4;
};

// One CTA per query.  dyn smem: keys[keep_cap + surv_cap] u64 | q_s[d] f32 (16B aligned)
__global__ void __launch_bounds__(512) select_merge_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    float* q_s = reinterpret_cast<float*>(sm_raw + size_t(p.keep_cap + p.surv_cap) * 8);
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[2];
    __shared__ int s_n, s_out;
    __shared__ int s_warp[32];
    __shared__ int s_off[512];

    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    if (p.fail[q]) return;
    const int L = p.list_len[q];
    uint64_t* my_list = p.list + size_t(q) * p.keep_cap;
    for (int i = tid; i < L; i += blockDim.x) keys[i] = my_list[i];
    if (tid == 0) {
        s_n = L;
        s_out = 0;
    }
    __syncthreads();
    int n;
    if (p.dense) {
        const float t = p.thr[q];
        const float* row = p.dense + size_t(q) * p.dense_ld;
        for (int j = tid; j < p.dense_n; j += blockDim.x) {
            float v = row[j];
            if (v >= t) {
                int slot = atomicAdd(&s_n, 1);
                keys[slot] = make_key(v, p.dense_row0 + uint32_t(j >> 8) * uint32_t(p.dense_tile_stride * 256) + uint32_t(j & 255));
            }
        }
        __syncthreads();
        n = s_n;
    } else {
        // gather this query's survivor segments: exclusive scan of the segment counts
        // (slot G = the shared overflow pool; a segment's count above seg_cap means the excess is there)
        const int G = p.groups + 1;
        int c = 0;
        bool over = false;
        if (tid < G) {
            c = p.seg_cnt[size_t(q) * G + tid];
            p.seg_cnt[size_t(q) * G + tid] = 0;   // ready for the next chunk
            if (tid < G - 1) c = min(c, p.seg_cap);
            else over = c > p.pool_cap;
        }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        const int any_over = __syncthreads_or(over ? 1 : 0);
        if (tid < 32) {
            int w = (tid < int(blockDim.x >> 5)) ? s_warp[tid] : 0;
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (tid >= o) wi += t;
            }
            s_warp[tid] = wi - w;                 // exclusive warp offsets
            if (tid == 31) s_n = wi;              // total survivors
        }
        __syncthreads();
        const int S = s_n;
        if (any_over || S > p.surv_cap) {  // survivors were dropped: this query goes to the dense fallback
            if (tid == 0) {
                p.fail[q] = 1;
                atomicAdd(&p.stats[ST_FAILED], 1ull);
            }
            return;
        }
        if (tid < G) s_off[tid] = s_warp[tid >> 5] + incl - c;
        __syncthreads();
        const uint64_t* sv = p.surv + size_t(q) * p.q_stride;
        const int warp_id = tid >> 5, lane_id = tid & 31, nw = blockDim.x >> 5;
        for (int g = warp_id; g < G; g += nw) {
            const int off = s_off[g];
            const int cg = ((g + 1 < G) ? s_off[g + 1] : S) - off;
            const uint64_t* src = sv + size_t(g) * p.seg_cap;
            for (int i = lane_id; i < cg; i += 32) keys[L + off + i] = src[i];
        }
        __syncthreads();
        if (tid == 0) atomicAdd(&p.stats[ST_SURVIVORS], (unsigned long long)S);
        n = L + S;
    }
    if (n == L) return;  // nothing new: list and threshold stay as they are
    if (n < p.k) {  // fewer than k rows seen: keep everything, threshold stays -inf
        for (int i = L + tid; i < n; i += blockDim.x) my_list[i] = keys[i];
        if (tid == 0) {
            p.list_len[q] = n;
            atomicMax(&p.stats[ST_MAX_LIST], (unsigned long long)n);
        }
        return;
    }
    // k-th best scan score, then the band below it
    const uint32_t vk = uint32_t(block_radix_select<32>(keys, n, p.k, hist, bcast) >> 32);
    // the list only ever held rows >= seed, so the cut cannot reach below it
    const float cutf = fmaxf(ord2f(vk) - p.band[q], p.seed[q]);
    const uint32_t cut = f2ord(cutf);
    // count the band first: if it does not fit, trim by exact score instead
    int cnt = 0;
    for (int i = tid; i < n; i += blockDim.x) cnt += (key_ord(keys[i]) >= cut);
    if (tid == 0) s_out = 0;
    __syncthreads();
    if (cnt) atomicAdd(&s_out, cnt);
    __syncthreads();
    cnt = s_out;
    __syncthreads();
    if (cnt <= p.keep_cap) {
        if (tid == 0) s_out = 0;
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            uint64_t key = keys[i];
            if (key_ord(key) >= cut) {
                int slot = atomicAdd(&s_out, 1);
                my_list[slot] = key;
            }
        }
        if (tid == 0) {
            p.list_len[q] = cnt;
            p.thr[q] = cutf;
            atomicMax(&p.stats[ST_MAX_LIST], (unsigned long long)cnt);
        }
        return;
    }
    // Tie flood: more than keep_cap rows inside the band.  Re-score every candidate exactly
    // (same routine as the final re-score) and keep exactly the k best by full key; exact
    // scores have zero error, so the invariant still holds with the same band.
    for (int i = tid; i < p.d; i += blockDim.x) q_s[i] = p.q[size_t(q) * p.d + i];
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int i = warp; i < n; i += nwarps) {
        const uint32_t row = key_row(keys[i]);
        float s = exact_dot_warp(q_s, p.xb + size_t(row) * p.d, p.d, p.vec4 != 0, lane);
        if (lane == 0) keys[i] = make_key(s, row);
    }
    __syncthreads();
    const uint64_t kth = block_radix_select<64>(keys, n, p.k, hist, bcast);
    if (tid == 0) s_out = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        uint64_t key = keys[i];
        if (key >= kth) {
            int slot = atomicAdd(&s_out, 1);
            my_list[slot] = key;
        }
    }
    if (tid == 0) {
        p.list_len[q] = p.k;
        p.thr[q] = fmaxf(ord2f(uint32_t(kth >> 32)) - p.band[q], p.seed[q]);
        atomicAdd(&p.stats[ST_EXACT_COMPACT], 1ull);
        atomicAdd(&p.stats[ST_RESCORED], (unsigned long long)n);
        atomicMax(&p.stats[ST_MAX_LIST], (unsigned long long)p.k);
    }
}

// ------------------------------------------------------------------------------------------
// block-wide bitonic sort, descending, n_pad a power of two (shared memory)
// ------------------------------------------------------------------------------------------
__device__ void block_bitonic_desc(uint64_t* keys, int n_pad) {
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                int lo = ((t / stride) * (stride << 1)) + (t % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

struct RescoreParams {
    const float* xb;      // fp32 rows of the shard
    const float* q;       // [nq][d]
    int d;
    int vec4;
    const uint64_t* list; // [nq][keep_cap]
    const int* list_len;
    int keep_cap;
    int n_pad;            // pow2 >= keep_cap
    int k;
    int64_t row0;         // global row of local row 0
    const int64_t* ids;   // external ids of this shard's rows, or nullptr
    float* out_scores;    // [*][k]
    int64_t* out_ids;     // [*][k]
    const int* out_index; // optional: output row of query q (dense fallback scatter), or nullptr
    const int* fail;      // skip failed queries (nullptr = none)
    int* fail_set;        // a list longer than n_pad (reduced-shared-memory launch) flags the query here
    unsigned long long* stats;
    // optional [nq]: list entries whose SCAN score is below cut[q] are dropped unscored (sharded search:
    // the shards agreed that such rows cannot be in the global top-k)
    const float* cut;
    // ... or computed here (node-wide search): counts planes [cnt_parts][plane_stride] of int[j] per query that
    // the shards stored into this rank's block; cut = T - 2*eps for the highest level T that at least k rows of the
    // WHOLE index reach in scan score, -inf if no level does.  Why it is safe: k rows with scan score >= T have
    // exact score >= T - eps, so the exact k-th best is >= T - eps, and a row of the top-k scans >= T - 2*eps.
    const int* cnt_planes;
    size_t cnt_plane_stride;
    int cnt_parts;
    const float* levels;  // [nq][lv_j], descending
    int lv_j;
    const float* band;    // [nq] 2*eps
    // scatter mode (sc_world > 0): output row r of this launch is query q_base + r of the batch; its list goes
    // to plane [sc_rank], row Q % sc_slice of the key buffer of rank Q / sc_slice (peer memory over NVLink):
    // u64 keys (exact score, GLOBAL row), best first, 0 = padding
    int sc_world;
    int sc_rank;
    long long sc_slice;
    long long q_base;
    int sc_key_stride;
    int sc_raise_fail;    // a query this shard could not finish is raised in every rank's qfail (seeded batches)
    PeerPtrs sc_keys;
    PeerPtrs sc_len;      // i32 [world][slice]: number of valid keys of each list (only those are stored)
    PeerPtrs sc_qfail;
};

// One CTA per query: exact fp32 score of every listed row, sort, emit the k best.
// dyn smem: keys[n_pad] u64 | q_s[d] f32
__global__ void __launch_bounds__(1024, 1) rescore_sort_kernel(RescoreParams p) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    float* q_s = reinterpret_cast<float*>(sm_raw + size_t(p.n_pad) * 8);
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    int L = p.list_len[q];
    const bool failed = p.fail && p.fail[q];
    if (failed || L > p.n_pad) {
        // L > n_pad: only possible when launched with less shared memory than keep_cap needs
        if (!failed && tid == 0 && p.fail_set) {
            p.fail_set[q] = 1;
            atomicAdd(&p.stats[ST_FAILED], 1ull);
        }
        if (p.sc_world > 0 && p.sc_raise_fail) {
            // the merging rank gets an empty list, and every rank learns that the query has to be searched again
            const long long Q = p.q_base + (long long)(p.out_index ? p.out_index[q] : q);
            const int dest = int(Q / p.sc_slice);
            if (tid == 0)
                static_cast<int*>(p.sc_len.p[dest])[size_t(p.sc_rank) * size_t(p.sc_slice) + size_t(Q - dest * p.sc_slice)] = 0;
            if (tid < p.sc_world) static_cast<int*>(p.sc_qfail.p[tid])[Q] = 1;
        }
        return;
    }
    for (int i = tid; i < p.d; i += blockDim.x) q_s[i] = p.q[size_t(q) * p.d + i];
    // the candidate list comes into shared memory in one coalesced sweep (a warp that fetched its entries one by one
    // from global memory paid a dependent-load latency per entry: with the short lists of a many-shard search that
    // was as long as the row gather itself)
    const uint64_t* my_list = p.list + size_t(q) * p.keep_cap;
    for (int i = tid; i < L; i += blockDim.x) keys[i] = my_list[i];
    constexpr int kRankSortMax = 512;       // up to here one counting pass beats log^2(n) barrier-separated bitonic passes
    const bool rank_sort = L <= kRankSortMax && p.n_pad >= 2 * kRankSortMax;
    int n_pad = 2;
    while (n_pad < L) n_pad <<= 1;
    if (!rank_sort)
        for (int i = L + tid; i < n_pad; i += blockDim.x) keys[i] = 0;  // sorts last
    __shared__ int s_kept;
    __shared__ uint32_t s_cut;
    if (tid == 0) {
        s_kept = 0;
        s_cut = p.cut ? f2ord(p.cut[q]) : 0u;
    }
    if (p.cnt_planes && tid < 32) {   // lv_j <= 64 levels: lane b looks at levels b and b + 32
        int c0 = 0, c1 = 0;
        for (int part = 0; part < p.cnt_parts; ++part) {
            const int* c = p.cnt_planes + size_t(part) * p.cnt_plane_stride + size_t(q) * p.lv_j;
            if (tid < p.lv_j) c0 += c[tid];
            if (tid + 32 < p.lv_j) c1 += c[tid + 32];
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, tid < p.lv_j && c0 >= p.k);
        const unsigned m1 = __ballot_sync(0xffffffffu, tid + 32 < p.lv_j && c1 >= p.k);
        if (tid == 0) {
            const int b = m0 ? __ffs(m0) - 1 : (m1 ? 32 + __ffs(m1) - 1 : -1);
            s_cut = b >= 0 ? f2ord(p.levels[size_t(q) * p.lv_j + b] - p.band[q]) : 0u;
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const uint32_t cut_ord = s_cut;
    int kept = 0;
    for (int i = warp; i < L; i += nwarps) {     // entry i is read and rewritten by the same warp
        const uint64_t cand = keys[i];
        if (key_ord(cand) < cut_ord) {   // warp-uniform
            __syncwarp();
            if (lane == 0) keys[i] = 0;
            continue;
        }
        const uint32_t row = key_row(cand);
        float s = exact_dot_warp(q_s, p.xb + size_t(row) * p.d, p.d, p.vec4 != 0, lane);
        __syncwarp();
        if (lane == 0) keys[i] = make_key(s, row);
        ++kept;
    }
    if (lane == 0 && kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    const int n_all = L;             // entries in shared memory: exact keys, zero where an entry was dropped
    L = s_kept;                      // zero keys order behind every kept one
    if (tid == 0) atomicAdd(&p.stats[ST_RESCORED], (unsigned long long)L);
    const size_t orow = p.out_index ? size_t(p.out_index[q]) : size_t(q);
    uint64_t* ok = nullptr;
    float* os = nullptr;
    int64_t* oi = nullptr;
    if (p.sc_world > 0) {
        const long long Q = p.q_base + (long long)orow;
        const int dest = int(Q / p.sc_slice);
        const size_t at = size_t(p.sc_rank) * size_t(p.sc_slice) + size_t(Q - dest * p.sc_slice);
        ok = static_cast<uint64_t*>(p.sc_keys.p[dest]) + at * size_t(p.sc_key_stride);
        if (tid == 0) static_cast<int*>(p.sc_len.p[dest])[at] = min(L, p.k);   // only the valid prefix travels
    } else {
        os = p.out_scores + orow * p.k;
        oi = p.out_ids + orow * p.k;
    }
    // entry `key` goes to position r of the sorted output (r < k)
    auto emit = [&](int r, uint64_t key) {
        const uint32_t row = key_row(key);
        if (ok) {
            ok[r] = (key & 0xFFFFFFFF00000000ull) | uint64_t(~(uint32_t(p.row0) + row));
        } else {
            os[r] = ord2f(key_ord(key));
            oi[r] = p.ids ? p.ids[row] : (p.row0 + int64_t(row));
        }
    };
    if (rank_sort) {
        // short list: every thread ranks its entry against all others (keys are distinct: one row, one key) and puts
        // it at its rank -- one pass over shared memory instead of ~40 barrier-separated compare-exchange passes.  The
        // ranked order is staged behind the list (keep_cap >= 2 * kRankSortMax) so that the stores go out coalesced.
        uint64_t* sorted = keys + kRankSortMax;
        for (int i = tid; i < n_all; i += blockDim.x) {
            const uint64_t key = keys[i];
            if (key == 0) continue;
            int r = 0;
#pragma unroll 8
            for (int j = 0; j < n_all; ++j) {
                const uint64_t x = keys[j];
                r += (x > key) || (x == key && j < i);
            }
            sorted[r] = key;
        }
        __syncthreads();
        for (int i = tid; i < min(L, p.k); i += blockDim.x) emit(i, sorted[i]);
    } else {
        block_bitonic_desc(keys, n_pad);
        for (int i = tid; i < min(L, p.k); i += blockDim.x) emit(i, keys[i]);
    }
    if (!ok) {
        for (int i = L + tid; i < p.k; i += blockDim.x) {   // padding behind the list
            os[i] = -FLT_MAX;
            oi[i] = -1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Multi-shard merge: [parts][plane_rows][w] (score, global row; -1 = padding) -> [grid][k], same key order
// as the single-shard search, so the result is bit-identical to it.  One CTA per query:
// compact the valid entries, radix-select the k-th key, sort only the k survivors.
// dyn smem: keys[parts*w] u64 | top[k_pad] u64
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) merge_kernel(const float* scores, const int64_t* rows,
                                                    int parts, int64_t plane_rows, int w, int k, int k_pad,
                                                    const int64_t* id_map, float* out_scores,
                                                    int64_t* out_ids) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    uint64_t* top = keys + size_t(parts) * w;
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[2];
    __shared__ int s_n, s_out;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const int n_in = parts * w;
    if (tid == 0) {
        s_n = 0;
        s_out = 0;
    }
    for (int i = tid; i < k_pad; i += blockDim.x) top[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_in; i += blockDim.x) {
        const int part = i / w, j = i - part * w;
        const size_t off = (size_t(part) * plane_rows + q) * w + j;
        const int64_t r = rows[off];
        if (r >= 0) keys[atomicAdd(&s_n, 1)] = make_key(scores[off], uint32_t(r));
    }
    __syncthreads();
    const int n = s_n;
    int m = n;                                  // how many go to the sort
    if (n > k) {
        const uint64_t kth = block_radix_select<64>(keys, n, k, hist, bcast);   // keys are distinct
        for (int i = tid; i < n; i += blockDim.x) {
            const uint64_t key = keys[i];
            if (key >= kth) {
                const int slot = atomicAdd(&s_out, 1);
                if (slot < k) top[slot] = key;   // duplicate keys (malformed input) must not overrun `top`
            }
        }
        m = k;
    } else {
        for (int i = tid; i < n; i += blockDim.x) top[i] = keys[i];
    }
    int n_pad = 2;
    while (n_pad < m) n_pad <<= 1;
    block_bitonic_desc(top, n_pad);
    for (int i = tid; i < k; i += blockDim.x) {
        const uint64_t key = i < m ? top[i] : 0ull;
        if (key != 0) {
            const uint32_t row = key_row(key);
            out_scores[q * k + i] = ord2f(key_ord(key));
            out_ids[q * k + i] = id_map ? id_map[row] : int64_t(row);
        } else {
            out_scores[q * k + i] = -FLT_MAX;
            out_ids[q * k + i] = -1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Query preparation: norms -> error band, reset per-query state, low-precision copy.
// One warp per query.
// ------------------------------------------------------------------------------------------
struct QueryPrepParams {
    const float* q;    // [nq][d]
    int nq, d;
    int lp_kind;       // 0 none, 1 f16, 2 bf16
    void* q_lp;        // [nq][d] half / bf16
    float coef;        // eps = coef * |q| * bmax_norm + abs_coef * (|q| + bmax_norm)
    float abs_coef;
    float bmax_norm;
    float* band;       // 2*eps (slightly inflated)
    float* seed;       // reset to -inf
    float* thr;
    int* list_len;
    int* fail;
    unsigned long long* stats;
};

__global__ void query_prep_kernel(QueryPrepParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.nq) return;
    const float* row = p.q + size_t(warp) * p.d;
    float ss = 0.f, mx = 0.f;
    for (int c = lane; c < p.d; c += 32) {
        float v = row[c];
        ss = fmaf(v, v, ss);
        mx = fmaxf(mx, fabsf(v));
        if (p.lp_kind == 1)
            reinterpret_cast<__half*>(p.q_lp)[size_t(warp) * p.d + c] = __float2half_rn(v);
        else if (p.lp_kind == 2)
            reinterpret_cast<__nv_bfloat16*>(p.q_lp)[size_t(warp) * p.d + c] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        const float qn = sqrtf(ss) * 1.0001f;
        const float eps = p.coef * qn * p.bmax_norm + p.abs_coef * (qn + p.bmax_norm);
        p.band[warp] = 2.0f * eps * 1.0001f;
        p.thr[warp] = -INFINITY;
        p.seed[warp] = -INFINITY;
        p.list_len[warp] = 0;
        p.fail[warp] = 0;
        if (p.lp_kind == 1 && !(mx < 65504.f)) atomicAdd(&p.stats[ST_RANGE_ERR], 1ull);
    }
}

// ------------------------------------------------------------------------------------------
// Index preparation (shard finalize): max row norm, max |x|, optional fp16/bf16 copy.
// One warp per row, grid-stride.
// ------------------------------------------------------------------------------------------
__global__ void index_prep_kernel(const float* xb, int64_t nrows, int d, int lp_kind, void* lp,
                                  unsigned int* max_norm2_bits, unsigned int* max_abs_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    float best_ss = 0.f, best_mx = 0.f;
    for (int64_t r = warp0; r < nrows; r += nwarps) {
        const float* row = xb + r * d;
        float ss = 0.f, mx = 0.f;
        for (int c = lane; c < d; c += 32) {
            float v = row[c];
            ss = fmaf(v, v, ss);
            mx = fmaxf(mx, fabsf(v));
            if (lp_kind == 1)
                reinterpret_cast<__half*>(lp)[r * d + c] = __float2half_rn(v);
            else if (lp_kind == 2)
                reinterpret_cast<__nv_bfloat16*>(lp)[r * d + c] = __float2bfloat16_rn(v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        // NaN/inf rows poison the bound on purpose: (ss != ss) -> +inf
        if (!(ss == ss)) ss = INFINITY;
        if (!(mx == mx)) mx = INFINITY;
        best_ss = fmaxf(best_ss, ss);
        best_mx = fmaxf(best_mx, mx);
    }
    if (lane == 0) {
        atomicMax(max_norm2_bits, __float_as_uint(best_ss));  // non-negative floats order as uints
        atomicMax(max_abs_bits, __float_as_uint(best_mx));
    }
}

// ------------------------------------------------------------------------------------------
// Seeded thresholds.  A sample of the index rows is scanned densely first; the
// j-th best sample score of a query is, in expectation, its (j / sample fraction)-th best score
// overall, so it seeds the filter threshold of the real scan far above -inf.  ANY seed is safe
// because the result is verified afterwards: every row with scan score >= seed was collected, an
// uncollected row has exact score < seed + eps, so if the k-th returned exact score is
// >= seed + eps nothing was missed.  Queries that fail the check are searched again unseeded.
// ------------------------------------------------------------------------------------------

// Best j of the `cols` sample values of each query (raw sample scores, or maxima of 32-column
// groups: the j-th largest group maximum is a lower bound of the j-th largest score, which only
// makes the seed more conservative).  One CTA per query: radix-select the j-th largest, then sort
// the few values at or above it.  Output best first, -inf padded.   j <= 64.
// dyn smem: keys[cols] u64 (orderable float in the high half)
// The [nq][j] result is stored at element offset out_off of every buffer in outs (node-wide search: plane `rank` of
// every rank's sample buffer, peer stores over NVLink; nouts = 1 for a local result).
__global__ void __launch_bounds__(256) sample_topj_kernel(const float* vals, int ld, int cols, int j,
                                                          PeerPtrs outs, int nouts, size_t out_off) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[2];
    __shared__ uint64_t top[64];
    __shared__ int s_cnt;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const float* row = vals + size_t(q) * ld;
    for (int i = tid; i < cols; i += blockDim.x) {
        float v = row[i];
        if (!(v == v)) v = -INFINITY;                       // NaN never seeds a threshold
        keys[i] = uint64_t(f2ord(v)) << 32;
    }
    if (tid < 64) top[tid] = uint64_t(f2ord(-INFINITY)) << 32;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (cols > 0) {
        const int jj = min(j, cols);
        const uint64_t kth = block_radix_select<32>(keys, cols, jj, hist, bcast);   // high half = j-th largest
        for (int i = tid; i < cols; i += blockDim.x) {
            if (keys[i] >= kth) {
                int slot = atomicAdd(&s_cnt, 1);
                if (slot < 64) top[slot] = keys[i];          // ties beyond 64 slots equal kth anyway
            }
        }
        __syncthreads();
    }
    // 64-element bitonic sort, descending
    for (int size = 2; size <= 64; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            if (tid < 32) {
                int lo = ((tid / stride) * (stride << 1)) + (tid % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = top[lo], b = top[hi];
                if ((a < b) == desc) {
                    top[lo] = b;
                    top[hi] = a;
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < j; i += blockDim.x) {
        const float v = ord2f(uint32_t(top[i] >> 32));
        for (int o = 0; o < nouts; ++o) (static_cast<float*>(outs.p[o]) + out_off)[size_t(q) * j + i] = v;
    }
}

// seed[q] = j-th best of the parts*j sample scores gathered from all shards ([parts][nq][j]).
// One warp per query (parts*j is small).  Fewer than j finite scores -> -inf (no seed).
__global__ void seed_from_samples_kernel(const float* topj, int parts, int nq, int j, int rank_j,
                                         float* seed) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const int n = parts * j;
    // rank_j-th largest by counting: value v is the answer if #(x > v) < rank_j <= #(x >= v)
    float best = -INFINITY;
    for (int c = lane; c < n; c += 32) {
        const int part = c / j, i = c - part * j;
        const float v = topj[(size_t(part) * nq + q) * j + i];
        int gt = 0, ge = 0;
        for (int t = 0; t < n; ++t) {
            const int pt = t / j, it = t - pt * j;
            const float x = topj[(size_t(pt) * nq + q) * j + it];
            gt += x > v;
            ge += x >= v;
        }
        if (gt < rank_j && rank_j <= ge) best = fmaxf(best, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) seed[q] = best;
}

// After the exact re-score: fail[q] = 1 unless the k-th returned score clears seed + eps.
__global__ void verify_seed_kernel(const float* scores /*[nq][k]*/, int nq, int k, const float* seed,
                                   const float* band /*2*eps*/, int* fail, unsigned long long* stats) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float s0 = seed[q];
    if (s0 == -INFINITY) return;                       // unseeded query: nothing to verify
    const float kth = scores[size_t(q) * k + (k - 1)];  // -FLT_MAX when fewer than k rows came back
    if (!(kth >= s0 + 0.5f * band[q] * 1.0001f)) {
        if (!fail[q]) {
            fail[q] = 1;
            if (stats) atomicAdd(&stats[ST_FAILED], 1ull);
        }
    }
}

// seed -> thr (start of a seeded pass) ; list emptied
__global__ void apply_seed_kernel(const float* seed_in, int nq, float bias, float* seed, float* thr, int* list_len) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float s0 = seed_in[q] + bias;   // bias: test hook (CLDRD_SEED_BIAS) to force seed misses
    seed[q] = s0;
    thr[q] = s0;
    list_len[q] = 0;
}

// gather failed queries into a compact matrix + remember where their results go
__global__ void gather_failed_kernel(const float* q, int d, const int* fail_index, int nfail,
                                     float* q_out) {
    const int f = blockIdx.x;
    if (f >= nfail) return;
    const float* src = q + size_t(fail_index[f]) * d;
    float* dst = q_out + size_t(f) * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) dst[c] = src[c];
}

}  // namespace cldrd
