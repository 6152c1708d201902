// Kernels of the node-wide sharded search (one shard per GPU of one NVLink / NVSwitch node): every exchange
// between the shards is a store into the peer's HBM by the kernel that produced the data, every barrier a flag
// kernel over the same peer memory.  No NCCL call, no host synchronisation inside a batch.
//
// Replaces what the reference's dead `co.shard = True` branch (retriever/retrieval_utils.py:174-182) meant faiss'
// IndexShards to do on host threads.
#pragma once
#include "select.cuh"

namespace cldrd {

// Exchange block owned by every rank (identical layout everywhere; peers map it with CUDA IPC or, inside one
// process, use the pointer directly).  Byte offsets from the block base.
struct NodeLayout {
    size_t flags = 0;    // u32 [CLDRD_MAX_PEERS]        barrier epochs, slot p written by rank p
    size_t ctrl = 0;     // i32 [16]                     [0]: output set chosen by the publishing rank for this batch
    size_t qfail = 0;    // i32 [QB]                     query must be searched again (any rank may raise it)
    size_t topj = 0;     // f32 [world][QB][J]           sample scores, plane p written by rank p
    size_t counts = 0;   // i32 [world][QB][J]           candidates above each level, plane p written by rank p
    size_t xkeys = 0;    // u64 [world][slice][cap_k]    re-scored lists of the queries this rank merges
    size_t xlen = 0;     // i32 [world][slice]           how many keys of each list are valid (the rest is stale)
    size_t res_d = 0;    // f32 [QB][cap_k]              merged scores (used on the rank that collects the result)
    size_t res_i = 0;    // i64 [QB][cap_k]              merged ids
    size_t qx = 0;       // f32 [QB][d]                  the batch's queries: every rank uploads 1/G and stores it everywhere
    size_t total = 0;
    int64_t slice = 0;   // rows per plane of xkeys = ceil(QB / world)
};

__host__ inline NodeLayout node_layout(int world, int cap_k, int d = 0) {
    NodeLayout l;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t QB = CLDRD_QUERY_BATCH, J = CLDRD_SEED_J;
    l.slice = int64_t((QB + size_t(world) - 1) / size_t(world));
    size_t o = 0;
    l.flags = o;  o = up(o + CLDRD_MAX_PEERS * 4);
    l.ctrl = o;   o = up(o + 16 * 4);
    l.qfail = o;  o = up(o + QB * 4);
    l.topj = o;   o = up(o + size_t(world) * QB * J * 4);
    l.counts = o; o = up(o + size_t(world) * QB * J * 4);
    l.xkeys = o;  o = up(o + size_t(world) * size_t(l.slice) * size_t(cap_k) * 8);
    l.xlen = o;   o = up(o + size_t(world) * size_t(l.slice) * 4);
    l.res_d = o;  o = up(o + QB * size_t(cap_k) * 4);
    l.res_i = o;  o = up(o + QB * size_t(cap_k) * 8);
    l.qx = o;     o = up(o + QB * size_t(d) * 4);
    l.total = o;
    return l;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Barrier over the ranks of the node, stream-ordered: everything this rank's stream did before it (peer stores
// included: kernel completion makes them visible system-wide) happens before anything a peer's stream does after
// its own barrier with the same epoch.  Lane p publishes `epoch` in rank p's flag slot [rank] and waits for rank
// p's epoch in the own slot [p].  Epochs only grow, so a late reader never misses one.  A peer that never
// arrives (crashed process, diverged call sequence) ends in an error code after `timeout_ns`, not in a hung GPU.
struct BarrierParams {
    uint32_t* peer_flags[CLDRD_MAX_PEERS];
    const uint32_t* my_flags;
    int world, rank;
    uint32_t epoch;
    unsigned long long timeout_ns;
    unsigned long long* err;      // stats[ST_KERNEL_ERR]
};
constexpr unsigned long long kErrBarrierTimeout = 900;

__global__ void node_barrier_kernel(BarrierParams b) {
    const int p = threadIdx.x;
    if (p >= b.world) return;
    __threadfence_system();
    st_release_sys(b.peer_flags[p] + b.rank, b.epoch);
    const unsigned long long t0 = global_ns();
    while (int32_t(ld_acquire_sys(b.my_flags + p) - b.epoch) < 0) {
        if (global_ns() - t0 > b.timeout_ns) {
            atomicMax(b.err, kErrBarrierTimeout + (unsigned long long)p);
            break;
        }
        __nanosleep(64);
    }
}

// levels[q][0..j) = the j best of the parts*j sample scores all shards stored into this rank's block, best first;
// the last one (+ bias) seeds the filter threshold.  One CTA per query; dyn smem parts*j floats.
__global__ void levels_seed_kernel(const float* topj, size_t plane_stride, int parts, int j, float bias, float* levels,
                                   float* seed, float* thr, int* list_len) {
    extern __shared__ float lv_s[];
    const int q = blockIdx.x;
    const int n = parts * j;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const int part = c / j, i = c - part * j;
        float v = topj[size_t(part) * plane_stride + size_t(q) * j + i];
        if (!(v == v)) v = -INFINITY;
        lv_s[c] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const float v = lv_s[c];
        int r = 0;
        for (int t = 0; t < n; ++t) {
            const float x = lv_s[t];
            r += (x > v) || (x == v && t < c);
        }
        if (r < j) levels[size_t(q) * j + r] = v;
        if (r == j - 1) {
            seed[q] = v + bias;
            thr[q] = v + bias;
            list_len[q] = 0;
        }
    }
}

// counts[q][b] = how many entries of this shard's candidate list scan at or above levels[q][b] (levels descending),
// stored into plane `rank` of every rank's counts buffer.  Failed queries count nothing: counts only have to be
// lower bounds.
__global__ void count_levels_peers_kernel(const uint64_t* list, const int* list_len, int keep_cap, const int* fail,
                                          const float* levels, int j, PeerPtrs outs, int nouts, size_t plane_off) {
    __shared__ int hist[64];
    __shared__ float lv[64];
    const int q = blockIdx.x;
    if (threadIdx.x < j) {
        hist[threadIdx.x] = 0;
        lv[threadIdx.x] = levels[size_t(q) * j + threadIdx.x];
    }
    __syncthreads();
    const int L = (fail && fail[q]) ? 0 : list_len[q];
    const uint64_t* my = list + size_t(q) * keep_cap;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float sc = ord2f(key_ord(my[i]));
        int lo = 0, hi = j;            // smallest b with lv[b] <= sc
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (lv[mid] <= sc) hi = mid;
            else lo = mid + 1;
        }
        if (lo < j) atomicAdd(&hist[lo], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // inclusive prefix over the (<= 64) levels, then one store per level and peer
        int a = threadIdx.x < j ? hist[threadIdx.x] : 0;
        int b2 = threadIdx.x + 32 < j ? hist[threadIdx.x + 32] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, a, o);
            const int u = __shfl_up_sync(0xffffffffu, b2, o);
            if (int(threadIdx.x) >= o) {
                a += t;
                b2 += u;
            }
        }
        b2 += __shfl_sync(0xffffffffu, a, 31);
        for (int p = 0; p < nouts; ++p) {
            int* out = static_cast<int*>(outs.p[p]) + plane_off + size_t(q) * j;
            if (int(threadIdx.x) < j) out[threadIdx.x] = a;
            if (int(threadIdx.x) + 32 < j) out[threadIdx.x + 32] = b2;
        }
    }
}

// Merge of the lists the shards scattered into this rank's block, for the slice of queries this rank owns: same
// u64 key order as the single-shard search, so the result is bit-identical to it; the seed check of the seeded
// search (k-th merged score >= seed + eps, else the query is raised in every rank's qfail) and the id_map gather are
// fused in; the rows go wherever the caller's pointers lead (own HBM, the collecting rank's HBM over NVLink, or
// page-locked host memory over PCIe).  One CTA per query.  dyn smem: keys[parts*k] u64 | top[k] u64
struct MergeKeysParams {
    const uint64_t* xkeys;     // [parts][slice][key_stride]
    const int* xlen;           // [parts][slice] valid keys per list
    int parts;
    long long slice;
    int key_stride;
    int k, k_pad;
    long long q_lo;            // batch index of the first query of this rank's slice
    const float* seed;         // [nq] of the batch (-inf = unseeded: nothing to verify)
    const float* band;         // [nq] 2*eps
    const long long* id_map;   // optional global row -> external id
    float* out_scores;         // [*][k]
    long long* out_ids;        // [*][k]
    // ... or, when out_sets > 0, one of several registered output sets, chosen per batch by ONE rank and published in
    // every rank's ctrl word before the barrier that precedes this kernel (rows start at out_row0 of the set)
    int out_sets;
    const int* ctrl;
    long long out_row0;
    float* set_scores[CLDRD_MAX_OUT_SETS];
    long long* set_ids[CLDRD_MAX_OUT_SETS];
    const int* out_rows;       // optional: output row of batch query Q (default Q)
    PeerPtrs qfail;            // i32 [QB] in every rank's block
    int world;
};

__global__ void __launch_bounds__(512) merge_keys_kernel(MergeKeysParams p) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    uint64_t* top = keys + size_t(p.parts) * p.k;      // [k] the merged order, staged so that the stores go out coalesced
    __shared__ int s_off[CLDRD_MAX_PEERS + 1];
    const long long r = blockIdx.x;
    const long long Q = p.q_lo + r;
    const int tid = threadIdx.x;
    // only the valid prefix of every shard's list is read (a shard of a G-GPU search sends about k / G keys)
    if (tid == 0) {
        int o = 0;
        for (int part = 0; part < p.parts; ++part) {
            s_off[part] = o;
            o += min(max(p.xlen[size_t(part) * size_t(p.slice) + size_t(r)], 0), p.k);
        }
        s_off[p.parts] = o;
    }
    __syncthreads();
    const int n = s_off[p.parts];
    for (int part = 0; part < p.parts; ++part) {
        const int lo = s_off[part], len = s_off[part + 1] - lo;
        const uint64_t* src = p.xkeys + (size_t(part) * size_t(p.slice) + size_t(r)) * size_t(p.key_stride);
        for (int j = tid; j < len; j += blockDim.x) keys[lo + j] = src[j];
    }
    __syncthreads();
    size_t orow = p.out_rows ? size_t(p.out_rows[Q]) : size_t(Q);
    float* os = p.out_scores;
    long long* oi = p.out_ids;
    if (p.out_sets > 0) {
        const int sel = min(max(*p.ctrl, 0), p.out_sets - 1);
        os = p.set_scores[sel];
        oi = p.set_ids[sel];
        orow += size_t(p.out_row0);
    }
    os += orow * p.k;
    oi += orow * p.k;
    const float s0 = p.seed ? p.seed[Q] : -INFINITY;
    auto check_seed = [&](float kth) {     // k-th merged score must clear seed + eps, else the query is searched again
        if (s0 != -INFINITY && !(kth >= s0 + 0.5f * p.band[Q] * 1.0001f))
            for (int w = 0; w < p.world; ++w) static_cast<int*>(p.qfail.p[w])[Q] = 1;
    };
    // Every list arrives sorted (best first), so a key's place in the merged order is its place in its own list plus,
    // for every other list, the number of keys there that order before it: one binary search per other list instead of
    // a select and a sort.  Equal keys of different shards (cannot happen: rows are distinct) would go lower shard first.
    for (int i = tid; i < n; i += blockDim.x) {
        int part = 0;
        while (i >= s_off[part + 1]) ++part;
        const uint64_t key = keys[i];
        int rank = i - s_off[part];
        for (int o = 0; o < p.parts; ++o) {
            if (o == part) continue;
            int lo = s_off[o], hi = s_off[o + 1];
            const int base = lo;
            while (lo < hi) {              // first index whose key does NOT order before `key`
                const int mid = (lo + hi) >> 1;
                const uint64_t x = keys[mid];
                if (x > key || (x == key && o < part)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo - base;
        }
        if (rank < p.k) {
            top[rank] = key;
            if (rank == p.k - 1) check_seed(ord2f(key_ord(key)));
        }
    }
    if (n < p.k && tid == 0) check_seed(-FLT_MAX);
    __syncthreads();
    // consecutive threads store consecutive entries: the destination may be the collecting rank's HBM over NVLink or
    // page-locked host memory over PCIe, where scattered 4- and 8-byte stores would each travel as their own packet
    const int m = min(n, p.k);
    for (int i = tid; i < p.k; i += blockDim.x) {
        if (i < m) {
            const uint64_t key = top[i];
            const uint32_t row = key_row(key);
            os[i] = ord2f(key_ord(key));
            oi[i] = p.id_map ? p.id_map[row] : (long long)row;
        } else {                                        // fewer than k rows in the whole index
            os[i] = -FLT_MAX;
            oi[i] = -1;
        }
    }
}

// Replicated queries without G uploads of the same bytes: this rank's part of the batch (uploaded host -> own block) is
// stored into the same place of every other rank's block over NVLink.  16-byte vectors; n16 = number of float4.
__global__ void node_spread_kernel(PeerPtrs qx, int world, int rank, size_t off16, size_t n16) {
    const float4* src = static_cast<const float4*>(qx.p[rank]) + off16;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) {
        const float4 v = src[i];
        for (int p = 0; p < world; ++p)
            if (p != rank) (static_cast<float4*>(qx.p[p]) + off16)[i] = v;
    }
}

// One rank tells all ranks which registered output set this batch goes to.
__global__ void node_publish_kernel(PeerPtrs ctrl, int world, int value) {
    if (threadIdx.x < world) static_cast<int*>(ctrl.p[threadIdx.x])[0] = value;
}

// End of a batch: the queries raised in qfail, in ascending order, the device counters and the completion mark go to
// a page-locked status slot the host reads after the batch's event.  One CTA of 1024 threads.
struct BatchStatus {
    int nfail;
    int pad;
    unsigned long long stats[ST_COUNT];
    int idx[CLDRD_QUERY_BATCH];
};

__global__ void __launch_bounds__(1024) node_tail_kernel(const int* qfail, int nq, const unsigned long long* stats,
                                                        BatchStatus* hs) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x;
    const int per = (nq + 1023) / 1024;
    const int lo = tid * per, hi = min(nq, lo + per);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += qfail[i] != 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const int w = s_warp[tid];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += t;
        }
        s_warp[tid] = wi - w;
        if (tid == 31) hs->nfail = wi;
    }
    __syncthreads();
    int at = s_warp[tid >> 5] + incl - c;
    for (int i = lo; i < hi; ++i)
        if (qfail[i] != 0) hs->idx[at++] = i;
    if (tid < ST_COUNT) hs->stats[tid] = stats[tid];
}

}  // namespace cldrd
