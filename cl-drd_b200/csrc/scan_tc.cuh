// Tensor-core scan for sm_100a: TMA -> shared (128B swizzle) -> tcgen05.mma -> TMEM -> fused
// per-query threshold filter (or dense dump).  The score tile never reaches HBM.
//
//   CTA tile      128 queries (TMEM lanes) x 256 index rows (TMEM columns), K streamed in
//                 128-byte blocks (64 fp16/bf16 or 32 tf32 elements), 4-stage mbarrier ring
//   accumulators  2 x 256 fp32 columns = all 512 TMEM columns: the epilogue of tile i overlaps
//                 the MMAs of tile i+1
//   warps         0: TMA producer   1: MMA issuer + TMEM owner   2..5: epilogue (one per TMEM
//                 lane quadrant; thread = one query: its threshold, its survivor count and its
//                 output cursor live in registers)
//   work units    unit u = (group g = u / num_m, query tile m = hashed rotation of u % num_m) covers `run_len`
//                 consecutive row tiles for ONE query tile.  Persistent CTAs draw units from a
//                 global counter (the producer warp fetches the next id while it streams the
//                 current unit and hands it to the other warps through a 4-deep shared-memory
//                 ring), so at any moment the CTAs in flight sit on ~148 CONSECUTIVE units = every
//                 query tile of ~3 short groups: each index row tile is pulled from HBM once and
//                 served to the other query tiles from L2 while it is still hot, however long the
//                 launch runs and however unevenly the SMs progress.
//   survivors     query q owns a slice of the survivor buffer cut into min(#groups, #CTAs)
//                 segments: one per group while groups are few, one per CTA otherwise.  A
//                 segment is written by exactly one thread at a time; its cursor is carried in
//                 a register across a unit and parked in seg_cnt between units (no atomics, no
//                 shared counters), and the survivors spread evenly over the segments.  The last
//                 quarter of the slice is an overflow pool shared by all segments of the query
//                 (atomic cursor, out-of-line filter_group_generic): a segment that fills up -
//                 rows one query likes stored next to each other - spills there instead of
//                 sending the query to the fallback.
#pragma once
#include "device_common.cuh"
#include "scan_simt.cuh"  // ScanParams

namespace cldrd {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_KB_BYTES = 128;
constexpr int TC_STAGES = 4;
constexpr int TC_A_STAGE = TC_BM * TC_KB_BYTES;  // 16 KiB
constexpr int TC_B_STAGE = TC_BN * TC_KB_BYTES;  // 32 KiB
constexpr int TC_STAGE_BYTES = TC_A_STAGE + TC_B_STAGE;
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_UNIT_RING = 4;
constexpr size_t TC_SMEM_BYTES = size_t(TC_STAGES) * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers + unit ring*/;

// kind: 0 = fp16, 1 = bf16, 2 = tf32.  Instruction descriptor (cute::UMMA::InstrDescriptor):
// c_format F32 [4,6) | a_format [7,10) | b_format [10,13) | K-major A and B (bits 15,16 = 0)
// | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t tc_idesc(int kind) {
    return (1u << 4) | (uint32_t(kind) << 7) | (uint32_t(kind) << 10) | (uint32_t(TC_BN >> 3) << 17) |
           (uint32_t(TC_BM >> 4) << 24);
}

// MODE: what the epilogue does with a score tile
constexpr int TC_FILTER = 0;  // append scores >= thr[q] to the query's survivor segment
constexpr int TC_DENSE = 1;   // write every score to dense[q][col] (first piece, fallback, tests)
constexpr int TC_MAXES = 2;   // write the max of every 32-column group to dense[q][col/32] (seed sample)

// Work unit u -> (query tile m, group g).  Units of one group are consecutive (so the CTAs in
// flight share its row tiles through L2), but the query tile is rotated by a per-group hash:
// without it a CTA that takes every gridDim-th unit would only ever see the query tiles of one
// residue class mod gcd(gridDim, num_m), and those queries' survivors would pile up in a few
// (query, CTA) segments instead of spreading over all of them.
__device__ __forceinline__ void unit_to_tile(int u, int num_m, int& m, int& g) {
    g = u / num_m;
    const int j = u - g * num_m;
    const uint32_t rot = (uint32_t(g) * 2654435761u) >> 12;
    m = int((uint32_t(j) + rot) % uint32_t(num_m));
}

// Rare, warp-uniform path of the filter: some thread's single-writer segment may not have room for
// 32 more survivors (a hot row range scanned by one CTA), or the tile is the chunk's ragged last
// one.  Re-reads the 32-column group from TMEM and appends with every check; survivors that do not
// fit the segment go to the query's shared overflow pool (atomic append).  Out of line so that the
// common path carries neither the checks nor the register pressure of a call.
__device__ __noinline__ int filter_group_generic(const ScanParams& p, uint32_t taddr_b, int qrow, float thr,
                                                 uint32_t row0b, int lim, int cnt, uint64_t* dst) {
    float v[32];
    tmem_ld32(taddr_b, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if ((v[j] >= thr) && (j < lim)) {
            const uint64_t key = make_key(v[j], row0b + uint32_t(j));
            if (cnt < p.seg_cap) {
                dst[cnt] = key;
            } else {
                const int pos = atomicAdd(p.seg_cnt + size_t(qrow) * (p.groups + 1) + p.groups, 1);
                if (pos < p.pool_cap) p.surv[size_t(qrow) * p.q_stride + size_t(p.groups) * p.seg_cap + pos] = key;
            }
            ++cnt;
        }
    }
    return cnt;
}

__device__ __forceinline__ float max8(const float* v) {
    return fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
}

// One 128 x 256 score tile, seen by one epilogue thread (= one query = TMEM lane): `taddr` is
// the accumulator stage at this warp's lane quadrant.  Shared by the 1-CTA and 2-CTA scans.
template <int MODE>
__device__ __forceinline__ void tc_epilogue_tile(const ScanParams& p, uint32_t taddr, int qrow, bool qvalid,
                                                 float thr, uint64_t* dst, int& cnt, int n, int valid_n) {
    float v[32];
    if (MODE == TC_MAXES) {
        // seed sample: only whole tiles are sampled, so every column is valid
        float mxs[TC_BN / 32];
#pragma unroll
        for (int b = 0; b < TC_BN / 32; ++b) {
            tmem_ld32(taddr + uint32_t(b * 32), v);
            float mx = v[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
            mxs[b] = mx;
        }
        if (qvalid) {
            float4* o = reinterpret_cast<float4*>(p.dense + size_t(qrow) * p.dense_ld + size_t(n) * (TC_BN / 32));
            o[0] = make_float4(mxs[0], mxs[1], mxs[2], mxs[3]);
            o[1] = make_float4(mxs[4], mxs[5], mxs[6], mxs[7]);
        }
    } else if (MODE == TC_DENSE) {
#pragma unroll 1
        for (int b = 0; b < TC_BN / 32; ++b) {
            tmem_ld32(taddr + uint32_t(b * 32), v);
            if (qvalid) {
                float* o = p.dense + size_t(qrow) * p.dense_ld + size_t(n) * TC_BN + b * 32;
                if (b * 32 + 32 <= valid_n && (p.dense_ld & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (b * 32 + j < valid_n) o[j] = v[j];
                }
            }
        }
    } else {
        const uint32_t row0 = uint32_t(p.row_begin) + uint32_t(n) * uint32_t(TC_BN * p.tile_stride);
        float w[64];
#pragma unroll 1
        for (int b2 = 0; b2 < TC_BN / 64; ++b2) {
            tmem_ld64(taddr + uint32_t(b2 * 64), w);       // two 32-column groups per TMEM round trip
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int b = 2 * b2 + h;
                const float* vv = w + 32 * h;
                float m8[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) m8[s] = max8(vv + 8 * s);
                const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
                if (__any_sync(0xffffffffu, mx >= thr)) {
                    // some thread of the warp has a survivor in this 32-column group
                    const int lim = valid_n - b * 32;   // >= 32 except on the chunk's last tile
                    if (__any_sync(0xffffffffu, cnt + 32 > p.seg_cap) || lim < 32) {
                        cnt = filter_group_generic(p, taddr + uint32_t(b * 32), qrow, thr, row0 + uint32_t(b * 32), lim, cnt, dst);
                    } else {
                        // common: room for the whole group, every column valid.  Expand only the
                        // 8-column sub-groups in which some thread of the warp has a hit.
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            if (__any_sync(0xffffffffu, m8[s] >= thr)) {
#pragma unroll
                                for (int j = 8 * s; j < 8 * s + 8; ++j) {
                                    if (vv[j] >= thr) {
                                        dst[cnt] = make_key(vv[j], row0 + uint32_t(b * 32 + j));
                                        ++cnt;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

template <int KIND, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               ScanParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 128B-swizzled operand tiles need 1024-byte alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* smemA = smem;
    unsigned char* smemB = smem + size_t(TC_STAGES) * TC_A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(TC_STAGES) * TC_STAGE_BYTES);
    uint64_t* full_bar = bars;                     // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + TC_STAGES;        // [STAGES] MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * TC_STAGES;    // [2] MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * TC_STAGES + 2;  // [2] epilogue -> MMA
    uint64_t* ufull_bar = bars + 2 * TC_STAGES + 4;                   // [RING] producer -> consumers: unit id published
    uint64_t* uempty_bar = bars + 2 * TC_STAGES + 4 + TC_UNIT_RING;   // [RING] consumers -> producer: slot read
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4 + 2 * TC_UNIT_RING);
    volatile int* unit_ring = reinterpret_cast<volatile int*>(tmem_base_slot + 2);    // [RING]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned long long* err = p.stats ? &p.stats[ST_KERNEL_ERR] : nullptr;

    const int num_m = (p.nq + TC_BM - 1) / TC_BM;
    const int num_n = (p.nrows + TC_BN - 1) / TC_BN;
    const int num_groups = (num_n + p.run_len - 1) / p.run_len;
    const int num_units = num_m * num_groups;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);
        }
        for (int r = 0; r < TC_UNIT_RING; ++r) {
            mbar_init(&ufull_bar[r], 1);
            mbar_init(&uempty_bar[r], 5);   // MMA warp + 4 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_base_slot, TC_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        int uq = 0;
        uint32_t uphase = 0;
        int u_next = 0;
        if (lane == 0) u_next = atomicAdd(p.unit_ctr, 1);
        for (;;) {
            const int u = __shfl_sync(0xffffffffu, u_next, 0);
            // publish the unit (or the end marker) to the MMA and epilogue warps
            mbar_wait(&uempty_bar[uq], uphase ^ 1, err, 500 + uq);
            if (lane == 0) {
                unit_ring[uq] = u < num_units ? u : -1;
                mbar_arrive(&ufull_bar[uq]);
            }
            if (++uq == TC_UNIT_RING) {
                uq = 0;
                uphase ^= 1;
            }
            if (u >= num_units) break;
            if (lane == 0) u_next = atomicAdd(p.unit_ctr, 1);   // latency hidden behind this unit's loads
            int m, g;
            unit_to_tile(u, num_m, m, g);
            const int n_end = min(num_n, (g + 1) * p.run_len);
            const int crd_q = m * TC_BM;
            for (int n = g * p.run_len; n < n_end; ++n) {
                const int crd_r = int(p.row_begin) + n * TC_BN * p.tile_stride;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, err, 100 + stage);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
                        // queries are re-read by every row tile: keep them in L2; index rows stream
                        tma_load_2d(smemA + size_t(stage) * TC_A_STAGE, &tmA, &full_bar[stage],
                                    kb * p.kb_elems, crd_q, kEvictLast);
                        tma_load_2d(smemB + size_t(stage) * TC_B_STAGE, &tmB, &full_bar[stage],
                                    kb * p.kb_elems, crd_r, kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = tc_idesc(KIND);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        int uq = 0;
        uint32_t uphase = 0;
        long long t_full = 0, t_tempty = 0, t_unit = 0;     // cycles the issuer spent waiting, by cause
        const long long t_begin = clock64();
        for (;;) {
            long long t0 = clock64();
            mbar_wait(&ufull_bar[uq], uphase, err, 600 + uq);
            t_unit += clock64() - t0;
            const int u = unit_ring[uq];
            __syncwarp();
            if (lane == 0) mbar_arrive(&uempty_bar[uq]);
            if (++uq == TC_UNIT_RING) {
                uq = 0;
                uphase ^= 1;
            }
            if (u < 0) break;
            const int g = u / num_m;
            const int n_end = min(num_n, (g + 1) * p.run_len);
            for (int n = g * p.run_len; n < n_end; ++n, ++it) {
                const uint32_t as = it & 1u;
                const uint32_t aphase = (it >> 1) & 1u;
                t0 = clock64();
                mbar_wait(&tempty_bar[as], aphase ^ 1, err, 200 + as);
                t_tempty += clock64() - t0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * uint32_t(TC_BN);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    t0 = clock64();
                    mbar_wait(&full_bar[stage], phase, err, 300 + stage);
                    t_full += clock64() - t0;
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t a_desc = umma_desc_sw128(smem_u32(smemA + size_t(stage) * TC_A_STAGE));
                        const uint64_t b_desc = umma_desc_sw128(smem_u32(smemB + size_t(stage) * TC_B_STAGE));
#pragma unroll
                        for (int kk = 0; kk < TC_KB_BYTES / 32; ++kk) {
                            // one MMA consumes 32 bytes of K: +2 in the 16-byte start-address field
                            tc_mma_ss<KIND == 2>(d_tmem, a_desc + uint64_t(kk * 2), b_desc + uint64_t(kk * 2),
                                                 idesc, uint32_t((kb | kk) != 0));
                        }
                        tc_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
                        if (kb == p.num_kb - 1) tc_commit(&tfull_bar[as]);  // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        if (lane == 0 && p.wait_cycles) {
            atomicAdd(&p.wait_cycles[0], (unsigned long long)t_full);
            atomicAdd(&p.wait_cycles[1], (unsigned long long)t_tempty);
            atomicAdd(&p.wait_cycles[2], (unsigned long long)t_unit);
            atomicAdd(&p.wait_cycles[3], (unsigned long long)(clock64() - t_begin));
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> filter =====================
        const int qd = warp & 3;  // the TMEM lane quadrant this warp may read
        uint32_t it = 0;
        unsigned long long tiles_done = 0;
        int uq = 0;
        uint32_t uphase = 0;
        for (;;) {
            mbar_wait(&ufull_bar[uq], uphase, err, 700 + uq);
            const int u = unit_ring[uq];
            __syncwarp();
            if (lane == 0) mbar_arrive(&uempty_bar[uq]);
            if (++uq == TC_UNIT_RING) {
                uq = 0;
                uphase ^= 1;
            }
            if (u < 0) break;
            int m, g;
            unit_to_tile(u, num_m, m, g);
            const int n_end = min(num_n, (g + 1) * p.run_len);
            const int qrow = m * TC_BM + qd * 32 + lane;
            const bool qvalid = qrow < p.nq;
            float thr = INFINITY;
            if (MODE == TC_FILTER && qvalid) thr = p.thr[qrow];
            // this thread's private segment: (query, group) while there are fewer groups than
            // CTAs -- every unit then owns a fresh segment -- else (query, this CTA)
            const int seg = p.seg_by_group ? g : int(blockIdx.x);
            uint64_t* dst = MODE != TC_FILTER ? nullptr : p.surv + size_t(qvalid ? qrow : 0) * p.q_stride + size_t(seg) * p.seg_cap;
            int* cnt_slot = MODE != TC_FILTER ? nullptr : p.seg_cnt + size_t(qvalid ? qrow : 0) * (p.groups + 1) + seg;
            int cnt = 0;
            if (MODE == TC_FILTER && qvalid) cnt = *cnt_slot;
            for (int n = g * p.run_len; n < n_end; ++n, ++it) {
                const uint32_t as = it & 1u;
                const uint32_t aphase = (it >> 1) & 1u;
                const int valid_n = min(TC_BN, p.nrows - n * TC_BN);
                const uint32_t taddr = tmem_base + (uint32_t(qd * 32) << 16) + as * uint32_t(TC_BN);
                mbar_wait(&tfull_bar[as], aphase, err, 400 + as);
                tc_fence_after();
                tc_epilogue_tile<MODE>(p, taddr, qrow, qvalid, thr, dst, cnt, n, valid_n);
                // accumulator stage drained: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
                ++tiles_done;
            }
            if (MODE == TC_FILTER && qvalid) *cnt_slot = cnt;
        }
        if (warp == 2 && lane == 0 && p.stats) atomicAdd(&p.stats[ST_TILES], tiles_done);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

}  // namespace cldrd
