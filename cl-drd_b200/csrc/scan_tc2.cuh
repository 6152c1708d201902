// 2-CTA tensor-core scan (tcgen05 cta_group::2): a CTA pair on one TPC computes a 256-query x
// 256-row tile.  Each CTA stages its own 128 queries (A half) and HALF of the row tile (128 rows
// of B); one tcgen05.mma issued by the leader reads both CTAs' shared memory and writes each
// CTA's 128 accumulator lanes into its own TMEM.  Per FLOP this halves the index-row bytes a CTA
// pulls through L2/TMA and holds in shared memory (32 KiB per stage instead of 48 KiB, so the
// ring is 6 deep instead of 4), which is what a power-capped scan needs.
//
//   both CTAs   warp 0  TMA producer for its own A tile and B half (cta_group::2 loads signal the
//                       LEADER's full barrier), warps 2..5 epilogue over their own TMEM lanes
//   leader only warp 1  MMA issuer; its commits are multicast to both CTAs' empty / tmem-full
//                       barriers; warp 0 also draws the work units and publishes each id into
//                       both CTAs' unit rings (remote shared-memory store + remote mbarrier arrive)
//   waits       plain CTA-scope mbarrier waits everywhere (operands arrive through the async proxy,
//               accumulators through TMEM) except where a thread reads DATA another CTA stored:
//               the unit ring in the peer (cluster-scope acquire; it costs an L1 invalidate)
//   epilogue    identical to the 1-CTA kernel (tc_epilogue_tile); the peer's warps hand the
//               accumulator stage back with a remote arrive on the leader's tmem-empty barrier
#pragma once
#include "scan_tc.cuh"

namespace cldrd {

constexpr int TC2_STAGES = 6;
constexpr int TC2_A_STAGE = TC_BM * TC_KB_BYTES;         // 16 KiB: this CTA's 128 queries
constexpr int TC2_B_STAGE = (TC_BN / 2) * TC_KB_BYTES;   // 16 KiB: this CTA's 128 of the 256 rows
constexpr int TC2_STAGE_BYTES = TC2_A_STAGE + TC2_B_STAGE;
constexpr int TC2_RING = 8;
constexpr size_t TC2_SMEM_BYTES = size_t(TC2_STAGES) * TC2_STAGE_BYTES + 1024 + 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_smem_addr), "r"(rank));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, unsigned long long* err_slot,
                                                  uint32_t code) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) {
            if (err_slot) atomicMax(err_slot, (unsigned long long)code);
            __threadfence_system();
            __trap();
        }
    }
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t crd0,
                                                int32_t crd1, uint64_t cache_hint) {
    const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(crd0), "r"(crd1),
          "l"(cache_hint)
        : "memory");
}
__device__ __forceinline__ void tc2_commit_mcast(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(uint16_t(3))
        : "memory");
}
template <bool kTF32>
__device__ __forceinline__ void tc2_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    if constexpr (kTF32) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// M = 256 (two CTAs x 128 lanes), N = 256
__host__ __device__ constexpr uint32_t tc2_idesc(int kind) {
    return (1u << 4) | (uint32_t(kind) << 7) | (uint32_t(kind) << 10) | (uint32_t(TC_BN >> 3) << 17) |
           (uint32_t((2 * TC_BM) >> 4) << 24);
}

// tmA: queries, box 128 x 128 B.  tmBh: index rows, box 128 rows x 128 B (half a row tile).
template <int KIND, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, ScanParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* smemA = smem;
    unsigned char* smemB = smem + size_t(TC2_STAGES) * TC2_A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(TC2_STAGES) * TC2_STAGE_BYTES);
    uint64_t* full_bar = bars;                                  // [STAGES] leader: TMA (both CTAs) -> MMA
    uint64_t* empty_bar = bars + TC2_STAGES;                    // [STAGES] each CTA: MMA commit -> its producer
    uint64_t* tfull_bar = bars + 2 * TC2_STAGES;                // [2] each CTA: MMA commit -> its epilogue
    uint64_t* tempty_bar = bars + 2 * TC2_STAGES + 2;           // [2] leader: both epilogues -> MMA
    uint64_t* ufull_bar = bars + 2 * TC2_STAGES + 4;            // [RING] each CTA: unit id published
    uint64_t* uempty_bar = bars + 2 * TC2_STAGES + 4 + TC2_RING;  // [RING] leader: all consumers of both CTAs
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC2_STAGES + 4 + 2 * TC2_RING);
    volatile int* unit_ring = reinterpret_cast<volatile int*>(tmem_base_slot + 2);  // [RING]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    unsigned long long* err = p.stats ? &p.stats[ST_KERNEL_ERR] : nullptr;

    const int num_m = (p.nq + TC_BM - 1) / TC_BM;
    const int num_mp = (num_m + 1) / 2;                          // query-tile pairs
    const int num_n = (p.nrows + TC_BN - 1) / TC_BN;
    const int num_groups = (num_n + p.run_len - 1) / p.run_len;
    const int num_units = num_mp * num_groups;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmBh);
        for (int s = 0; s < TC2_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 8);      // 4 epilogue warps of each CTA
        }
        for (int r = 0; r < TC2_RING; ++r) {
            mbar_init(&ufull_bar[r], 1);
            mbar_init(&uempty_bar[r], 10);     // leader: MMA + 4 epilogue; peer: producer + 4 epilogue
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem2_alloc(tmem_base_slot, TC_TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                        // both CTAs' barriers and TMEM are ready
    __syncthreads();                           // (the cluster barrier already orders this; racecheck only models this one)
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    // every consumer of a unit id frees the ring slot on the LEADER's barrier
    auto release_unit_slot = [&](int uq) {
        if (leader) mbar_arrive(&uempty_bar[uq]);
        else mbar_arrive_cluster(map_to_cta(smem_u32(&uempty_bar[uq]), 0));
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        int uq = 0;
        uint32_t uphase = 0;
        int u_next = 0;
        if (leader && lane == 0) u_next = atomicAdd(p.unit_ctr, 1);
        for (;;) {
            int u;
            if (leader) {
                u = __shfl_sync(0xffffffffu, u_next, 0);
                mbar_wait(&uempty_bar[uq], uphase ^ 1, err, 500 + uq);
                if (lane == 0) {
                    const int id = u < num_units ? u : -1;
                    unit_ring[uq] = id;
                    st_cluster_u32(map_to_cta(smem_u32(const_cast<int*>(&unit_ring[uq])), 1), uint32_t(id));
                    mbar_arrive(&ufull_bar[uq]);
                    mbar_arrive_cluster(map_to_cta(smem_u32(&ufull_bar[uq]), 1));
                }
                __syncwarp();
            } else {
                mbar_wait_cluster(&ufull_bar[uq], uphase, err, 550 + uq);
                u = unit_ring[uq];
                __syncwarp();
                if (lane == 0) release_unit_slot(uq);
                if (u < 0) u = num_units;
            }
            if (++uq == TC2_RING) {
                uq = 0;
                uphase ^= 1;
            }
            if (u >= num_units) break;
            if (leader && lane == 0) u_next = atomicAdd(p.unit_ctr, 1);
            int mp, g;
            unit_to_tile(u, num_mp, mp, g);
            const int n_end = min(num_n, (g + 1) * p.run_len);
            const int crd_q = (2 * mp + int(rank)) * TC_BM;      // a tile past nq is zero-filled by TMA
            for (int n = g * p.run_len; n < n_end; ++n) {
                const int crd_r = int(p.row_begin) + (n * TC_BN) * p.tile_stride + int(rank) * (TC_BN / 2);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, err, 100 + stage);
                    if (lane == 0) {
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * TC2_STAGE_BYTES);  // both CTAs' bytes
                        tma_load_2d_2sm(smemA + size_t(stage) * TC2_A_STAGE, &tmA, &full_bar[stage], kb * p.kb_elems,
                                        crd_q, kEvictLast);
                        tma_load_2d_2sm(smemB + size_t(stage) * TC2_B_STAGE, &tmBh, &full_bar[stage], kb * p.kb_elems,
                                        crd_r, kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == TC2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc = tc2_idesc(KIND);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            int uq = 0;
            uint32_t uphase = 0;
            long long t_full = 0, t_tempty = 0, t_unit = 0;   // cycles the issuer spent waiting, by cause
            const long long t_begin = clock64();
            for (;;) {
                long long t0 = clock64();
                mbar_wait(&ufull_bar[uq], uphase, err, 600 + uq);
                t_unit += clock64() - t0;
                const int u = unit_ring[uq];
                __syncwarp();
                if (lane == 0) release_unit_slot(uq);
                if (++uq == TC2_RING) {
                    uq = 0;
                    uphase ^= 1;
                }
                if (u < 0) break;
                const int g = u / num_mp;
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
                            const uint64_t a_desc = umma_desc_sw128(smem_u32(smemA + size_t(stage) * TC2_A_STAGE));
                            const uint64_t b_desc = umma_desc_sw128(smem_u32(smemB + size_t(stage) * TC2_B_STAGE));
#pragma unroll
                            for (int kk = 0; kk < TC_KB_BYTES / 32; ++kk)
                                tc2_mma_ss<KIND == 2>(d_tmem, a_desc + uint64_t(kk * 2), b_desc + uint64_t(kk * 2), idesc,
                                                      uint32_t((kb | kk) != 0));
                            tc2_commit_mcast(&empty_bar[stage]);                       // frees the slot in BOTH CTAs
                            if (kb == p.num_kb - 1) tc2_commit_mcast(&tfull_bar[as]);  // wakes BOTH epilogues
                        }
                        __syncwarp();
                        if (++stage == TC2_STAGES) {
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
        }
    } else {
        // ===================== epilogue (both CTAs, own TMEM lanes) =====================
        const int qd = warp & 3;
        uint32_t it = 0;
        unsigned long long tiles_done = 0;
        int uq = 0;
        uint32_t uphase = 0;
        for (;;) {
            mbar_wait_cluster(&ufull_bar[uq], uphase, err, 700 + uq);
            const int u = unit_ring[uq];
            __syncwarp();
            if (lane == 0) release_unit_slot(uq);
            if (++uq == TC2_RING) {
                uq = 0;
                uphase ^= 1;
            }
            if (u < 0) break;
            int mp, g;
            unit_to_tile(u, num_mp, mp, g);
            const int m = 2 * mp + int(rank);
            const int n_end = min(num_n, (g + 1) * p.run_len);
            const int qrow = m * TC_BM + qd * 32 + lane;
            const bool qvalid = qrow < p.nq;
            float thr = INFINITY;
            if (MODE == TC_FILTER && qvalid) thr = p.thr[qrow];
            const int seg = p.seg_by_group ? g : int(blockIdx.x >> 1);   // one segment per CTA PAIR: its two CTAs hold different query tiles
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
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(&tempty_bar[as]);
                    else mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[as]), 0));
                }
                ++tiles_done;
            }
            if (MODE == TC_FILTER && qvalid) *cnt_slot = cnt;
        }
        if (warp == 2 && lane == 0 && p.stats) atomicAdd(&p.stats[ST_TILES], tiles_done);
    }

    tc_fence_before();
    cluster_sync_all();      // nobody leaves (or frees TMEM) while the pair still uses this CTA's memory
    if (warp == 1) {
        tc_fence_after();
        tmem2_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

}  // namespace cldrd
