// Device helpers shared by the kernels of libcldrd.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cldrd {

// ------------------------------------------------------------------------------------------
// Candidate keys.  One u64 orders candidates exactly like faiss' flat index reports them:
// higher score first, ties -> lower row.   key = orderable(score) << 32 | ~row
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    return (uint64_t(f2ord(s)) << 32) | uint64_t(~row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~uint32_t(k); }
__device__ __forceinline__ uint32_t key_ord(uint64_t k) { return uint32_t(k >> 32); }

// Per-search device counters (cldrd_shard_last_stats).
enum StatSlot {
    ST_FAILED = 0,      // queries whose survivor buffer overflowed -> dense fallback
    ST_MAX_LIST = 1,    // max candidate-list length seen
    ST_SURVIVORS = 2,   // survivors pushed by the fused filter
    ST_RESCORED = 3,    // candidates re-scored in fp32
    ST_EXACT_COMPACT = 4,  // in-kernel exact compactions (tie floods)
    ST_RANGE_ERR = 5,   // query value outside the fp16 range in an fp16 scan
    ST_TILES = 6,       // tcgen05 tiles executed
    ST_KERNEL_ERR = 7,  // mbarrier watchdog code
    ST_COUNT = 8
};

// ------------------------------------------------------------------------------------------
// Exact fp32 dot product: the ONE routine every returned score comes from, so a row's score
// is bit-identical no matter which shard, scan mode or fallback produced the candidate.
// 32 lanes each FMA a strided slice, then an xor butterfly (commutative adds -> all lanes
// hold the same bits).  q_s is the query in shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float exact_dot_warp(const float* __restrict__ q_s,
                                                const float* __restrict__ row, int d, bool vec4,
                                                int lane) {
    float acc = 0.f;
    if (vec4) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const float4* q4 = reinterpret_cast<const float4*>(q_s);
        const int n4 = d >> 2;
        // eight 16-byte loads in flight per lane: a 768-dim row is ONE round trip to HBM (the gather is latency-bound
        // when the lists are short and the clocks are down after the scan); the FMA order is the same as a plain
        // stride-32 loop, so the result does not depend on the unroll
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = lane; c < n4; c += 256) {
            float4 b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = (c + 32 * u < n4) ? __ldg(r4 + c + 32 * u) : zero;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c + 32 * u < n4) {
                    const float4 a = q4[c + 32 * u];
                    acc = fmaf(a.x, b[u].x, acc);
                    acc = fmaf(a.y, b[u].y, acc);
                    acc = fmaf(a.z, b[u].z, acc);
                    acc = fmaf(a.w, b[u].w, acc);
                }
            }
        }
    } else {
        for (int c = lane; c < d; c += 32) acc = fmaf(q_s[c], __ldg(row + c), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc + 0.0f;  // -0.0 -> +0.0 so that equal scores have equal keys
}

// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap with a code in stats[ST_KERNEL_ERR], never
// in a hung GPU.  ~2^31 cycles is about a second; no legitimate wait comes near it.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity,
                                          unsigned long long* err_slot, uint32_t code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) {
            if (err_slot) atomicMax(err_slot, (unsigned long long)code);
            __threadfence_system();
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load global -> shared, completion on an mbarrier.  crd0 = innermost (K) element,
// crd1 = row.  Out-of-bounds elements are zero-filled and still counted in the tx bytes.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t crd0, int32_t crd1, uint64_t cache_hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(crd0), "r"(crd1), "l"(cache_hint)
        : "memory");
}
// L2 eviction-priority descriptors (same encodings CUTLASS uses for TMA cache hints)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// tcgen05.commit: the mbarrier gets one arrival when every tcgen05 op issued so far by this
// thread has completed (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single-thread issue.
template <bool kTF32>
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    if constexpr (kTF32) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}"
            :
            : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of the
// warp's TMEM quadrant).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 64 consecutive fp32 columns per thread in one TMEM round trip
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (the layout TMA writes for
// a box whose inner extent is 128 bytes): 8-row groups 1024 B apart, version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units      [0,14)
    d |= uint64_t(1) << 16;                     // leading byte offset (ignored)      [16,30)
    d |= uint64_t(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B [32,46)
    d |= uint64_t(1) << 46;                     // descriptor version = 1             [46,48)
    d |= uint64_t(2) << 61;                     // layout type = SWIZZLE_128B         [61,64)
    return d;
}

}  // namespace cldrd
