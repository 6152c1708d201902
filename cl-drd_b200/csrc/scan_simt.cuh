// fp32 FFMA scan: 128 queries x 128 rows per CTA, fused threshold filter or dense dump.
// This is the correctness anchor and the path for shapes TMA cannot describe (d*4 not a
// multiple of 16 bytes, unaligned buffers).  It is NOT the fast path: see scan_tc.cuh.
#pragma once
#include "device_common.cuh"

namespace cldrd {

struct ScanParams {
    // operands
    const float* xb;      // fp32 rows of the shard (SIMT scan)
    const float* q;       // [nq][d] fp32 (SIMT scan)
    int d;
    int nq;
    int64_t row_begin;    // first local row of the chunk
    int nrows;            // rows (score columns) in the chunk
    int tile_stride;      // 1 = contiguous chunk; s > 1 = sampled scan: column c reads local row
                          // row_begin + (c / 256) * s * 256 + c % 256   (every s-th 256-row tile)
    // fused filter.  Survivors of query q land in its private slice surv[q*q_stride ..), which
    // is cut into `groups` single-writer segments of seg_cap entries followed by one shared
    // overflow pool of pool_cap entries (appended with atomics, only when a segment is full: a
    // hot row range that one CTA happens to scan).  seg_cnt[q*(groups+1) + g] counts segment g
    // (it keeps running past seg_cap: the excess went to the pool), slot `groups` counts the pool.
    const float* thr;     // [nq]
    uint64_t* surv;       // [nq][q_stride]
    int* seg_cnt;         // [nq][groups]
    int q_stride;
    int seg_cap;
    int pool_cap;
    int groups;           // segments per query: SIMT scan 1 (atomic append), TC scan min(#groups, #CTAs)
    int seg_by_group;     // TC scan: segment index = group (1) or CTA (0)
    int run_len;          // TC scan: consecutive row tiles per work unit
    int* unit_ctr;        // TC scan: global work-unit counter of this launch (zeroed by the host)
    // dense dump
    float* dense;         // [nq][dense_ld]
    int dense_ld;
    unsigned long long* stats;
    unsigned long long* wait_cycles;   // TC scan, optional [4]: MMA issuer cycles waiting on operands / TMEM / unit ids, total
    // tensor-core scan only
    int num_kb;           // K blocks of 128 bytes
    int kb_elems;         // elements per K block (64 half, 32 tf32)
};

constexpr int SIMT_BM = 128, SIMT_BN = 128, SIMT_BK = 16, SIMT_LD = 132;

template <bool DENSE, bool VEC>
__global__ void __launch_bounds__(256) scan_simt_kernel(ScanParams p) {
    __shared__ __align__(16) float As[SIMT_BK][SIMT_LD];
    __shared__ __align__(16) float Bs[SIMT_BK][SIMT_LD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SIMT_BM;
    const int n0 = blockIdx.x * SIMT_BN;
    const float* __restrict__ A = p.q;
    // source row of this CTA's first column (a 128-column tile never straddles a 256-row block)
    const int64_t src0 = p.row_begin + int64_t(n0 >> 8) * p.tile_stride * 256 + (n0 & 255);
    const float* __restrict__ B = p.xb + size_t(src0 - n0) * p.d;
    const int d = p.d;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // each thread stages 2 x (1 row, 4 consecutive k) of A and of B per K block
    float4 ra[2], rb[2];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            const int k = k0 + kq;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            const int qa = m0 + r, rbw = n0 + r;
            if (VEC) {
                if (qa < p.nq && k < d) va = *reinterpret_cast<const float4*>(A + size_t(qa) * d + k);
                if (rbw < p.nrows && k < d) vb = __ldg(reinterpret_cast<const float4*>(B + size_t(rbw) * d + k));
            } else {
                if (qa < p.nq) {
                    const float* s = A + size_t(qa) * d;
                    if (k + 0 < d) va.x = s[k + 0];
                    if (k + 1 < d) va.y = s[k + 1];
                    if (k + 2 < d) va.z = s[k + 2];
                    if (k + 3 < d) va.w = s[k + 3];
                }
                if (rbw < p.nrows) {
                    const float* s = B + size_t(rbw) * d;
                    if (k + 0 < d) vb.x = __ldg(s + k + 0);
                    if (k + 1 < d) vb.y = __ldg(s + k + 1);
                    if (k + 2 < d) vb.z = __ldg(s + k + 2);
                    if (k + 3 < d) vb.w = __ldg(s + k + 3);
                }
            }
            ra[i] = va;
            rb[i] = vb;
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            As[kq + 0][r] = ra[i].x;
            As[kq + 1][r] = ra[i].y;
            As[kq + 2][r] = ra[i].z;
            As[kq + 3][r] = ra[i].w;
            Bs[kq + 0][r] = rb[i].x;
            Bs[kq + 1][r] = rb[i].y;
            Bs[kq + 2][r] = rb[i].z;
            Bs[kq + 3][r] = rb[i].w;
        }
    };

    load_tile(0);
    for (int k0 = 0; k0 < d; k0 += SIMT_BK) {
        __syncthreads();
        store_tile();
        __syncthreads();
        if (k0 + SIMT_BK < d) load_tile(k0 + SIMT_BK);
#pragma unroll
        for (int k = 0; k < SIMT_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int qi = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (qi >= p.nq) continue;
        float t = 0.f;
        if (!DENSE) t = p.thr[qi];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (col >= p.nrows) continue;
            const float v = acc[i][j];
            if (DENSE) {
                p.dense[size_t(qi) * p.dense_ld + col] = v;
            } else if (v >= t) {
                const uint64_t key = make_key(v, uint32_t(src0 - n0 + col));
                int* cnts = p.seg_cnt + size_t(qi) * (p.groups + 1);
                uint64_t* slice = p.surv + size_t(qi) * p.q_stride;
                const int slot = atomicAdd(&cnts[0], 1);
                if (slot < p.seg_cap) {
                    slice[slot] = key;
                } else {
                    const int pos = atomicAdd(&cnts[p.groups], 1);
                    if (pos < p.pool_cap) slice[size_t(p.groups) * p.seg_cap + pos] = key;
                }
            }
        }
    }
}

}  // namespace cldrd
