// HBM-bound helper kernels: the backward preprocess delta = rowsum(dO * O)
// (reference code/_flash_attention_kernel_optimized.py:210-211) and the (O, LSE) partial merge used by
// the sequence-sharded ring.  Coalesced 16-byte loads, D/8 threads per row.
#pragma once
#include "fa_ptx.cuh"

namespace fa {

// lanes of the warp that share one row (TPR consecutive lanes); rows never straddle a warp
template <int TPR> __device__ __forceinline__ uint32_t row_group_mask() {
    return (TPR == 32 ? 0xffffffffu : ((1u << TPR) - 1u)) << ((threadIdx.x & 31) & ~(TPR - 1));
}

template <bool kBf16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (kBf16) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        } else {
            const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
            const float2 t = __half22float2(h);
            f[2 * i] = t.x; f[2 * i + 1] = t.y;
        }
    }
}

// delta[row] = sum_d fp32(dO[row,d]) * fp32(O[row,d])
// Element strides (batch, head, row) of a [B,H,S,D] tensor whose D is contiguous
struct RowStrides { long long b, h, r; };

// Four rows per thread with all eight 16-byte loads issued before the first use: these launches are a few tens of MB,
// so what counts is bytes in flight per SM in a single wave, not the loop.
template <int D, bool kBf16>
__global__ void __launch_bounds__(256) fa_delta_kernel(const uint4* __restrict__ o, const uint4* __restrict__ dout,
                                                       float* __restrict__ delta, long long rows, int H, int Sq,
                                                       RowStrides so, RowStrides sd, float4* __restrict__ zero_acc) {
    constexpr int TPR = D / 8;                       // threads per row
    constexpr int RPB = 256 / TPR;                   // rows per block per step
    constexpr int U = 4;                             // rows per thread per iteration
    const int sub = threadIdx.x % TPR;
    const long long stride = (long long)gridDim.x * RPB;
    pdl_wait();
    for (long long row0 = (long long)blockIdx.x * RPB + threadIdx.x / TPR; row0 < rows; row0 += stride * U) {
        uint4 a[U], b[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            if (row < rows) {
                const long long bh = row / Sq, sq = row % Sq, bb = bh / H, hh = bh % H;
                a[u] = __ldg(o + ((bb * so.b + hh * so.h + sq * so.r) >> 3) + sub);
                b[u] = __ldg(dout + ((bb * sd.b + hh * sd.h + sq * sd.r) >> 3) + sub);
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            if (row < rows) {                        // uniform over the TPR lanes of a row
                float fa_[8], fb[8];
                unpack8<kBf16>(a[u], fa_); unpack8<kBf16>(b[u], fb);
                float acc = 0.f;
                #pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(fa_[i], fb[i], acc);
                #pragma unroll
                for (int off = TPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(row_group_mask<TPR>(), acc, off);
                if (sub == 0) delta[row] = acc;
                if (zero_acc) {                      // fused backward: clear the fp32 dQ accumulator [rows, D] on the way
                    zero_acc[row * (D / 4) + sub * 2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    zero_acc[row * (D / 4) + sub * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}

inline int launch_delta(const void* o, const void* dout, float* delta, long long rows, int H, int Sq, RowStrides so,
                        RowStrides sd, int D, int dtype, int sms, cudaStream_t st, float* zero_acc = nullptr) {
    const int rpb = 256 / (D / 8) * 4;
    long long blocks = (rows + rpb - 1) / rpb;
    const long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
    const uint4* o4 = (const uint4*)o; const uint4* d4 = (const uint4*)dout;
    if (D == 64) { if (dtype) launch_pdl(fa_delta_kernel<64, true>, (int)blocks, 256, 0, st, o4, d4, delta, rows, H, Sq, so, sd, (float4*)zero_acc);
                   else launch_pdl(fa_delta_kernel<64, false>, (int)blocks, 256, 0, st, o4, d4, delta, rows, H, Sq, so, sd, (float4*)zero_acc); }
    else         { if (dtype) launch_pdl(fa_delta_kernel<128, true>, (int)blocks, 256, 0, st, o4, d4, delta, rows, H, Sq, so, sd, (float4*)zero_acc);
                   else launch_pdl(fa_delta_kernel<128, false>, (int)blocks, 256, 0, st, o4, d4, delta, rows, H, Sq, so, sd, (float4*)zero_acc); }
    return (int)cudaGetLastError();
}

// In-place merge of a partial attention result into fp32 running accumulators.
template <int D, bool kBf16>
__global__ void __launch_bounds__(256) fa_merge_kernel(float4* __restrict__ o_acc, float* __restrict__ lse_acc,
                                                       const uint4* __restrict__ o_part, const float* __restrict__ lse_part,
                                                       long long rows, int Sq, int Sq_acc, int q_off) {
    constexpr int TPR = D / 8;
    constexpr int RPB = 256 / TPR;
    const int sub = threadIdx.x % TPR;
    pdl_wait();
    for (long long row = (long long)blockIdx.x * RPB + threadIdx.x / TPR; row < rows; row += (long long)gridDim.x * RPB) {
        const long long arow = (row / Sq) * Sq_acc + q_off + (row % Sq);     // row inside the accumulator
        const float la = lse_acc[arow], lb = lse_part[row];
        const float mx = fmaxf(la, lb);
        float wa, wb, lnew;
        if (mx == -INFINITY) { wa = 0.f; wb = 0.f; lnew = -INFINITY; }
        else {
            const float ea = __expf(la - mx), eb = __expf(lb - mx);     // exp(-inf) = 0
            const float s = ea + eb;
            lnew = mx + __logf(s);
            wa = ea / s; wb = eb / s;
        }
        float4 a0 = o_acc[arow * TPR * 2 + sub * 2], a1 = o_acc[arow * TPR * 2 + sub * 2 + 1];
        float fb[8];
        unpack8<kBf16>(__ldg(o_part + row * TPR + sub), fb);
        a0.x = a0.x * wa + fb[0] * wb; a0.y = a0.y * wa + fb[1] * wb; a0.z = a0.z * wa + fb[2] * wb; a0.w = a0.w * wa + fb[3] * wb;
        a1.x = a1.x * wa + fb[4] * wb; a1.y = a1.y * wa + fb[5] * wb; a1.z = a1.z * wa + fb[6] * wb; a1.w = a1.w * wa + fb[7] * wb;
        o_acc[arow * TPR * 2 + sub * 2] = a0; o_acc[arow * TPR * 2 + sub * 2 + 1] = a1;
        // every thread of the row group has read lse_acc[row] before its leader overwrites it
        __syncwarp(row_group_mask<TPR>());
        if (sub == 0) lse_acc[arow] = lnew;
    }
}

inline int launch_merge(float* o_acc, float* lse_acc, const void* o_part, const float* lse_part, long long rows,
                        int Sq, int Sq_acc, int q_off, int D, int dtype, int sms, cudaStream_t st) {
    const int rpb = 256 / (D / 8);
    long long blocks = (rows + rpb - 1) / rpb;
    const long long cap = (long long)sms * 16;
    if (blocks > cap) blocks = cap;
    float4* oa = (float4*)o_acc; const uint4* op = (const uint4*)o_part;
    if (D == 64) { if (dtype) launch_pdl(fa_merge_kernel<64, true>, (int)blocks, 256, 0, st, oa, lse_acc, op, lse_part, rows, Sq, Sq_acc, q_off);
                   else launch_pdl(fa_merge_kernel<64, false>, (int)blocks, 256, 0, st, oa, lse_acc, op, lse_part, rows, Sq, Sq_acc, q_off); }
    else         { if (dtype) launch_pdl(fa_merge_kernel<128, true>, (int)blocks, 256, 0, st, oa, lse_acc, op, lse_part, rows, Sq, Sq_acc, q_off);
                   else launch_pdl(fa_merge_kernel<128, false>, (int)blocks, 256, 0, st, oa, lse_acc, op, lse_part, rows, Sq, Sq_acc, q_off); }
    return (int)cudaGetLastError();
}

}  // namespace fa
