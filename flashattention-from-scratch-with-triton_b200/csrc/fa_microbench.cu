// Micro-benchmarks of the on-chip paths the attention kernels lean on (one CTA, clock64 timing):
//   tmem_ld : tcgen05.ld 32x32b.x32 throughput with 1..8 warps (is the TMEM read port a limiter for the math warps?)
//   mufu    : ex2.approx throughput per SM, for reference
// build/fa_microbench   (prints bytes/clk/SM and instr/clk)
#include "fa_ptx.cuh"
#include <cstdio>
using namespace fa;

__global__ void __launch_bounds__(512, 1) k_tmem_ld(int nwarps, int iters, int batch, long long* out_clk, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            uint32_t v[4][32];
            #pragma unroll
            for (int b = 0; b < 4; ++b) if (b < batch) tmem_ld32(tmem + ((i * 4 + b) * 32 & 511), v[b]);
            tc_wait_ld();
            #pragma unroll
            for (int b = 0; b < 4; ++b) if (b < batch) acc += __uint_as_float(v[b][0] ^ v[b][13] ^ v[b][31]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) *out_clk = t1 - t0;
    sink[threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

__global__ void __launch_bounds__(512, 1) k_mufu(int nwarps, int iters, long long* out_clk, float* sink) {
    const int warp = threadIdx.x >> 5;
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps)
        for (int i = 0; i < iters; ++i) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = ex2_approx(x[j]) - 1.0f;
        }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) *out_clk = t1 - t0;
    float a = 0; for (int j = 0; j < 8; ++j) a += x[j];
    sink[threadIdx.x] = a;
}

int main() {
    long long* d_clk; float* d_sink; long long h;
    cudaMalloc(&d_clk, 8); cudaMalloc(&d_sink, 4096);
    const int iters = 2000;
    for (int batch : {1, 2, 4})
        for (int nw : {1, 2, 4, 8, 16}) {
            k_tmem_ld<<<1, 512>>>(nw, iters, batch, d_clk, d_sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("tmem_ld failed\n"); return 1; }
            cudaMemcpy(&h, d_clk, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * iters * batch * 32 * 32 * 4;
            printf("tmem_ld warps=%2d loads_in_flight=%d : %lld clk, %.1f B/clk/SM, %.1f clk per x32 load per warp\n", nw, batch, h,
                   bytes / h, (double)h / (iters * batch));
        }
    for (int nw : {4, 8, 16}) {
        k_mufu<<<1, 512>>>(nw, iters, d_clk, d_sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d_clk, 8, cudaMemcpyDeviceToHost);
        printf("mufu ex2+fadd warps=%2d : %lld clk, %.2f ex2/clk/SM\n", nw, h, (double)nw * 32 * iters * 8 / h);
    }
    return 0;
}
