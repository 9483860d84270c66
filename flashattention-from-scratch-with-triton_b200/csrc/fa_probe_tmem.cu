// Bring-up probe: which (TMEM lane, column) does each register of the 16x256b / 16x128b tcgen05.ld / st shapes hold?
// Fill 128 lanes x 64 columns through the known 32x32b shape with value = lane * 256 + column, read back with the other shapes
// and print the mapping of warp 0 and warp 1 (profiles/r02_probe_tmem_shapes.txt).  Build: build.py --bringup.
#include "fa_ptx.cuh"
#include <cstdio>
using namespace fa;

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    // fill: thread r <-> lane r, columns 0..63
    uint32_t v[32];
    for (int q = 0; q < 2; ++q) {
        for (int i = 0; i < 32; ++i) v[i] = (uint32_t)tid * 256u + (uint32_t)(q * 32 + i);
        tmem_st32(tmem + lane_field + q * 32, v);
    }
    tc_wait_st(); tc_fence_before(); __syncthreads(); tc_fence_after();
    // ---- 16x256b.x2: lanes [base, base+16), 16 columns; 8 registers per thread
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        const uint32_t addr = tmem + lane_field + ((uint32_t)(half * 16) << 16) + 8;      // start at column 8
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory");
        tc_wait_ld();
        for (int i = 0; i < 8; ++i) out[((0 * 2 + half) * 128 + tid) * 8 + i] = r[i];
    }
    // ---- 16x128b.x2: lanes [base, base+16), 8 columns; 4 registers per thread
    for (int half = 0; half < 2; ++half) {
        uint32_t r[4];
        const uint32_t addr = tmem + lane_field + ((uint32_t)(half * 16) << 16) + 8;
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
        tc_wait_ld();
        for (int i = 0; i < 4; ++i) out[((1 * 2 + half) * 128 + tid) * 8 + i] = r[i];
    }
    // ---- store through 16x128b.x2 into columns 64.. and read back through 32x32b
    tc_fence_before(); __syncthreads(); tc_fence_after();
    for (int half = 0; half < 2; ++half) {
        const uint32_t addr = tmem + lane_field + ((uint32_t)(half * 16) << 16) + 64;
        const uint32_t a = 0x1000000u + (uint32_t)tid * 256u + half * 16u;
        asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};"
                     :: "r"(addr), "r"(a + 0), "r"(a + 1), "r"(a + 2), "r"(a + 3) : "memory");
    }
    tc_wait_st(); tc_fence_before(); __syncthreads(); tc_fence_after();
    {
        uint32_t r[16];
        tmem_ld16(tmem + lane_field + 64, r); tc_wait_ld();
        for (int i = 0; i < 8; ++i) out[(4 * 128 + tid) * 8 + i] = r[i];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    uint32_t* d; cudaMalloc(&d, 5 * 128 * 8 * 4); cudaMemset(d, 0xff, 5 * 128 * 8 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t h[5 * 128 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[4] = {"ld 16x256b.x2 lanes+0 col 8..", "ld 16x256b.x2 lanes+16 col 8..", "ld 16x128b.x2 lanes+0 col 8..", "ld 16x128b.x2 lanes+16 col 8.."};
    for (int s = 0; s < 4; ++s) {
        printf("== %s   (register -> lane:col), warps 0 and 1\n", names[s]);
        for (int t = 0; t < 64; ++t) {
            printf("T%-3d", t);
            for (int i = 0; i < (s < 2 ? 8 : 4); ++i) { uint32_t x = h[(s * 128 + t) * 8 + i]; printf("  r%d=%3u:%-2u", i, x >> 8, x & 255u); }
            printf("\n");
        }
    }
    printf("== st 16x128b.x2 at col 64 (value = 0x1000000 + writer_tid*256 + half*16 + reg), read by 32x32b: lane <- cols 64..71\n");
    for (int t = 0; t < 64; ++t) {
        printf("lane %-3d", t);
        for (int i = 0; i < 8; ++i) { uint32_t x = h[(4 * 128 + t) * 8 + i] - 0x1000000u; printf("  c%d=T%u/h%u/r%u", 64 + i, x >> 8, (x >> 4) & 15u, x & 15u); }
        printf("\n");
    }
    return 0;
}
