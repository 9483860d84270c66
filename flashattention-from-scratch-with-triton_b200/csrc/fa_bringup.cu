// Bring-up of the hand-encoded tcgen05 / TMA primitives: one 128-row tile
// GEMM per test, checked against a scalar host computation.  Descriptor fields are runtime
// parameters so one run can try candidate encodings; the attention kernels use the variant
// marked "expected".  Standalone binary: build/fa_bringup  (prints PASS/FAIL lines, exit 0 iff
// every "expected" variant passes).
#include "fa_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

using namespace fa;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return (EncodeTiledFn)fn;
}
// 2-D row-major [rows][cols] 16-bit tensor, box = [box_rows][64 cols], SWIZZLE_128B
static CUtensorMap make_map_2d(void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16) {
    static EncodeTiledFn enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
    return m;
}

// 2-D row-major [rows][cols] fp32 tensor, box = [box_rows][32 cols] (128 B), SWIZZLE_128B: the dQ accumulator of the
// fused backward, target of cp.reduce.async.bulk.tensor (.add)
static CUtensorMap make_map_2d_f32(void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static EncodeTiledFn enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(f32) failed %d\n", (int)r); exit(2); }
    return m;
}

struct GemmParams {
    int a_chunks, b_chunks;         // number of [rows x 64] boxes loaded for A and B
    int a_rows, b_rows;             // rows per box
    uint32_t a_lbo, a_sbo, a_kstep; // A smem descriptor fields (bytes), per-k-step start advance
    uint32_t a_chunk_ksteps;        // k-steps per A chunk before jumping to the next chunk (K-major)
    uint32_t b_lbo, b_sbo, b_kstep, b_chunk_ksteps, b_chunk_bytes, a_chunk_bytes;
    uint32_t idesc;
    int nk;                         // number of K=16 steps
    int n;                          // N (columns of D)
    int a_from_tmem;                // 1: A is written to TMEM by the threads (packed 16-bit pairs)
    int a_tmem_kstep_cols;          // TMEM column advance per k-step for A
    int bf16;
    int store_mode;                 // 0: D fp32 to global via registers; 1: also TMA-store a 16-bit copy;
                                    // 2: also reduce-add the fp32 tile twice into R_out through mapRed (TMA .add)
    // extra K = 16 step from two compact NO-SWIZZLE tiles written by the threads: A' = [128 x 16] with ones in columns 0-2,
    // B' = [n x 16] with a 3-way bf16 split of x[col] in columns 0-2  =>  D[m][col] += x[col]  (statistics through the MMA)
    int extra_kstep; uint32_t x_lbo, x_sbo;
};

// D[128 x n] = A[128 x K] * B   (B either [n x K] K-major or [K x n] MN-major, per descriptors)
__global__ void __launch_bounds__(128, 1)
bringup_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRed,
             const uint16_t* __restrict__ A_gmem, int a_ld,
             float* __restrict__ D_out, GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                     // up to 2 chunks x 16 KB
    uint8_t* sB = smem + 32768;             // up to 2 chunks x 16 KB
    uint8_t* sO = smem + 65536;             // 2 chunks x 16 KB staging for the TMA-store test
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        uint32_t bytes = 0;
        if (!p.a_from_tmem) bytes += p.a_chunks * p.a_rows * 128;
        bytes += p.b_chunks * p.b_rows * 128;
        mbar_arrive_expect_tx(&bar_load, bytes);
        if (!p.a_from_tmem)
            for (int c = 0; c < p.a_chunks; ++c) tma_load_2d(sA + c * p.a_chunk_bytes, &mapA, &bar_load, c * 64, 0);
        for (int c = 0; c < p.b_chunks; ++c) tma_load_2d(sB + c * p.b_chunk_bytes, &mapB, &bar_load, c * 64, 0);
    }
    if (p.a_from_tmem) {
        // thread r owns A row r: pack pairs (2j, 2j+1) -> TMEM column 256 + j of lane r
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int K = p.nk * 16;
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
            uint32_t v[16];
            #pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t lo = A_gmem[(size_t)tid * a_ld + 2 * (c0 + j)];
                uint32_t hi = A_gmem[(size_t)tid * a_ld + 2 * (c0 + j) + 1];
                v[j] = lo | (hi << 16);
            }
            tmem_st16(tmem + lane_base + 256 + c0, v);
        }
        tc_wait_st();
        tc_fence_before();
    }
    uint8_t* sXA = smem + 131072 - 8192;    // compact tiles of the extra k-step (inside the staging area, unused by these tests)
    uint8_t* sXB = sXA + 4096;
    if (p.extra_kstep) {
        // thread r writes row r of both tiles: core matrix (r/8, k-chunk j) at (r/8)*256 + j*128, row r%8 at +16 bytes each
        const uint32_t one = 0x3f80u;                               // bf16 1.0
        const float x = 0.37f * (float)tid - 11.5f + 1e-3f * (float)(tid * tid % 97);
        const __nv_bfloat16 hi = __float2bfloat16(x), mid = __float2bfloat16(x - __bfloat162float(hi));
        const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi) - __bfloat162float(mid));
        uint16_t h16, m16, l16; memcpy(&h16, &hi, 2); memcpy(&m16, &mid, 2); memcpy(&l16, &lo, 2);
        const uint32_t off = (tid >> 3) * 256 + (tid & 7) * 16;
        *reinterpret_cast<uint4*>(sXA + off) = make_uint4(one | (one << 16), one, 0u, 0u);
        *reinterpret_cast<uint4*>(sXA + off + 128) = make_uint4(0u, 0u, 0u, 0u);
        if (tid < p.n) {
            *reinterpret_cast<uint4*>(sXB + off) = make_uint4((uint32_t)h16 | ((uint32_t)m16 << 16), (uint32_t)l16, 0u, 0u);
            *reinterpret_cast<uint4*>(sXB + off + 128) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bar_load, 0, 1);
        tc_fence_after();
        for (int k = 0; k < p.nk; ++k) {
            uint32_t a_off = (k / p.a_chunk_ksteps) * p.a_chunk_bytes + (k % p.a_chunk_ksteps) * p.a_kstep;
            uint32_t b_off = (k / p.b_chunk_ksteps) * p.b_chunk_bytes + (k % p.b_chunk_ksteps) * p.b_kstep;
            uint64_t bd = make_smem_desc(smem_u32(sB) + b_off, p.b_lbo, p.b_sbo);
            if (p.a_from_tmem) {
                umma_ts(tmem, tmem + 256 + k * p.a_tmem_kstep_cols, bd, p.idesc, k > 0);
            } else {
                uint64_t ad = make_smem_desc(smem_u32(sA) + a_off, p.a_lbo, p.a_sbo);
                umma_ss(tmem, ad, bd, p.idesc, k > 0);
            }
        }
        if (p.extra_kstep)
            umma_ss(tmem, make_smem_desc_noswizzle(smem_u32(sXA), p.x_lbo, p.x_sbo),
                    make_smem_desc_noswizzle(smem_u32(sXB), p.x_lbo, p.x_sbo), p.idesc & ~((1u << 15) | (1u << 16)), 1);
        tc_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0, 2);
    tc_fence_after();
    // epilogue: thread r reads row r
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < p.n; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + c0, v);
        tc_wait_ld();
        #pragma unroll
        for (int j = 0; j < 32; ++j) D_out[(size_t)tid * p.n + c0 + j] = __uint_as_float(v[j]);
        if (p.store_mode == 2) {
            // fp32 staging, one [128 rows][32 floats] SWIZZLE_128B box per 32 columns
            #pragma unroll
            for (int g = 0; g < 8; ++g)
                sts128(smem_u32(sO) + (c0 / 32) * 16384 + sw128_offset(tid, g), v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if (p.store_mode == 1) {
            // stage a 16-bit copy in SWIZZLE_128B layout for the TMA store test
            int chunk = c0 / 64;
            #pragma unroll
            for (int g = 0; g < 4; ++g) {         // 4 x 16-byte groups = 32 elements
                uint32_t w[4];
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float lo = __uint_as_float(v[g * 8 + 2 * j]), hi = __uint_as_float(v[g * 8 + 2 * j + 1]);
                    w[j] = p.bf16 ? pack2<true>(lo, hi) : pack2<false>(lo, hi);
                }
                uint32_t c16 = (c0 % 64) / 8 + g;
                *reinterpret_cast<uint4*>(sO + chunk * 16384 + sw128_offset(tid, c16)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    if (p.store_mode == 1) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            for (int c = 0; c < p.n / 64; ++c) tma_store_2d(&mapOut, sO + c * 16384, c * 64, 0);
            tma_store_commit();
            tma_store_wait_all0();
        }
    }
    if (p.store_mode == 2) {
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            for (int rep = 0; rep < 2; ++rep) {
                for (int c = 0; c < p.n / 32; ++c) tma_reduce_add_2d(&mapRed, sO + c * 16384, c * 32, 0);
                tma_store_commit();
            }
            tma_store_wait_all0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static uint16_t f2h16(float f, bool bf16) {
    if (bf16) { __nv_bfloat16 b = __float2bfloat16(f); uint16_t u; memcpy(&u, &b, 2); return u; }
    __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u;
}
static float h2f16(uint16_t u, bool bf16) {
    if (bf16) { __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b); }
    __half h; memcpy(&h, &u, 2); return __half2float(h);
}

struct TestCfg {
    const char* name; bool expected;
    int K, N; bool b_mn_major; bool a_from_tmem; bool bf16; int store_mode;
    uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep; int a_tmem_kstep_cols;
    bool a_mn_major = false;        // A stored [K rows][128 M columns] (M contiguous): the fused backward's dS^T buffer
    int extra_kstep = 0; uint32_t x_lbo = 0, x_sbo = 0;
};

static bool run_test(const TestCfg& t) {
    const int M = 128, K = t.K, N = t.N;
    std::vector<uint16_t> hA((size_t)M * K), hB((size_t)(t.b_mn_major ? K * N : N * K));
    std::vector<float> fA(hA.size()), fB(hB.size());
    uint32_t s = 12345u + K * 7 + N * 3 + t.b_mn_major;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 9) & 0xff) - 128) / 64.0f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = f2h16(rnd(), t.bf16); fA[i] = h2f16(hA[i], t.bf16); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = f2h16(rnd(), t.bf16); fB[i] = h2f16(hB[i], t.bf16); }
    std::vector<float> ref((size_t)M * N, 0.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += (t.a_mn_major ? fA[(size_t)k * M + m] : fA[(size_t)m * K + k]) * (t.b_mn_major ? fB[(size_t)k * N + n] : fB[(size_t)n * K + k]);
        if (t.extra_kstep) acc += 0.37f * (float)n - 11.5f + 1e-3f * (float)(n * n % 97);
        ref[(size_t)m * N + n] = acc;
    }
    uint16_t *dA, *dB, *dO16; float *dD, *dR;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, (size_t)M * N * 4)); CK(cudaMalloc(&dO16, (size_t)M * N * 2));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)M * N * 4)); CK(cudaMemset(dO16, 0xff, (size_t)M * N * 2));
    CK(cudaMalloc(&dR, (size_t)M * N * 4));
    { std::vector<float> ones((size_t)M * N, 1.0f); CK(cudaMemcpy(dR, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice)); }

    GemmParams p; memset(&p, 0, sizeof(p));
    p.a_chunks = K / 64; p.a_rows = 128; p.a_chunk_bytes = 128 * 128; p.a_chunk_ksteps = 4;
    if (t.a_mn_major) { p.a_chunks = M / 64; p.a_rows = K; p.a_chunk_bytes = K * 128; p.a_chunk_ksteps = 1 << 20; }
    p.a_lbo = t.a_lbo; p.a_sbo = t.a_sbo; p.a_kstep = t.a_kstep;
    if (t.b_mn_major) { p.b_chunks = N / 64; p.b_rows = K; p.b_chunk_bytes = K * 128; p.b_chunk_ksteps = 1 << 20; }
    else              { p.b_chunks = K / 64; p.b_rows = N; p.b_chunk_bytes = N * 128; p.b_chunk_ksteps = 4; }
    p.b_lbo = t.b_lbo; p.b_sbo = t.b_sbo; p.b_kstep = t.b_kstep;
    p.idesc = make_idesc(t.bf16, t.a_mn_major, t.b_mn_major, 128, N);
    p.nk = K / 16; p.n = N; p.a_from_tmem = t.a_from_tmem; p.a_tmem_kstep_cols = t.a_tmem_kstep_cols;
    p.bf16 = t.bf16; p.store_mode = t.store_mode;
    p.extra_kstep = t.extra_kstep; p.x_lbo = t.x_lbo; p.x_sbo = t.x_sbo;

    CUtensorMap mA = t.a_mn_major ? make_map_2d(dA, K, M, K, t.bf16) : make_map_2d(dA, M, K, 128, t.bf16);
    CUtensorMap mR = make_map_2d_f32(dR, M, N, 128);
    CUtensorMap mB = t.b_mn_major ? make_map_2d(dB, K, N, K, t.bf16) : make_map_2d(dB, N, K, N, t.bf16);
    CUtensorMap mO = make_map_2d(dO16, M, N, 128, t.bf16);
    const int smem_bytes = 131072 + 1024;
    CK(cudaFuncSetAttribute(bringup_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    bringup_gemm<<<1, 128, smem_bytes>>>(mA, mB, mO, mR, dA, K, dD, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("FAIL %-40s launch error: %s\n", t.name, cudaGetErrorString(e));
        return false;   // context is likely dead after a trap; caller exits
    }
    std::vector<float> hD((size_t)M * N), hR((size_t)M * N); std::vector<uint16_t> hO((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hR.data(), dR, hR.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hO.data(), dO16, hO.size() * 2, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxerr16 = 0; int bad = 0, first_bad = -1;
    for (size_t i = 0; i < hD.size(); ++i) {
        double d = fabs((double)hD[i] - ref[i]);
        if (!(d <= 1e-3)) { if (first_bad < 0) first_bad = (int)i; ++bad; }
        if (d > maxerr || d != d) maxerr = d;
        if (t.store_mode == 2) {             // 1 + two reduce-adds of the tile
            double dr = fabs((double)hR[i] - (1.0 + 2.0 * ref[i]));
            if (!(dr <= 2e-3)) { if (first_bad < 0) first_bad = (int)i; ++bad; }
            if (dr > maxerr16) maxerr16 = dr;
        }
        if (t.store_mode == 1) {
            double d16 = fabs((double)h2f16(hO[i], t.bf16) - ref[i]);
            double tol = 0.02 * fabs(ref[i]) + 0.02;
            if (!(d16 <= tol)) { if (first_bad < 0) first_bad = (int)i; ++bad; }
            if (d16 > maxerr16) maxerr16 = d16;
        }
    }
    bool ok = bad == 0;
    printf("%s %-40s %s maxerr=%.3e maxerr16=%.3e bad=%d", ok ? "PASS" : "FAIL", t.name,
           t.expected ? "[expected]" : "[alt]     ", maxerr, maxerr16, bad);
    if (!ok) printf(" first_bad=(%d,%d) got=%f ref=%f", first_bad / N, first_bad % N, hD[first_bad], ref[first_bad]);
    printf("\n");
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dO16); cudaFree(dR);
    return ok;
}

int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
           prop.multiProcessorCount, (size_t)prop.sharedMemPerBlockOptin);
    // name, expected, K, N, b_mn, a_tmem, bf16, store, a_lbo,a_sbo,a_kstep, b_lbo,b_sbo,b_kstep, a_tmem_cols
    std::vector<TestCfg> tests = {
        {"ss_kmajor_K64_N128_bf16",        true,  64, 128, false, false, true,  0, 0, 1024, 32, 0, 1024, 32, 0},
        {"ss_kmajor_K128_N128_bf16",       true, 128, 128, false, false, true,  0, 0, 1024, 32, 0, 1024, 32, 0},
        {"ss_kmajor_K128_N128_fp16",       true, 128, 128, false, false, false, 0, 0, 1024, 32, 0, 1024, 32, 0},
        {"ss_kmajor_K64_N64_bf16",         true,  64,  64, false, false, true,  0, 0, 1024, 32, 0, 1024, 32, 0},
        {"ss_kmajor_lbo16_K128",           false,128, 128, false, false, true,  0, 16, 1024, 32, 16, 1024, 32, 0},
        {"ss_bmn_K128_N64_bf16",           true, 128,  64, true,  false, true,  0, 0, 1024, 32, 16384, 1024, 2048, 0},
        {"ss_bmn_K128_N128_bf16",          true, 128, 128, true,  false, true,  0, 0, 1024, 32, 16384, 1024, 2048, 0},
        {"ss_bmn_K128_N128_swapped_lbo_sbo", false,128,128, true, false, true,  0, 0, 1024, 32, 1024, 16384, 2048, 0},
        {"ss_bmn_K64_N128_bf16",           true,  64, 128, true,  false, true,  0, 0, 1024, 32, 8192, 1024, 2048, 0},
        {"ts_atmem_bmn_K128_N128_bf16",    true, 128, 128, true,  true,  true,  0, 0, 0, 0, 16384, 1024, 2048, 8},
        {"ts_atmem_bmn_K128_N64_fp16",     true, 128,  64, true,  true,  false, 0, 0, 0, 0, 16384, 1024, 2048, 8},
        {"ts_atmem_bmn_kstep16cols",       false,128, 128, true,  true,  true,  0, 0, 0, 0, 16384, 1024, 2048, 16},
        {"ts_atmem_bkmajor_K128_N128",     true, 128, 128, false, true,  true,  0, 0, 0, 0, 0, 1024, 32, 8},
        {"tma_store_sw128_N128_bf16",      true, 128, 128, false, false, true,  1, 0, 1024, 32, 0, 1024, 32, 0},
        {"tma_store_sw128_N64_fp16",       true,  64,  64, false, false, false, 1, 0, 1024, 32, 0, 1024, 32, 0},
        // fused backward: dQ[128 q x 64] = dS (A MN-major: smem holds dS^T [kv rows][128 q], two 64-column chunks)
        //                                  * K (B MN-major [kv rows][64 d]); fp32 tile reduce-added by TMA
        {"ss_amn_bmn_K128_N64_bf16",       true, 128,  64, true,  false, true,  0, 16384, 1024, 2048, 16384, 1024, 2048, 0, true},
        {"ss_amn_bmn_K128_N64_fp16",       true, 128,  64, true,  false, false, 0, 16384, 1024, 2048, 16384, 1024, 2048, 0, true},
        {"ss_amn_bkmajor_K64_N128_bf16",   true,  64, 128, false, false, true,  0, 8192, 1024, 2048, 0, 1024, 32, 0, true},
        {"ss_amn_swapped_lbo_sbo",         false,128,  64, true,  false, true,  0, 1024, 16384, 2048, 16384, 1024, 2048, 0, true},
        // statistics through the MMA: one extra K = 16 step from compact no-swizzle tiles (two candidate LBO / SBO readings)
        {"ss_extra_kstep_noswizzle_lbo128_sbo256", true, 64, 128, false, false, true, 0, 0, 1024, 32, 0, 1024, 32, 0, false, 1, 128, 256},
        {"ss_extra_kstep_noswizzle_lbo256_sbo128", false, 64, 128, false, false, true, 0, 0, 1024, 32, 0, 1024, 32, 0, false, 1, 256, 128},
        {"tma_reduce_add_f32_sw128_N64",   true, 128,  64, true,  false, true,  2, 0, 1024, 32, 16384, 1024, 2048, 0},
        {"tma_reduce_add_f32_sw128_N128",  true, 128, 128, false, false, true,  2, 0, 1024, 32, 0, 1024, 32, 0},
    };
    int fails = 0;
    for (size_t i = 0; i < tests.size(); ++i) {
        if (only >= 0 && (int)i != only) continue;
        bool ok = run_test(tests[i]);
        if (!ok && tests[i].expected) ++fails;
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            printf("context lost after test %zu; stopping (rerun with index to isolate)\n", i);
            return 3;
        }
    }
    printf("bringup: %d expected-variant failures\n", fails);
    return fails ? 1 : 0;
}
