// Hand-written PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld / st / fences) and the UMMA shared-memory / instruction
// descriptors.  No CUTLASS/CuTe: every encoding below is spelled out and is exercised by
// csrc/fa_bringup.cu (unit GEMMs against a scalar check) before any attention kernel relies on it.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <utility>

namespace fa {

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// ordered variant: stays between the named barriers that bracket a warpgroup's turn on the MUFU unit
__device__ __forceinline__ float ex2_approx_ordered(float x) {
    float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {          // FMNMX3 (sm_100)
    float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

// Debug record written by a waiter that gave up (see mbar_wait): {tag, blockIdx.x, threadIdx.x, parity}
struct HangRecord { unsigned int tag, block, thread, parity; };
static __device__ HangRecord g_hang_record;
static __device__ unsigned int g_hang_flag;

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA store, tcgen05.mma smem reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box.  After FA_WAIT_TIMEOUT_NS without progress the first waiter records who
// it is in g_hang_record / g_hang_flag; every waiter then abandons its wait, the kernel drains and TRAPS at its end
// (hang_trap_if_set): the launch fails with a sticky CUDA error, so this and every later call of the process report it instead of
// handing back garbage.  The timeout is wall clock (globaltimer keeps running under time-slicing or a debugger), hence generous.
// Development builds (-DFA_HANG_TRAP=0) do not trap: the kernel finishes with garbage results and fa_sm100_last_hang() returns the
// record.  (The trap sits at the kernel's end, not in the wait loop: inside the loop it changes the register allocation of every
// hot loop that waits on a barrier.)
#ifndef FA_WAIT_TIMEOUT_NS
#define FA_WAIT_TIMEOUT_NS 20000000000ull
#endif
#ifndef FA_HANG_TRAP
#define FA_HANG_TRAP 1
#endif
__device__ __forceinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, uint32_t tag) {
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0) {
            if (*reinterpret_cast<volatile unsigned int*>(&g_hang_flag)) return;
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > FA_WAIT_TIMEOUT_NS) {
                if (atomicExch(&g_hang_flag, 1u) == 0u) {
                    g_hang_record.tag = tag; g_hang_record.block = blockIdx.x;
                    g_hang_record.thread = threadIdx.x; g_hang_record.parity = parity;
                    __threadfence();
                }
                return;
            }
        }
    }
}
// last statement of every kernel that waits on mbarriers
__device__ __forceinline__ void hang_trap_if_set() {
    if (FA_HANG_TRAP && threadIdx.x == 0 && *reinterpret_cast<volatile unsigned int*>(&g_hang_flag)) __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0) {
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity, tag);
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src,
                                             int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// Reduce-add of a shared-memory tile into global memory through a tensor map (element type from the map: fp32 for the
// fused backward's dQ accumulator).  Same bulk-group completion as the stores; rows outside the tensor are dropped.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// smem of all committed store groups has been read (safe to overwrite the staging buffer)
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed store groups are complete (globally visible)
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 1-D bulk copy global -> shared (rows of fp32 statistics), completes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit
// ----------------------------------------------------------------------------------------------
// One full warp calls this; the TMEM base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive(1) on `bar` once every tcgen05 op this thread issued so far has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit):
//   [0,14)  start address >> 4         [16,30) leading-dim byte offset >> 4 (LBO)
//   [32,46) stride-dim byte offset >> 4 (SBO)      [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0            [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
// All our tiles are [rows][64 x 16-bit] = 128-byte rows written by TMA with SWIZZLE_128B, so one
// swizzle atom is 8 rows x 128 B = 1024 B and consecutive 8-row groups are 1024 B apart (SBO).
//  * K-major operand (contraction dim contiguous):  the MMA's K=16 slice is 32 B inside the
//    128-B row -> advance the start address by 32 B per k-step; LBO unused (one atom along K).
//  * MN-major operand (M/N dim contiguous, contraction dim = rows): the K=16 slice is 16 rows
//    = 2048 B -> advance by 2048 B per k-step; LBO = byte distance between 64-element column
//    chunks (used when N = 128 spans two chunks).
constexpr uint64_t kDescSw128 = (uint64_t(2) << 61) | (uint64_t(1) << 46);
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return kDescSw128 | (uint64_t((sbo_bytes >> 4) & 0x3fff) << 32) |
           (uint64_t((lbo_bytes >> 4) & 0x3fff) << 16) | uint64_t((saddr >> 4) & 0x3fff);
}
// Same descriptor without swizzle (bits 61-63 = 0): compact [rows][16 x 16-bit] K-major operands made of 8-row x 16-byte core
// matrices (128 contiguous bytes each).  Candidate use: one extra K = 16 step that adds per-column statistics to a score tile
// inside the MMA; encoding pinned by fa_bringup "ss_extra_kstep_noswizzle_*".
__device__ __forceinline__ uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t(1) << 46) | (uint64_t((sbo_bytes >> 4) & 0x3fff) << 32) |
           (uint64_t((lbo_bytes >> 4) & 0x3fff) << 16) | uint64_t((saddr >> 4) & 0x3fff);
}
// Instruction descriptor, kind::f16 (32-bit):
//   [4,6) D format: 1 = f32   [7,10) A format: 0 = f16, 1 = bf16   [10,13) B format
//   [15] A major: 0 = K, 1 = MN   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, bool a_mn_major, bool b_mn_major,
                                                  uint32_t M, uint32_t N) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) |
           ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: 128 lanes x K 16-bit elements packed 2 per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld / st, shape 32x32b: thread t of the warp <-> TMEM lane (warp%4)*32 + t,
// N consecutive 32-bit columns per thread.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr),
           "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// 16-bit packing: element 2j -> low half, 2j+1 -> high half (memory / K order)
// ----------------------------------------------------------------------------------------------
template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (kBf16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    } else {
        __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
}

// Byte offset of 16-byte chunk `c16` (0..7) of row `row` inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t c16) {
    return row * 128u + ((c16 ^ (row & 7u)) << 4);
}

// explicit shared-window accesses (pointers derived from the aligned dynamic-smem base are otherwise
// compiled as generic LD/ST)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t saddr, float a) {
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(saddr), "f"(a) : "memory");
}

// packed fp32x2 math (FFMA2 / FADD2 / FMUL2 on sm_100): two elements per issue slot
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ uint64_t pack_u2(uint32_t lo, uint32_t hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// exp2 of a pair on the FMA pipe (MUFU relief, the FlashAttention-4 trick): x = n + f with n = round(x),
// f in [-0.5, 0.5]; 2^f by a degree-3 near-minimax polynomial (max rel. error 7.5e-5, far below the 16-bit
// rounding of P); 2^n spliced into the exponent field.  x is clamped at -126 so the splice never wraps.
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& p0, float& p1) {
    float x0, x1; unpack_f2(x, x0, x1);
    const uint64_t xm = pack_f2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    const uint64_t magic = pack_f2(12582912.f, 12582912.f);              // 1.5 * 2^23: low mantissa bits = round(x)
    const uint64_t t = fadd2(xm, magic);
    const uint64_t nf = fadd2(t, pack_f2(-12582912.f, -12582912.f));
    const uint64_t f = ffma2(nf, pack_f2(-1.f, -1.f), xm);
    uint64_t p = ffma2(pack_f2(0.055179596f, 0.055179596f), f, pack_f2(0.24261186f, 0.24261186f));
    p = ffma2(p, f, pack_f2(0.69325954f, 0.69325954f));
    p = ffma2(p, f, pack_f2(0.99992800f, 0.99992800f));
    float t0, t1, q0, q1; unpack_f2(t, t0, t1); unpack_f2(p, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// One leader lane of a CONVERGED warp.  The MMA / TMA warps keep all 32 lanes in uniform control flow and elect
// only around the tcgen05 / TMA / arrive instructions: descriptors computed in converged code stay in uniform
// registers, whereas code inside a divergent `if (lane == 0)` makes ptxas wrap every UTCHMMA / UTMALDG in an
// ELECT + R2UR.BROADCAST loop (~13 extra instructions per MMA on the latency-critical issue path).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Self-electing variants: every lane of the converged warp calls them, one leader lane executes the instruction.
#define FA_ELECT_PROLOGUE "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
__device__ __forceinline__ void umma_ss_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(FA_ELECT_PROLOGUE ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(FA_ELECT_PROLOGUE ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint64_t* bar) {
    asm volatile(FA_ELECT_PROLOGUE "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_e(uint64_t* bar) {
    asm volatile(FA_ELECT_PROLOGUE "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_e(uint64_t* bar, uint32_t bytes) {
    asm volatile(FA_ELECT_PROLOGUE "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(FA_ELECT_PROLOGUE
                 "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d_e(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(FA_ELECT_PROLOGUE
                 "@e cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Dropout (the tutorial's own next step, Phase_6.md:54-114): a counter-based generator keyed by (seed, batch*head, query row,
// key column), so the forward and both backward kernels regenerate the same keep mask in whatever orientation they hold the
// score tile.  One 32-bit word covers the 4 key columns k & ~3 .. (k | 3) of one query row, one byte per element;
// keep <=> byte >= thresh, i.e. p_drop = thresh / 256 exactly and kept elements are scaled by 256 / (256 - thresh).
// mix32 is the "lowbias32" integer finaliser.  The test suite pins this definition with a numpy restatement (dropout_keep_mask).
// ----------------------------------------------------------------------------------------------
struct DropoutParams { uint32_t seed0, seed1, thresh; float scale; };
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_row_key(uint32_t seed0, uint32_t bh, uint32_t q) {
    return mix32((bh * 0x9E3779B1u + q) ^ seed0);
}
// the 4 random bytes of key columns 4*k4 .. 4*k4+3 of the row with key `row_key`
__host__ __device__ __forceinline__ uint32_t dropout_word(uint32_t row_key, uint32_t seed1, uint32_t k4) {
    return mix32(row_key ^ (k4 * 0x85EBCA6Bu + seed1));
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t word, uint32_t k, uint32_t thresh) {
    return ((word >> ((k & 3u) * 8u)) & 0xffu) >= thresh;
}

// Work order of the persistent kernels.  Heads are cut into chunks of `hc` heads whose tensors fit L2 together; inside a
// chunk items go tile-major (t = 0 is the heaviest tile under causal), so the dynamic scheduler hands out heavy items first
// and the last items of a launch are the lightest.  (A plain head-major order leaves one of the heaviest items of the last
// head for the very end: a tail of +15..50 % on short causal problems.)
__device__ __forceinline__ void item_to_head_tile(int item, int n_heads, int n_tiles, int hc, int& head, int& t) {
    const int per_chunk = hc * n_tiles;
    const int chunk = item / per_chunk, rem = item - chunk * per_chunk;
    const int h0 = chunk * hc;
    const int nh = min(hc, n_heads - h0);
    t = rem / nh; head = h0 + rem - t * nh;
}

// The work counters of the persistent kernels reset themselves, so no memset has to precede a launch: word 0 hands out items,
// word 1 counts the CTAs whose scheduler warp has drawn its last item (>= n_items); the last of them zeroes both.  Every fetch on
// word 0 has returned by then, and the next user of the slot is a later launch.
__device__ __forceinline__ void sched_retire(unsigned int* sched) {
    if (atomicAdd(sched + 1, 1u) == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
}

// First item of a persistent CTA.  Static (item = blockIdx.x, no atomic in front of the first TMA load) when the launch owns the
// GPU; DRAWN FROM THE COUNTER like every later item when the SMs are shared (dyn != 0: sequence-parallel runs, where NCCL kernels
// occupy a few SMs while the attention launch starts).  A CTA that cannot be scheduled until an SM frees would otherwise hold its
// static item — one of the HEAVIEST, the order is heavy-first — hostage until the very end of the launch: measured +15..20 % on
// every launch that overlapped a transfer (profiles/r02, C5 at 2 GPUs).  Whole warp, converged.
__device__ __forceinline__ int sched_first(unsigned int* sched, int dyn) {
    int item = (int)blockIdx.x;
    if (dyn) {
        if (lane_id() == 0) item = (int)atomicAdd(sched, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
    }
    return item;
}
__device__ __forceinline__ int sched_next(unsigned int* sched, int dyn) {
    int item = 0;
    if (lane_id() == 0) item = (int)atomicAdd(sched, 1u) + (dyn ? 0 : (int)gridDim.x);
    return __shfl_sync(0xffffffffu, item, 0);
}

// Programmatic dependent launch.  Every kernel of the library is launched with the programmatic-stream-serialization attribute
// (launch_pdl) and executes pdl_wait() before its first global-memory access: the launch and the prologue (barrier init, TMEM
// allocation, tensor-map prefetch) overlap the tail of the previous kernel in the stream; pdl_wait returns once that kernel has
// completed and its writes are visible.  A short step (C2: fwd, delta, fused, convert = 0.2 ms) otherwise pays a launch gap per boundary.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef FA_PDL
#define FA_PDL 1
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = FA_PDL ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// The opt-in to > 48 KB of dynamic shared memory is a per-device (per-context) attribute of the kernel: set it once per device
// ordinal, not once per process, or the first launch on a second GPU of the same process fails with "invalid value".
template <auto Kernel>
inline cudaError_t ensure_smem(int bytes, int dev) {
    static std::atomic<unsigned long long> done{0};           // one bit per device ordinal (< 64, checked by the C API)
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// register re-distribution between warpgroups (all 4 warps of a warpgroup must execute it)
template <int N> __device__ __forceinline__ void reg_alloc()   { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }

// named barrier over a subset of warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

}  // namespace fa
