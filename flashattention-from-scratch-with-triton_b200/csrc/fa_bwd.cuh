// Backward attention for sm_100a.  Replaces flash_attention_dQ_kernel and flash_attention_dKV_kernel
// (reference code/_flash_attention_kernel_optimized.py:164-258 and :291-386); delta comes from the
// preprocess kernel in fa_aux.cuh.  Same two-kernel, atomic-free (deterministic) structure as the
// reference; both kernels recompute P from the saved LSE.
//
// Common CTA shape: 384 threads
//   warps 0-3  compute warpgroup A : TMEM lane r = tid & 127, columns [0,64) of the 128-wide score tile
//   warps 4-7  compute warpgroup B : same lanes, columns [64,128)
//   warp  8    MMA issuer (one thread)      warp 9  TMA producer (one thread)
//   warp  10   statistics loader (dKV only: per-column -LSE*log2e and delta -> smem)   warp 11 idle
//
// dKV kernel (one CTA per 128-row K/V tile, loops over Q tiles i), transposed so kv rows are lanes:
//   S^T = K Q_i^T, dP^T = V dO_i^T  (SS MMAs, N = 128)        TMEM: S^T [0,128) dP^T [128,256)
//   P^T = exp2(S^T c - LSE_i log2e) -> 16-bit over S^T; dV += P^T dO_i    (A from TMEM, B = dO_i MN-major)
//   dS^T = P^T o (dP^T - delta_i)   -> 16-bit over dP^T; dK += dS^T Q_i   TMEM: dV [256,256+D) dK [256+D,..)
// dQ kernel (one CTA per 128-row Q tile, loops over K/V tiles j):
//   S = Q K_j^T, dP = dO V_j^T; dS = P o (dP - delta) -> 16-bit into its OWN double-buffered TMEM region;
//   dQ += dS K_j (B = K_j MN-major).  TMEM: S [0,128) dP [128,256) dQ [256,256+D) dS0 dS1 (64 cols each)
//   = exactly 512 columns at D=128.  S and dP are released as soon as they are in registers, so S(j+1) and
//   dP(j+1) run under the exp/dS math of tile j and the math warps never wait for the tensor pipe.
// Each warpgroup packs its 64 columns into the first 32 TMEM columns of ITS OWN half of the region
// it overwrites, so the two warpgroups never touch each other's unread data; the MMA issuer
// addresses the two 32-column runs explicitly per K=16 step.
#pragma once
#include "fa_ptx.cuh"
#include "fa_fwd.cuh"   // fwd_tile_iters

namespace fa {

struct BwdParams {
    int BH, H, Sq, Sk, causal;     // 4-D tensor maps [B, H, S, D]: coordinates (col, row, h, b)
    int Hk, G;                     // K/V heads and query heads per K/V head (GQA/MQA; G = 1: reference layout)
    float scale, scale_log2;
    const float* lse;      // [BH, Sq]
    const float* delta;    // [BH, Sq]
    int n_qtiles, n_ktiles;
    unsigned int* sched_dkv;   // work counters (zeroed before launch) for the persistent kernels
    unsigned int* sched_dq;
    int sms;
    int dyn_first;             // draw the first item from the counter too (shared SMs, see sched_first)
    int dev;                   // device ordinal of the launch (host side only: per-device kernel attributes)
    int hc_dkv, hc_dq;         // heads per scheduling chunk (item_to_head_tile) for the K/V-tile and the Q-tile kernels
    // Optional range masks (FwdParams): row_lo/row_hi [B, Sq] = keys visible to a query row (used by the dQ kernel);
    // col_lo/col_hi [B, Sk] = queries that see a key row (the same mask seen from the K/V side, used by the dK/dV kernel).
    // All four non-decreasing along the sequence; NULL = unrestricted.
    const int* row_lo; const int* row_hi;
    const int* col_lo; const int* col_hi;
    DropoutParams drop;        // thresh = 0: no dropout (kDropout instantiations only)
};

// Turn-taking between the two math warpgroups around the exp loop (named barriers 3/4, as in the forward):
// while one warpgroup is on the MUFU unit the other runs its FMA/LDS-heavy dS phase.
#ifndef FA_BWD_STAGGER
#define FA_BWD_STAGGER 0   // measured: no gain (both warpgroups share one tile, the MMA wait dominates); kept as a knob
#endif
// Of every 16 score columns, this many (0, 4, 8) take their exp2 from the FMA-pipe polynomial (ex2_poly2) instead of MUFU: the
// exp phase of both backward loops queues on the MUFU unit (ncu: stall_mio) while the FMA pipe idles.  Measured A/B
// (profiles/r01_ab_bwd_poly.txt): 4 is best at D = 64 (-5..-7 %), 8 for the dK/dV kernel at D = 128 (-3..-4 %), dQ at 128 is flat.
#ifdef FA_BWD_POLY
template <int D> struct BwdPoly { static constexpr int kDkv = FA_BWD_POLY, kDq = FA_BWD_POLY; };
#else
template <int D> struct BwdPoly { static constexpr int kDkv = (D == 128) ? 8 : 4, kDq = 4; };
#endif
constexpr int kBwdThreads = 384;
// setmaxnreg split (2 math warpgroups + 1 warpgroup of MMA / TMA / statistics warps) of the CTA's 384 x 168 registers:
// 2 * compute + other <= 504, or setmaxnreg.inc never returns
template <int D> struct BwdRegs { static constexpr int kCompute = 208, kOther = 80; };
static_assert(2 * BwdRegs<64>::kCompute + BwdRegs<64>::kOther <= 504 && 2 * BwdRegs<128>::kCompute + BwdRegs<128>::kOther <= 504, "register pool");
constexpr float kLog2e = 1.44269504088896340736f;

template <int D> struct BwdCfg {
    static constexpr int kChunks = D / 64;
    static constexpr int kTileBytes = 128 * D * 2;
    static constexpr int kStages = (D == 128) ? 2 : 4;
    // resident pair (K,V for dKV; Q,dO for dQ) + kStages x streamed pair (+ 1 KB statistics per stage)
    static constexpr int kOffRes = 0;
    static constexpr int kOffStage = 2 * kTileBytes;
    static constexpr int kStageBytes = 2 * kTileBytes;
    static constexpr int kStatStages = 8;                 // deep ring: statistics are tiny and latency-bound
    static constexpr int kOffStat = kOffStage + kStages * kStageBytes;
    static constexpr int kOffBar = kOffStat + kStatStages * 1024;
    static constexpr int kNumBars = 8 + 3 * kStages + 2 * kStatStages + 4;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
};

// score-tile MMA: D[tmem 128x128] = A[smem 128 x D, K-major] * B[smem 128 x D, K-major]^T
template <int D, bool kBf16>
__device__ __forceinline__ void issue_scores(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr) {
    constexpr uint32_t idesc = make_idesc(kBf16, false, false, 128, 128);
    #pragma unroll
    for (int k = 0; k < D / 16; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss_e(d_tmem, make_smem_desc(a_addr + off, 0, 1024), make_smem_desc(b_addr + off, 0, 1024), idesc, k > 0);
    }
}
// gradient MMA: D[tmem 128 x D] (+)= A[tmem: 128 lanes x 128 16-bit, two 32-column runs at a_tmem and
// a_tmem+64] * B[smem 128 rows x D, MN-major]
// (kSplitA) or one contiguous 64-column run (!kSplitA)
template <int D, bool kBf16, bool kSplitA = true>
__device__ __forceinline__ void issue_grad(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, bool acc) {
    constexpr uint32_t idesc = make_idesc(kBf16, false, true, 128, D);
    #pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_ts_e(d_tmem, a_tmem + (kSplitA ? (k >> 2) * 64 + (k & 3) * 8 : k * 8),
                make_smem_desc(b_addr + k * 2048, 16384, 1024), idesc, acc || k > 0);
}

// TMEM accumulator half-row (D/2 columns starting at h*D/2) -> scaled 16-bit -> SWIZZLE_128B staging tile
template <int D, bool kBf16>
__device__ __forceinline__ void stage_grad_half(uint32_t t_acc, uint8_t* stage, int r, int h, float mul, bool zero) {
    constexpr int kCols = D / 2;
    #pragma unroll
    for (int q = 0; q < kCols / 32; ++q) {
        uint32_t v[32];
        if (!zero) { tmem_ld32(t_acc + h * kCols + q * 32, v); tc_wait_ld(); }
        else {
            #pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        const int col0 = h * kCols + q * 32;             // first of 32 columns held in v
        uint8_t* chunk = stage + (col0 >> 6) * 16384;
        #pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                w[i] = pack2<kBf16>(__uint_as_float(v[g * 8 + 2 * i]) * mul, __uint_as_float(v[g * 8 + 2 * i + 1]) * mul);
            sts128(smem_u32(chunk) + sw128_offset(r, ((col0 & 63) >> 3) + g), w[0], w[1], w[2], w[3]);
        }
    }
}

// =================================================================================================
// dK / dV  — persistent: one CTA per SM walks (b,h, kv-tile) items from a dynamic scheduler; the next item's
// K/V/Q/dO loads, statistics and first score MMAs run under the current item's epilogue.
// =================================================================================================
#ifndef FA_BWD_PERSISTENT
#define FA_BWD_PERSISTENT 1     // 0: grid = one CTA per item through the same code (A/B switch)
#endif

// Statistics through the MMA: instead of 32 broadcast LDS.128 per thread and tile (30 % of the shared-memory
// pipe in the D=128 dK/dV kernel), -LSE/scale and -delta ride the two score MMAs as ONE extra K = 16 step:
//   S'^T = K Q^T + ones (x) a^T,  a_q = 3-way bf16 split of -LSE_q / scale;   dP'^T = V dO^T + ones (x) b^T,  b_q = split of -delta_q
// from compact no-swizzle tiles (encoding and precision pinned by fa_bringup ss_extra_kstep_noswizzle_*: max error 3.8e-6).
// The math warps then compute P = exp2(c S') and dS = P o dP' with no shared-memory reads.  Off until measured.
#ifndef FA_STATS_MMA
#define FA_STATS_MMA 0
#endif
constexpr bool kStatsMma = FA_STATS_MMA != 0;

template <int D> struct DkvCfg {
    static constexpr int kChunks = D / 64;
    static constexpr int kTileBytes = 128 * D * 2;
    static constexpr int kStages = (D == 128) ? 2 : 4;
    // plain: 8 slots x 1 KB (128 fp32 + 128 fp32); through the MMA: 3 slots x 2 compact [128 x 16] bf16 tiles + the ones tile
    static constexpr int kStatStages = kStatsMma ? 3 : 8;
    static constexpr int kStatSlotBytes = kStatsMma ? 8192 : 1024;
    static constexpr int kStatBytes = kStatStages * kStatSlotBytes + (kStatsMma ? 4096 : 0);
    static constexpr bool kSepStage = (D == 64);          // own dK/dV staging: next K/V can land under the epilogue
    static constexpr int kOffRes = 0;
    static constexpr int kOffStage = 2 * kTileBytes;
    static constexpr int kStageBytes = 2 * kTileBytes;
    static constexpr int kOffStat = kOffStage + kStages * kStageBytes;
    static constexpr int kOffOut = kOffStat + kStatBytes;
    static constexpr int kOffBar = kOffOut + (kSepStage ? 2 * kTileBytes : 0);
    static constexpr int kNumBars = 16 + 3 * kStages + 2 * kStatStages;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 32 + 1024;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// kDropout: instantiation that regenerates the forward's keep mask (fa_ptx.cuh): dV uses the dropped-out, rescaled P^T,
// dP is masked and rescaled before dS = P o (dP - delta).  The plain instantiation is untouched.
template <int D, bool kBf16, bool kDropout = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                  const __grid_constant__ CUtensorMap mapdK, const __grid_constant__ CUtensorMap mapdV,
                  const BwdParams p) {
    using C = DkvCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sK = smem + C::kOffRes;
    uint8_t* sV = sK + C::kTileBytes;
    uint8_t* sStage = smem + C::kOffStage;             // per stage: Q_i then dO_i
    float* sStat = reinterpret_cast<float*>(smem + C::kOffStat);   // per slot: 128 x (-LSE*log2e), 128 x (-delta)
    uint8_t* sOutV = C::kSepStage ? smem + C::kOffOut : sV;        // dV / dK staging for the TMA store
    uint8_t* sOutK = C::kSepStage ? smem + C::kOffOut + C::kTileBytes : sK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* k_full = bars;            uint64_t* v_full = bars + 1;
    uint64_t* s_full = bars + 2;        uint64_t* dp_full = bars + 3;
    uint64_t* p_full = bars + 4;        uint64_t* ds_full = bars + 5;
    uint64_t* acc_full = bars + 6;      uint64_t* s_full1 = bars + 7;    // second S^T buffer (kDoubleS)
    uint64_t* acc_empty = bars + 8;     uint64_t* kv_free = bars + 9;
    uint64_t* sched_full = bars + 10;   uint64_t* sched_empty = bars + 12;   // [2] each
    uint64_t* q_full = bars + 16;                       // [kStages]
    uint64_t* do_full = q_full + C::kStages;            // [kStages]
    uint64_t* stage_empty = do_full + C::kStages;       // [kStages]
    uint64_t* stat_full = stage_empty + C::kStages;     // [kStatStages]
    uint64_t* stat_empty = stat_full + C::kStatStages;  // [kStatStages]
    volatile int* sched_item = reinterpret_cast<volatile int*>(bars + C::kNumBars);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_items = (p.BH / p.G) * p.n_ktiles;       // one item per (batch, K/V head, kv tile)

    if (tid == 0) {
        mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(s_full1, 1); mbar_init(dp_full, 1);
        mbar_init(p_full, 256); mbar_init(ds_full, 256); mbar_init(acc_full, 1); mbar_init(acc_empty, 256);
        mbar_init(kv_free, C::kSepStage ? 1 : 2);         // MMA thread (+ the store's read-done when staging aliases K/V)
        for (int i = 0; i < 2; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 10); }
        for (int i = 0; i < C::kStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&do_full[i], 1); mbar_init(&stage_empty[i], 1); }
        for (int i = 0; i < C::kStatStages; ++i) { mbar_init(&stat_full[i], 1); mbar_init(&stat_empty[i], kStatsMma ? 1 : 8); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    // D=64 leaves 128 TMEM columns free: S^T is double-buffered there, so S^T(i+1) is ready before the math
    // of tile i ends (the S -> P -> dV -> S chain otherwise dominates when the MMAs are short).
    constexpr bool kDoubleS = (D == 64);
    constexpr uint32_t kColST = 0, kColDPT = kDoubleS ? 256 : 128, kColDV = kDoubleS ? 384 : 256, kColDK = kColDV + D;

    // item -> (batch*Hk + kv head, kv tile, first q tile, q tiles per query head, iterations); ascending kv tile =
    // heavy first under causal.  With GQA the item walks the q tiles of every query head of the group, so the
    // reduction of dK/dV over the group happens in the TMEM accumulators (deterministic, no atomics).
    auto decode = [&](int item, int& bh, int& jt, int& i_start, int& i_end, int& n_it) {
        item_to_head_tile(item, p.BH / p.G, p.n_ktiles, p.hc_dkv, bh, jt);     // bh = b * Hk + hk
        i_start = p.causal ? jt : 0;                       // first Q tile with a row >= kv_block_start (:341)
        i_end = p.n_qtiles;
        if (p.col_lo) {                                    // range mask: q tiles [first query of the first kv row, last query of the last)
            const size_t cb = (size_t)(bh / p.Hk) * p.Sk;
            i_start = max(i_start, __ldg(p.col_lo + cb + min(jt * 128, p.Sk - 1)) >> 7);
            i_end = min(i_end, ((min(__ldg(p.col_hi + cb + min(jt * 128 + 127, p.Sk - 1)), p.Sq) - 1) >> 7) + 1);
        }
        n_it = max(i_end - i_start, 0) * p.G;
    };
    // consumer side of the scheduler broadcast: a whole warp calls it (lane 0 releases the slot), or one thread alone
    auto next_item = [&](uint32_t ix, bool solo = false) -> int {
        const uint32_t slot = ix & 1;
        mbar_wait(&sched_full[slot], (ix >> 1) & 1, 440);
        int item = sched_item[slot];
        if (!solo) item = __shfl_sync(0xffffffffu, item, 0);      // provably warp-uniform -> uniform registers downstream
        if (solo) mbar_arrive(&sched_empty[slot]); else mbar_arrive_e(&sched_empty[slot]);
        return item;
    };

    if (warp == 11) {
        reg_dealloc<BwdRegs<D>::kOther>();
    } else if (warp == 10) {
        // ------------------------------ statistics loader ------------------------------
        // Runs up to kStatStages tiles ahead of the math; the global loads of tile it+1 are in flight
        // while tile it waits for its slot, so their latency never reaches the critical path.
        reg_dealloc<BwdRegs<D>::kOther>();
        const int lane = lane_id();
        const uint32_t stat_addr = smem_u32(sStat);
        uint32_t gs = 0;
        if constexpr (kStatsMma) {
            // once: zero every compact tile (their second 16-byte K chunk is never written again) and build the ones tile
            // [128 x 16] = {1, 1, 1, 0 ...} behind the slots; the first stat_full arrive publishes all of it
            for (uint32_t o = lane * 16; o < (uint32_t)C::kStatBytes; o += 32 * 16) sts128(stat_addr + o, 0u, 0u, 0u, 0u);
            __syncwarp();
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t q = lane * 4 + u;
                sts128(stat_addr + C::kStatStages * C::kStatSlotBytes + (q >> 3) * 256 + (q & 7) * 16, 0x3f803f80u, 0x3f80u, 0u, 0u);
            }
            fence_proxy_async_smem();
            __syncwarp();
        }
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            // (q tile, row offset of the query head) walker over the item's iterations; with GQA it visits every head of the
            // group.  Lane l owns rows 4l..4l+3 of the tile: one 16-byte load each for LSE and delta (scalar, bounds-checked
            // loads only for a ragged last tile or an unaligned S_q) and one 16-byte shared store each.
            int s_qtile = i_start;
            size_t s_row0 = ((size_t)(bh / p.Hk) * p.H + (size_t)(bh % p.Hk) * p.G) * p.Sq;
            const bool vec_ok = (p.Sq & 3) == 0;
            auto fetch = [&](float4& nl, float4& dl) {             // statistics of the walker's current tile, then advance
                const int q0 = s_qtile * 128 + lane * 4;
                const size_t off = s_row0 + q0;
                if (++s_qtile == i_end) { s_qtile = i_start; s_row0 += p.Sq; }
                float l[4];
                if (vec_ok && q0 + 4 <= p.Sq) {
                    const float4 lv = __ldg(reinterpret_cast<const float4*>(p.lse + off));
                    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.delta + off));
                    dl = make_float4(-dv.x, -dv.y, -dv.z, -dv.w);          // the math warps add -delta (no negation in their loop)
                    l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
                } else {
                    float d[4];
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool in = q0 + u < p.Sq;                 // out-of-range query rows: P = exp2(-inf) = 0
                        l[u] = in ? __ldg(p.lse + off + u) : INFINITY;
                        d[u] = in ? -__ldg(p.delta + off + u) : 0.f;
                    }
                    dl = make_float4(d[0], d[1], d[2], d[3]);
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u) l[u] = (l[u] == INFINITY || l[u] == -INFINITY) ? -INFINITY : -l[u] * kLog2e;
                nl = make_float4(l[0], l[1], l[2], l[3]);
            };
            auto publish = [&](const float4& nl, const float4& dl) {
                const uint32_t ss = gs % C::kStatStages;
                mbar_wait(&stat_empty[ss], ((gs / C::kStatStages) & 1) ^ 1, 400);
                if constexpr (kStatsMma) {
                    // row q = 4 lane + u of the two compact tiles: [hi, mid, lo, 0 ...] of nl / c = -LSE / scale and of -delta
                    const float inv_c = 1.0f / p.scale_log2;
                    const float nv[4] = {nl.x, nl.y, nl.z, nl.w}, dv[4] = {dl.x, dl.y, dl.z, dl.w};
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t q = lane * 4 + u, off = (q >> 3) * 256 + (q & 7) * 16;
                        #pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            const float x = w ? dv[u] : ((nv[u] == -INFINITY) ? -INFINITY : nv[u] * inv_c);
                            const __nv_bfloat16 hi = __float2bfloat16(x);
                            const float r1 = (x == -INFINITY) ? 0.f : x - __bfloat162float(hi);
                            const __nv_bfloat16 mid = __float2bfloat16(r1);
                            const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
                            const uint32_t w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
                            sts128(stat_addr + ss * C::kStatSlotBytes + w * 4096 + off, w0, (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
                        }
                    }
                    fence_proxy_async_smem();
                } else {
                    sts128(stat_addr + ss * 1024 + lane * 16, __float_as_uint(nl.x), __float_as_uint(nl.y), __float_as_uint(nl.z), __float_as_uint(nl.w));
                    sts128(stat_addr + ss * 1024 + 512 + lane * 16, __float_as_uint(dl.x), __float_as_uint(dl.y), __float_as_uint(dl.z), __float_as_uint(dl.w));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_full[ss]);
                ++gs;
            };
            // two tiles of global loads in flight: with short MMAs (D=64) this warp sits right at the critical path
            float4 nlA, dlA, nlB, dlB;
            if (n_it > 0) fetch(nlA, dlA);
            if (n_it > 1) fetch(nlB, dlB);
            for (int it = 0; it < n_it; it += 2) {
                publish(nlA, dlA);
                if (it + 2 < n_it) fetch(nlA, dlA);
                if (it + 1 < n_it) {
                    publish(nlB, dlB);
                    if (it + 3 < n_it) fetch(nlB, dlB);
                }
            }
        }
    } else if (warp == 9) {
        // ----------------------------- TMA producer + scheduler -----------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        {   // whole warp, converged
            if (lane_id() == 0) { tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); tma_prefetch_desc(&mapdO); }
            __syncwarp();
            uint32_t git = 0, nacc = 0;
            int item = sched_first(p.sched_dkv, FA_BWD_PERSISTENT ? p.dyn_first : 0);
            for (uint32_t ix = 0;; ++ix) {
                const uint32_t slot = ix & 1;
                mbar_wait(&sched_empty[slot], ((ix >> 1) & 1) ^ 1, 441);
                sched_item[slot] = item;
                mbar_arrive_e(&sched_full[slot]);
                if (item >= n_items) break;
                int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
                mbar_wait(kv_free, (ix & 1) ^ 1, 442);        // K/V smem of the previous item released
                if (n_it > 0) {
                    mbar_arrive_expect_tx_e(k_full, C::kTileBytes);
                    #pragma unroll
                    for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sK + c * 16384, &mapK, k_full, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                    int l_qtile = i_start, hq = (bh % p.Hk) * p.G;     // (q tile, query head) walker of the item
                    const int bq = bh / p.Hk;
                    for (int it = 0; it < n_it; ++it, ++git) {
                        const uint32_t st = git % C::kStages;
                        uint8_t* sQi = sStage + st * C::kStageBytes;
                        uint8_t* sdOi = sQi + C::kTileBytes;
                        const int q0 = l_qtile * 128, hcur = hq;
                        if (++l_qtile == i_end) { l_qtile = i_start; ++hq; }
                        mbar_wait(&stage_empty[st], ((git / C::kStages) & 1) ^ 1, 410);
                        mbar_arrive_expect_tx_e(&q_full[st], C::kTileBytes);
                        #pragma unroll
                        for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sQi + c * 16384, &mapQ, &q_full[st], c * 64, q0, hcur, bq);
                        if (it == 0) {
                            mbar_arrive_expect_tx_e(v_full, C::kTileBytes);
                            #pragma unroll
                            for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sV + c * 16384, &mapV, v_full, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                        }
                        mbar_arrive_expect_tx_e(&do_full[st], C::kTileBytes);
                        #pragma unroll
                        for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sdOi + c * 16384, &mapdO, &do_full[st], c * 64, q0, hcur, bq);
                    }
                    ++nacc;
                }
                item = FA_BWD_PERSISTENT ? sched_next(p.sched_dkv, p.dyn_first) : n_items;
            }
            if (lane_id() == 0) sched_retire(p.sched_dkv);
        }
    } else if (warp == 8) {
        // ---------------------------------- MMA issuer ----------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aSt = smem_u32(sStage);
        uint32_t git = 0, gi = 0, nacc = 0;
        for (uint32_t ix = 0;; ++ix) {           // whole warp, converged
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            auto qfull = [&](uint32_t g) { mbar_wait(&q_full[g % C::kStages], (g / C::kStages) & 1, 421); };
            auto dofull = [&](uint32_t g) { mbar_wait(&do_full[g % C::kStages], (g / C::kStages) & 1, 423); };
            auto stage_addr = [&](uint32_t g) { return aSt + (g % C::kStages) * C::kStageBytes; };
            // statistics through the MMA: one more K = 16 step, ones (x) split(-LSE/scale) on S^T(G), ones (x) split(-delta) on dP^T(G);
            // the slot of iteration G is awaited at S^T(G) (always issued first) and released behind dP^T(G)
            const uint32_t aStat = smem_u32(sStat), aOnes = aStat + C::kStatStages * C::kStatSlotBytes;
            auto stat_scores = [&](uint32_t d_tmem, uint32_t G) {
                if constexpr (kStatsMma) {
                    mbar_wait(&stat_full[G % C::kStatStages], (G / C::kStatStages) & 1, 427); tc_fence_after();
                    umma_ss_e(d_tmem, make_smem_desc_noswizzle(aOnes, 128, 256),
                              make_smem_desc_noswizzle(aStat + (G % C::kStatStages) * C::kStatSlotBytes, 128, 256),
                              make_idesc(true, false, false, 128, 128), 1);
                }
            };
            auto stat_dp = [&](uint32_t G) {
                if constexpr (kStatsMma) {
                    umma_ss_e(tmem + kColDPT, make_smem_desc_noswizzle(aOnes, 128, 256),
                              make_smem_desc_noswizzle(aStat + (G % C::kStatStages) * C::kStatSlotBytes + 4096, 128, 256),
                              make_idesc(true, false, false, 128, 128), 1);
                    tc_commit_e(&stat_empty[G % C::kStatStages]);
                }
            };
            if (n_it > 0) {
                mbar_wait(k_full, nacc & 1, 420);
                qfull(git); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColST + (kDoubleS ? (gi & 1) * 128 : 0), aK, stage_addr(git));
                stat_scores(tmem + kColST + (kDoubleS ? (gi & 1) * 128 : 0), gi);
                tc_commit_e((kDoubleS && (gi & 1)) ? s_full1 : s_full);
                if (kDoubleS && n_it > 1) {
                    qfull(git + 1); tc_fence_after();
                    issue_scores<D, kBf16>(tmem + kColST + ((gi + 1) & 1) * 128, aK, stage_addr(git + 1));
                    stat_scores(tmem + kColST + ((gi + 1) & 1) * 128, gi + 1);
                    tc_commit_e(((gi + 1) & 1) ? s_full1 : s_full);
                }
                mbar_wait(v_full, nacc & 1, 422);
                dofull(git); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColDPT, aV, stage_addr(git) + C::kTileBytes); stat_dp(gi); tc_commit_e(dp_full);
            }
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it, gt = git + it;
                const uint32_t aQ = stage_addr(gt), adO = aQ + C::kTileBytes;
                const bool more = it + 1 < n_it;
                const uint32_t sbuf = kDoubleS ? (g & 1) * 128 : 0;
                mbar_wait(p_full, g & 1, 424);
                if (it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 429);     // previous item's dV/dK drained from TMEM
                tc_fence_after();
                issue_grad<D, kBf16>(tmem + kColDV, tmem + kColST + sbuf, adO, it > 0);   // dV += P^T dO_i
                if (kDoubleS) {
                    if (it + 2 < n_it) {                                                   // S^T(i+2) into the buffer dV(i) just read
                        qfull(gt + 2); tc_fence_after();
                        issue_scores<D, kBf16>(tmem + kColST + sbuf, aK, stage_addr(gt + 2));
                        stat_scores(tmem + kColST + sbuf, g + 2);
                        tc_commit_e((g & 1) ? s_full1 : s_full);
                    }
                } else if (more) {
                    qfull(gt + 1); tc_fence_after();
                    issue_scores<D, kBf16>(tmem + kColST, aK, stage_addr(gt + 1)); stat_scores(tmem + kColST, g + 1);
                    tc_commit_e(s_full);                                                                       // S^T(i+1)
                }
                mbar_wait(ds_full, g & 1, 426); tc_fence_after();
                issue_grad<D, kBf16>(tmem + kColDK, tmem + kColDPT, aQ, it > 0);           // dK += dS^T Q_i
                tc_commit_e(&stage_empty[gt % C::kStages]);
                if (more) {
                    dofull(gt + 1); tc_fence_after();
                    issue_scores<D, kBf16>(tmem + kColDPT, aV, stage_addr(gt + 1) + C::kTileBytes); stat_dp(g + 1);
                    tc_commit_e(dp_full);                                                                      // dP^T(i+1)
                }
            }
            if (n_it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 429);
            tc_commit_e(acc_full);                       // every MMA of the item is done -> accumulators final
            tc_commit_e(kv_free);                        // ... and K/V are no longer read
            gi += n_it; git += n_it; if (n_it > 0) ++nacc;
        }
    } else {
        // ------------------------------- compute warpgroups -------------------------------
        reg_alloc<BwdRegs<D>::kCompute>();
        const int h = warp >> 2;                         // column half
        const int r = tid & 127;                         // kv row in tile == TMEM lane
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tST = tmem + lane_field + kColST + h * 64;
        const uint32_t tDPT = tmem + lane_field + kColDPT + h * 64;
        const float c2 = p.scale_log2;
        if (FA_BWD_STAGGER && h == 1) named_bar_arrive(3, 256);       // warpgroup A takes the first turn
        uint32_t gi = 0;
        bool store_pending = false;
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            const int kv_g = jt * 128 + r;
            int q_lo = p.causal ? kv_g : 0, q_hi = p.Sq;  // queries that see my kv row
            if (p.col_lo) {
                const size_t ci = (size_t)(bh / p.Hk) * p.Sk + min(kv_g, p.Sk - 1);
                q_lo = max(q_lo, __ldg(p.col_lo + ci)); q_hi = min(q_hi, __ldg(p.col_hi + ci));
            }
            int qtile = i_start;                          // q tile of the current iteration (wraps per query head of the group)
            uint32_t bhq = (uint32_t)((bh / p.Hk) * p.H + (bh % p.Hk) * p.G);   // batch*H + query head of the current iteration
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it;
                const uint32_t ss = g % C::kStatStages;
                const uint32_t stat = smem_u32(sStat) + ss * 1024 + h * 256;
                const int q0 = qtile * 128 + h * 64;                 // global query index of my column 0
                const uint32_t bhq_it = bhq;
                if (++qtile == i_end) { qtile = i_start; ++bhq; }
                if constexpr (!kStatsMma) mbar_wait(&stat_full[ss], (g / C::kStatStages) & 1, 430);
                const uint32_t tSTi = tST + (kDoubleS ? (g & 1) * 128 : 0);
                if (kDoubleS) mbar_wait((g & 1) ? s_full1 : s_full, (g >> 1) & 1, 431);
                else mbar_wait(s_full, g & 1, 431);
                tc_fence_after();
                float pv[64];
                {
                    uint32_t s[2][32];
                    tmem_ld32(tSTi, s[0]); tmem_ld32(tSTi + 32, s[1]);
                    tc_wait_ld();
                    if (FA_BWD_STAGGER) named_bar_sync(3 + h, 256);
                    const uint64_t c2v = pack_f2(c2, c2);
                    #pragma unroll
                    for (int c = 0; c < 64; c += 4) {
                        uint64_t xa, xb;
                        if constexpr (kStatsMma) {               // the MMA already added -LSE / scale
                            xa = fmul2(pack_u2(s[c >> 5][c & 31], s[c >> 5][(c & 31) + 1]), c2v);
                            xb = fmul2(pack_u2(s[c >> 5][(c & 31) + 2], s[c >> 5][(c & 31) + 3]), c2v);
                        } else {
                            const float4 nl = lds128(stat + c * 4);
                            xa = ffma2(pack_u2(s[c >> 5][c & 31], s[c >> 5][(c & 31) + 1]), c2v, pack_f2(nl.x, nl.y));
                            xb = ffma2(pack_u2(s[c >> 5][(c & 31) + 2], s[c >> 5][(c & 31) + 3]), c2v, pack_f2(nl.z, nl.w));
                        }
                        if ((c & 15) < BwdPoly<D>::kDkv) {
                            ex2_poly2(xa, pv[c], pv[c + 1]); ex2_poly2(xb, pv[c + 2], pv[c + 3]);
                        } else {
                            float x0, x1, x2, x3;
                            unpack_f2(xa, x0, x1); unpack_f2(xb, x2, x3);
                            pv[c] = ex2_approx(x0); pv[c + 1] = ex2_approx(x1); pv[c + 2] = ex2_approx(x2); pv[c + 3] = ex2_approx(x3);
                        }
                    }
                    if (FA_BWD_STAGGER) named_bar_arrive(4 - h, 256);
                }
                if (q0 < q_lo || q0 + 64 > q_hi) {           // tile straddles the diagonal / a range end: keep q_lo <= q < q_hi
                    const int cmin = q_lo - q0, cmax = q_hi - 1 - q0;
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) if (c < cmin || c > cmax) pv[c] = 0.f;
                }
                uint64_t keep = ~0ull;                       // keep bit per column (query) of my kv row
                if constexpr (kDropout) {
                    keep = 0ull;
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) {
                        const uint32_t w = dropout_word(dropout_row_key(p.drop.seed0, bhq_it, (uint32_t)(q0 + c)), p.drop.seed1, (uint32_t)kv_g >> 2);
                        keep |= (uint64_t)dropout_keep(w, (uint32_t)kv_g, p.drop.thresh) << c;
                    }
                }
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t pk[16];
                    #pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = q * 32 + 2 * i;
                        if constexpr (kDropout)                  // dV sees the dropped-out, rescaled P^T; pv keeps P for dS
                            pk[i] = pack2<kBf16>(((keep >> c) & 1) ? pv[c] * p.drop.scale : 0.f, ((keep >> (c + 1)) & 1) ? pv[c + 1] * p.drop.scale : 0.f);
                        else
                            pk[i] = pack2<kBf16>(pv[c], pv[c + 1]);
                    }
                    tmem_st16(tSTi + q * 16, pk);
                }
                tc_wait_st(); tc_fence_before();
                mbar_arrive(p_full);
                mbar_wait(dp_full, g & 1, 432);
                tc_fence_after();
                {
                    uint32_t dp[2][32];
                    tmem_ld32(tDPT, dp[0]); tmem_ld32(tDPT + 32, dp[1]);
                    tc_wait_ld();
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint32_t pk[16];
                        #pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const int c = q * 32 + 2 * i;
                            float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);             // through the MMA: dP' already holds dP - delta
                            if constexpr (!kStatsMma) dl = lds128(stat + 512 + c * 4);
                            float d0, d1, d2, d3;        // dS = P o (dP - delta), packed: FADD2 + FMUL2
                            uint64_t dpa = pack_u2(dp[q][2 * i], dp[q][2 * i + 1]), dpb = pack_u2(dp[q][2 * i + 2], dp[q][2 * i + 3]);
                            if constexpr (kDropout) {
                                float e0, e1, e2, e3; unpack_f2(dpa, e0, e1); unpack_f2(dpb, e2, e3);
                                dpa = pack_f2(((keep >> c) & 1) ? e0 * p.drop.scale : 0.f, ((keep >> (c + 1)) & 1) ? e1 * p.drop.scale : 0.f);
                                dpb = pack_f2(((keep >> (c + 2)) & 1) ? e2 * p.drop.scale : 0.f, ((keep >> (c + 3)) & 1) ? e3 * p.drop.scale : 0.f);
                            }
                            if constexpr (!kStatsMma) { dpa = fadd2(dpa, pack_f2(dl.x, dl.y)); dpb = fadd2(dpb, pack_f2(dl.z, dl.w)); }
                            unpack_f2(fmul2(pack_f2(pv[c], pv[c + 1]), dpa), d0, d1);
                            unpack_f2(fmul2(pack_f2(pv[c + 2], pv[c + 3]), dpb), d2, d3);
                            pk[i] = pack2<kBf16>(d0, d1); pk[i + 1] = pack2<kBf16>(d2, d3);
                        }
                        tmem_st16(tDPT + q * 16, pk);
                    }
                }
                tc_wait_st(); tc_fence_before();
                mbar_arrive(ds_full);
                if constexpr (!kStatsMma) {
                    __syncwarp();
                    if (lane_id() == 0) mbar_arrive(&stat_empty[ss]);    // this warp is done with the slot's statistics
                }
            }
            gi += n_it;
            // ---- epilogue: dV, dK*scale -> 16-bit -> smem staging -> TMA store
            mbar_wait(acc_full, ix & 1, 433); tc_fence_after();
            // The previous item's store must have read the staging before anybody overwrites it.  tid 0 has waited for that read
            // (here when the staging is separate, right behind its store when it aliases K/V); the barrier puts EVERY thread behind
            // it.  Without it (round 1, aliased staging) an item with n_it == 0 — whose accumulators are "complete" at once —
            // let the other threads zero the staging while the previous item's dK/dV store was still reading it: the tile in
            // front of an empty item came out partly zero, run-to-run different (found by the C5 stress run, r02).
            if (C::kSepStage && tid == 0 && store_pending) tma_store_wait_read0();
            named_bar_sync(1, 256);
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDV, sOutV, r, h, 1.0f, n_it == 0);
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDK, sOutK, r, h, p.scale, n_it == 0);
            tc_fence_before();
            mbar_arrive(acc_empty);                          // TMEM accumulators drained
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (tid == 0) {
                #pragma unroll
                for (int c = 0; c < C::kChunks; ++c) {
                    tma_store_4d(&mapdV, sOutV + c * 16384, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                    tma_store_4d(&mapdK, sOutK + c * 16384, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                }
                tma_store_commit();
                if (!C::kSepStage) { tma_store_wait_read0(); mbar_arrive(kv_free); }   // staging aliases K/V
            }
            store_pending = true;
        }
        if (tid == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
    hang_trap_if_set();
}

// =================================================================================================
// dQ — persistent: one CTA per SM walks (b,h, q-tile) items (late = heavy tiles first) from a dynamic scheduler.
// =================================================================================================
template <int D> struct DqCfg {
    static constexpr int kChunks = D / 64;
    static constexpr int kTileBytes = 128 * D * 2;
    // Q, dO resident; K_j lives from S(j) to dQ(j) (long), V_j only for dP(j) (short) -> separate rings,
    // K one slot deeper so that the loads keep a full tile of lead over the MMAs.
    static constexpr int kKStages = (D == 128) ? 3 : 4;
    static constexpr int kVStages = (D == 128) ? 2 : 4;
    static constexpr bool kSepStage = (D == 64);          // own dQ staging: the next item's Q/dO can land under the epilogue
    static constexpr int kOffRes = 0;
    static constexpr int kOffK = 2 * kTileBytes;
    static constexpr int kOffV = kOffK + kKStages * kTileBytes;
    static constexpr int kOffOut = kOffV + kVStages * kTileBytes;
    static constexpr int kOffBar = kOffOut + (kSepStage ? kTileBytes : 0);
    static constexpr int kNumBars = 16 + 2 * kKStages + 2 * kVStages;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 32 + 1024;
};

template <int D, bool kBf16, bool kDropout = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                 const __grid_constant__ CUtensorMap mapdQ, const BwdParams p) {
    using C = DqCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + C::kOffRes;
    uint8_t* sdO = sQ + C::kTileBytes;
    uint8_t* sKr = smem + C::kOffK;                    // K ring
    uint8_t* sVr = smem + C::kOffV;                    // V ring
    uint8_t* sOut = C::kSepStage ? smem + C::kOffOut : sQ;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;            uint64_t* do_full = bars + 1;
    uint64_t* s_full = bars + 2;        uint64_t* dp_full = bars + 3;
    uint64_t* s_empty = bars + 4;       uint64_t* ds_full = bars + 5;
    uint64_t* acc_full = bars + 6;      uint64_t* dp_empty = bars + 7;
    uint64_t* acc_empty = bars + 8;     uint64_t* qdo_free = bars + 9;
    uint64_t* sched_full = bars + 10;   uint64_t* sched_empty = bars + 12;   // [2] each
    uint64_t* k_full = bars + 16;                       // [kKStages]
    uint64_t* k_empty = k_full + C::kKStages;           // [kKStages]
    uint64_t* v_full = k_empty + C::kKStages;           // [kVStages]
    uint64_t* v_empty = v_full + C::kVStages;           // [kVStages]
    volatile int* sched_item = reinterpret_cast<volatile int*>(bars + C::kNumBars);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_items = p.BH * p.n_qtiles;

    if (tid == 0) {
        mbar_init(q_full, 1); mbar_init(do_full, 1); mbar_init(s_full, 1); mbar_init(dp_full, 1);
        mbar_init(s_empty, 256); mbar_init(dp_empty, 256); mbar_init(ds_full, 256);
        mbar_init(acc_full, 1); mbar_init(acc_empty, 256);
        mbar_init(qdo_free, C::kSepStage ? 1 : 2);        // MMA thread (+ the store's read-done when staging aliases Q)
        for (int i = 0; i < 2; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 9); }   // MMA warp + 8 math warps
        for (int i = 0; i < C::kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < C::kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDS = 256 + D;   // dS buffers: kColDS + 64*(g&1)

    // item -> (bh, q tile, number of kv tiles); descending q tile = heavy first under causal (:219 truncation)
    auto decode = [&](int item, int& bh, int& iq, int& jb, int& n_it) {
        int qt; item_to_head_tile(item, p.BH, p.n_qtiles, p.hc_dq, bh, qt);
        iq = p.n_qtiles - 1 - qt;
        n_it = fwd_item_iters(p.row_lo, p.row_hi, bh / p.H, iq * 128, 0, p.Sq, p.Sk, p.causal, jb);   // kv tiles jb .. jb + n_it - 1
    };
    auto next_item = [&](uint32_t ix) -> int {           // whole warp
        const uint32_t slot = ix & 1;
        mbar_wait(&sched_full[slot], (ix >> 1) & 1, 540);
        const int item = __shfl_sync(0xffffffffu, sched_item[slot], 0);   // provably warp-uniform -> uniform registers downstream
        mbar_arrive_e(&sched_empty[slot]);
        return item;
    };

    if (warp >= 10) {
        reg_dealloc<BwdRegs<D>::kOther>();
    } else if (warp == 9) {
        // ----------------------------- TMA producer + scheduler (whole warp, converged) -----------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        if (lane_id() == 0) { tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); tma_prefetch_desc(&mapdO); }
        __syncwarp();
        uint32_t g = 0;
        int item = sched_first(p.sched_dq, FA_BWD_PERSISTENT ? p.dyn_first : 0);
        for (uint32_t ix = 0;; ++ix) {
            const uint32_t slot = ix & 1;
            mbar_wait(&sched_empty[slot], ((ix >> 1) & 1) ^ 1, 541);
            sched_item[slot] = item;
            mbar_arrive_e(&sched_full[slot]);
            if (item >= n_items) break;
            int bh, iq, jb, n_it; decode(item, bh, iq, jb, n_it);
            mbar_wait(qdo_free, (ix & 1) ^ 1, 542);          // Q/dO smem of the previous item released
            if (n_it > 0) {                                  // n_it == 0: no row of the tile sees a key (range masks) -> dQ = 0, no loads
                mbar_arrive_expect_tx_e(q_full, C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sQ + c * 16384, &mapQ, q_full, c * 64, iq * 128, bh % p.H, bh / p.H);
            }
            for (int it = 0; it < n_it; ++it, ++g) {
                const uint32_t ks = g % C::kKStages, vs = g % C::kVStages;
                uint8_t* sKj = sKr + ks * C::kTileBytes;
                uint8_t* sVj = sVr + vs * C::kTileBytes;
                mbar_wait(&k_empty[ks], ((g / C::kKStages) & 1) ^ 1, 510);
                mbar_arrive_expect_tx_e(&k_full[ks], C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sKj + c * 16384, &mapK, &k_full[ks], c * 64, (jb + it) * 128, (bh % p.H) / p.G, bh / p.H);
                if (it == 0) {
                    mbar_arrive_expect_tx_e(do_full, C::kTileBytes);
                    #pragma unroll
                    for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sdO + c * 16384, &mapdO, do_full, c * 64, iq * 128, bh % p.H, bh / p.H);
                }
                mbar_wait(&v_empty[vs], ((g / C::kVStages) & 1) ^ 1, 511);
                mbar_arrive_expect_tx_e(&v_full[vs], C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < C::kChunks; ++c) tma_load_4d_e(sVj + c * 16384, &mapV, &v_full[vs], c * 64, (jb + it) * 128, (bh % p.H) / p.G, bh / p.H);
            }
            item = FA_BWD_PERSISTENT ? sched_next(p.sched_dq, p.dyn_first) : n_items;
        }
        if (lane_id() == 0) sched_retire(p.sched_dq);
    } else if (warp == 8) {
        // ---------------------------------- MMA issuer (whole warp, converged) ----------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aKr = smem_u32(sKr), aVr = smem_u32(sVr);
        uint32_t gi = 0, nld = 0;                      // nld: items that loaded Q / dO (phase of q_full / do_full)
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, iq, jb, n_it; decode(item, bh, iq, jb, n_it);
            auto kfull = [&](uint32_t g) { mbar_wait(&k_full[g % C::kKStages], (g / C::kKStages) & 1, 521); };
            auto vfull = [&](uint32_t g) { mbar_wait(&v_full[g % C::kVStages], (g / C::kVStages) & 1, 523); };
            if (n_it > 0) {
                mbar_wait(q_full, nld & 1, 520);
                kfull(gi); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColS, aQ, aKr + (gi % C::kKStages) * C::kTileBytes); tc_commit_e(s_full);
                mbar_wait(do_full, nld & 1, 522);
                vfull(gi); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColDP, adO, aVr + (gi % C::kVStages) * C::kTileBytes); tc_commit_e(dp_full);
                tc_commit_e(&v_empty[gi % C::kVStages]);
                ++nld;
            } else {
                mbar_wait(acc_empty, (ix & 1) ^ 1, 529);   // keep the accumulator hand-shake in step: the math warps store zeros
            }
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it;
                if (it + 1 < n_it) {
                    mbar_wait(s_empty, g & 1, 524);                // S(j) is in registers
                    kfull(g + 1); tc_fence_after();
                    issue_scores<D, kBf16>(tmem + kColS, aQ, aKr + ((g + 1) % C::kKStages) * C::kTileBytes); tc_commit_e(s_full);     // S(j+1)
                    mbar_wait(dp_empty, g & 1, 528);               // dP(j) is in registers
                    vfull(g + 1); tc_fence_after();
                    issue_scores<D, kBf16>(tmem + kColDP, adO, aVr + ((g + 1) % C::kVStages) * C::kTileBytes); tc_commit_e(dp_full);  // dP(j+1)
                    tc_commit_e(&v_empty[(g + 1) % C::kVStages]);
                }
                mbar_wait(ds_full, g & 1, 526);
                if (it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 529);     // previous item's dQ drained from TMEM
                tc_fence_after();
                issue_grad<D, kBf16, false>(tmem + kColDQ, tmem + kColDS + (g & 1) * 64, aKr + (g % C::kKStages) * C::kTileBytes, it > 0);   // dQ += dS K_j
                tc_commit_e(&k_empty[g % C::kKStages]);
            }
            tc_commit_e(acc_full);                     // every MMA of the item is done -> dQ final
            tc_commit_e(qdo_free);                     // ... and Q/dO are no longer read
            gi += n_it;
        }
    } else {
        reg_alloc<BwdRegs<D>::kCompute>();
        const int h = warp >> 2;
        const int r = tid & 127;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + lane_field + kColS + h * 64;
        const uint32_t tDP = tmem + lane_field + kColDP + h * 64;
        const uint32_t tDS = tmem + lane_field + kColDS + h * 32;      // + 64 * (g & 1)
        const float c2 = p.scale_log2;
        if (FA_BWD_STAGGER && h == 1) named_bar_arrive(3, 256);       // warpgroup A takes the first turn
        uint32_t gi = 0;
        bool store_pending = false;
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, iq, jb, n_it; decode(item, bh, iq, jb, n_it);
            const int row_g = iq * 128 + r;
            float nl = -INFINITY, dl = 0.f;
            if (row_g < p.Sq) {
                const float l = __ldg(p.lse + (size_t)bh * p.Sq + row_g);
                nl = (l == -INFINITY) ? -INFINITY : -l * kLog2e;
                dl = __ldg(p.delta + (size_t)bh * p.Sq + row_g);
            }
            int k_lo = 0, k_hi = p.Sk;                    // keys this row may see (before the causal clip)
            if (p.row_lo) {
                const size_t ri = (size_t)(bh / p.H) * p.Sq + min(row_g, p.Sq - 1);
                k_lo = __ldg(p.row_lo + ri); k_hi = min(__ldg(p.row_hi + ri), p.Sk);
            }
            const uint32_t drop_key = kDropout ? dropout_row_key(p.drop.seed0, (uint32_t)bh, (uint32_t)row_g) : 0u;
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it;
                mbar_wait(s_full, g & 1, 530);
                tc_fence_after();
                float pv[64];
                {
                    uint32_t s[2][32];
                    tmem_ld32(tS, s[0]); tmem_ld32(tS + 32, s[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(s_empty);
                    if (FA_BWD_STAGGER) named_bar_sync(3 + h, 256);
                    const uint64_t c2v = pack_f2(c2, c2), nlv = pack_f2(nl, nl);
                    #pragma unroll
                    for (int c = 0; c < 64; c += 2) {
                        const uint64_t xa = ffma2(pack_u2(s[c >> 5][c & 31], s[c >> 5][(c & 31) + 1]), c2v, nlv);
                        if ((c & 15) < BwdPoly<D>::kDq) {
                            ex2_poly2(xa, pv[c], pv[c + 1]);
                        } else {
                            float x0, x1;
                            unpack_f2(xa, x0, x1);
                            pv[c] = ex2_approx(x0); pv[c + 1] = ex2_approx(x1);
                        }
                    }
                    if (FA_BWD_STAGGER) named_bar_arrive(4 - h, 256);
                }
                const int k0 = (jb + it) * 128 + h * 64;      // global key index of my column 0
                int cmax = k_hi - 1 - k0;
                if (p.causal) cmax = min(cmax, row_g - k0);
                const int cmin = k_lo - k0;
                if (cmax < 63 || cmin > 0) {
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) if (c > cmax || c < cmin) pv[c] = 0.f;
                }
                mbar_wait(dp_full, g & 1, 531);
                tc_fence_after();
                {
                    uint32_t dp[2][32];
                    tmem_ld32(tDP, dp[0]); tmem_ld32(tDP + 32, dp[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(dp_empty);
                    // dS(g) goes to buffer g&1; its previous reader dQ(g-2) completed before dp_full(g) fired
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint32_t pk[16];
                        const uint64_t ndl = pack_f2(-dl, -dl);
                        #pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int c = q * 32 + 2 * i;
                            float d0, d1;
                            uint64_t dpv = pack_u2(dp[q][2 * i], dp[q][2 * i + 1]);
                            if constexpr (kDropout) {               // dP reaches P only through the kept, rescaled elements
                                const uint32_t kcol = (uint32_t)(k0 + c);
                                const uint32_t w = dropout_word(drop_key, p.drop.seed1, kcol >> 2);
                                float e0, e1; unpack_f2(dpv, e0, e1);
                                e0 = dropout_keep(w, kcol, p.drop.thresh) ? e0 * p.drop.scale : 0.f;
                                e1 = dropout_keep(w, kcol + 1, p.drop.thresh) ? e1 * p.drop.scale : 0.f;
                                dpv = pack_f2(e0, e1);
                            }
                            unpack_f2(fmul2(pack_f2(pv[c], pv[c + 1]), fadd2(dpv, ndl)), d0, d1);
                            pk[i] = pack2<kBf16>(d0, d1);
                        }
                        tmem_st16(tDS + (g & 1) * 64 + q * 16, pk);
                    }
                }
                tc_wait_st(); tc_fence_before();
                mbar_arrive(ds_full);
            }
            gi += n_it;
            // ---- epilogue: dQ*scale -> 16-bit -> smem staging -> TMA store
            mbar_wait(acc_full, ix & 1, 532); tc_fence_after();
            // as in the dK/dV kernel: every thread behind tid 0's wait for the previous store's read of the staging (an item with
            // n_it == 0 reaches this point immediately)
            if (C::kSepStage && tid == 0 && store_pending) tma_store_wait_read0();
            named_bar_sync(1, 256);
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDQ, sOut, r, h, p.scale, n_it == 0);
            tc_fence_before();
            mbar_arrive(acc_empty);                          // dQ drained from TMEM
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (tid == 0) {
                #pragma unroll
                for (int c = 0; c < C::kChunks; ++c) tma_store_4d(&mapdQ, sOut + c * 16384, c * 64, iq * 128, bh % p.H, bh / p.H);
                tma_store_commit();
                if (!C::kSepStage) { tma_store_wait_read0(); mbar_arrive(qdo_free); }   // staging aliases Q
            }
            store_pending = true;
        }
        if (tid == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
    hang_trap_if_set();
}

template <int D, bool kBf16, bool kDropout>
int launch_bwd_td(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                  const CUtensorMap& mdq, const CUtensorMap& mdk, const CUtensorMap& mdv, const BwdParams& p,
                  int parts, cudaStream_t st) {
    {
        cudaError_t e = ensure_smem<fa_bwd_dkv_kernel<D, kBf16, kDropout>>(DkvCfg<D>::kSmemBytes, p.dev);
        if (e != cudaSuccess) return (int)e;
        e = ensure_smem<fa_bwd_dq_kernel<D, kBf16, kDropout>>(DqCfg<D>::kSmemBytes, p.dev);
        if (e != cudaSuccess) return (int)e;
    }
    // same order as the reference launcher (code/My_FlashAttention_optimized.py:111-126): dQ, then dK/dV
    if (parts & 2) {
        const int items = p.BH * p.n_qtiles;
        const int grid = FA_BWD_PERSISTENT ? (items < p.sms ? items : p.sms) : items;
        cudaError_t e = launch_pdl(fa_bwd_dq_kernel<D, kBf16, kDropout>, grid, kBwdThreads, DqCfg<D>::kSmemBytes, st, mq, mk, mv, mdo, mdq, p);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    if (parts & 4) {
        const int items = (p.BH / p.G) * p.n_ktiles;
        const int grid = FA_BWD_PERSISTENT ? (items < p.sms ? items : p.sms) : items;
        cudaError_t e = launch_pdl(fa_bwd_dkv_kernel<D, kBf16, kDropout>, grid, kBwdThreads, DkvCfg<D>::kSmemBytes, st, mq, mk, mv, mdo, mdk, mdv, p);
        if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaGetLastError();
}
template <int D, bool kBf16>
int launch_bwd_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                 const CUtensorMap& mdq, const CUtensorMap& mdk, const CUtensorMap& mdv, const BwdParams& p,
                 int parts, cudaStream_t st) {
    if (kStatsMma && p.drop.thresh) return (int)cudaErrorNotSupported;   // experimental build: dP - delta is formed inside the MMA, dropout must scale dP alone
    return p.drop.thresh ? launch_bwd_td<D, kBf16, true>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st)
                         : launch_bwd_td<D, kBf16, false>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st);
}

inline int launch_bwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                      const CUtensorMap& mdq, const CUtensorMap& mdk, const CUtensorMap& mdv, const BwdParams& p,
                      int D, int dtype, int parts, cudaStream_t st) {
    if (D == 64) return dtype ? launch_bwd_t<64, true>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st)
                              : launch_bwd_t<64, false>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st);
    return dtype ? launch_bwd_t<128, true>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st)
                 : launch_bwd_t<128, false>(mq, mk, mv, mdo, mdq, mdk, mdv, p, parts, st);
}

}  // namespace fa
