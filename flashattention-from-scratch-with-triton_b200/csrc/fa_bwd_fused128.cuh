// Fused single-pass backward for sm_100a, D = 128.  One kernel replaces
// flash_attention_dQ_kernel AND flash_attention_dKV_kernel (reference code/_flash_attention_kernel_optimized.py:164-258,
// :291-386): 5 GEMMs per (kv tile, q tile) pair instead of the reference split's 3 + 4 and one exponential per score
// element instead of two.  On B200 the D = 128 backward is POWER-bound (profiles/r02_ncu_full_summary_c4.json: the dQ / dK-dV
// kernels hold the tensor pipe 84 % / 79 % busy while the SM clock sinks to 1.47 / 1.53 GHz of 1.97), so executing 5/7 of
// the MMA work for the same gradients is the lever — not a higher pipe utilisation.
//
// Same structure as the D = 64 kernel (fa_bwd_fused.cuh): one CTA walks (batch, K/V head, kv tile) items, transposed
// scores (kv rows = TMEM lanes), per q tile i
//   S^T = K Q_i^T, dP^T = V dO_i^T                         (SS, N = 128)
//   P^T = exp2(S^T c - LSE_i log2e)   -> 16-bit into TMEM; dV += P^T dO_i      (A from TMEM)
//   dS^T = P^T o (dP^T - delta_i)     -> 16-bit into TMEM; dK += dS^T Q_i      (A from TMEM)
//                                     -> and into a swizzled smem tile [128 kv][128 q]; read MN-major it is dS[q, kv]:
//   dQ_i (partial) = dS K             (SS, both operands MN-major) -> TMEM [128 q lanes x 128]
//   the math warps move the fp32 partial to smem and the reducer warp adds it into the fp32 accumulator [B*H, Sq, 128]
//   with cp.reduce.async.bulk.tensor (.add); fa_dq_convert_kernel then writes dq = scale * accumulator in 16 bits.
//
// What D = 128 changes: S^T, dP^T, dV, dK fill all 512 TMEM columns, so the dQ partial lives in the HOLE the 16-bit
// operands leave behind.  The two math warpgroups place P^T in columns [0,64) (the low half of S^T) and dS^T in [192,256)
// (the HIGH half of dP^T), which frees ONE contiguous 128-column run [64,192) for dQ_i — a single N = 128 MMA, no operand
// re-read.  Because a warpgroup then writes 16-bit data over columns the OTHER warpgroup reads its scores from, two
// 256-thread mbarriers (s_loaded, dp_loaded) order "both halves are in registers" before either store.  The price:
// S^T(i+1) / dP^T(i+1) cannot be issued before dQ_i has been drained (dq_taken), so one tile is in flight at a time and
// the tensor pipe runs S dP | dV | dQ dK with the exp / dS math in the gaps.
//
// Shared memory (227 KB, no slack: the dynamic segment must start 1024-aligned, checked at run time):
//   K 32 | V 32 | Q x2 64 | dO 32 | dS^T 32 | dQ staging 32 | statistics 2 x 1 KB | barriers
// dO is single-buffered (free after dV_i, needed again by dP^T(i+1) a whole dS phase + two MMAs later).  The fp32 dQ
// partial is 64 KB = 4 boxes [128 q][32 d]: two go to the dedicated 32 KB, the other two into the Q_i stage buffer, which is
// dead once dK_i has completed and is not needed again before Q_(i+2) — the producer reloads it only after the reduce
// has read it (q_empty counts the MMA commit AND the reducer).  The TMA reduce of a 64 KB partial takes ~0.9 us per SM
// (measured: staging over the dS^T tile, which is rewritten half an iteration later, cost 27 % of the math warps' time in
// that wait); in the Q buffer it has a whole iteration.  The dV / dK epilogue staging is the dS^T tile + the dedicated 32 KB.
// Warps: 0-7 math (two warpgroups, 64 score columns each), 8 MMA issuer, 9 TMA producer + scheduler, 10 statistics, 11 dQ reducer.
#pragma once
#include "fa_bwd.cuh"
#include "fa_bwd_fused.cuh"

namespace fa {

// of every 16 score columns, this many (0, 4, 8) take their exp2 from the FMA-pipe polynomial instead of MUFU
#ifndef FA_F128_POLY
#define FA_F128_POLY 8
#endif
// the two math warpgroups take turns on the exp phase (named barriers 3/4)
#ifndef FA_F128_STAGGER
#define FA_F128_STAGGER 0
#endif
// MMA order after dS(i): 0 = dQ(i) then dK(i) (the drain of dQ overlaps dK), 1 = dK(i) then dQ(i)
#ifndef FA_F128_DK_FIRST
#define FA_F128_DK_FIRST 0
#endif

// timing experiments only (wrong results): 1 = no TMA reduce, 2 = no dQ MMA, 4 = no dS^T smem copy, 8 = no dQ staging stores,
// 16 = no statistics LDS (constants instead)
#ifndef FA_F128_SKIP
#define FA_F128_SKIP 0
#endif

struct Fused128Cfg {
    static constexpr int D = 128;
    static constexpr int kTileBytes = 128 * D * 2;                     // 32 KB
    static constexpr int kQStages = 2;
    static constexpr int kStatStages = 2;
    static constexpr int kOffK = 0;
    static constexpr int kOffV = kTileBytes;
    static constexpr int kOffQ = 2 * kTileBytes;
    static constexpr int kOffdO = kOffQ + kQStages * kTileBytes;
    static constexpr int kOffDS = kOffdO + kTileBytes;                 // dS^T tile; dV store staging
    static constexpr int kOffDQ = kOffDS + kTileBytes;                 // dQ staging boxes 2,3 (boxes 0,1: the Q_i stage); dK store staging
    static constexpr int kOffStat = kOffDQ + kTileBytes;
    static constexpr int kOffBar = kOffStat + kStatStages * 1024;
    static constexpr int kNumBars = 32;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 32;
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

template <bool kBf16>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_fused128_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                       const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                       const __grid_constant__ CUtensorMap mapdK, const __grid_constant__ CUtensorMap mapdV,
                       const __grid_constant__ CUtensorMap mapdQacc, const BwdParams p) {
    using C = Fused128Cfg;
    constexpr int D = 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();              // SWIZZLE_128B tiles need 1024-byte alignment and there is no room for slack
    uint8_t* sK = smem + C::kOffK;
    uint8_t* sV = smem + C::kOffV;
    uint8_t* sQ = smem + C::kOffQ;
    uint8_t* sdO = smem + C::kOffdO;
    uint8_t* sDS = smem + C::kOffDS;                   // also: dV store staging
    float* sStat = reinterpret_cast<float*>(smem + C::kOffStat);
    uint8_t* sOutV = sDS;
    uint8_t* sOutK = smem + C::kOffDQ;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* k_full = bars + 0;        uint64_t* v_full = bars + 1;        uint64_t* kv_free = bars + 2;
    uint64_t* s_full = bars + 3;        uint64_t* dp_full = bars + 4;
    uint64_t* s_loaded = bars + 5;      uint64_t* dp_loaded = bars + 6;     // both warpgroups hold S^T / dP^T in registers
    uint64_t* p_full = bars + 7;        uint64_t* ds_full = bars + 8;
    uint64_t* dq_full = bars + 9;       uint64_t* dq_taken = bars + 10;     // dQ partial in TMEM / in registers
    uint64_t* dqs_full = bars + 11;     uint64_t* dqs_empty = bars + 12;    // dQ partial staged in smem / read by the reduce
    uint64_t* acc_full = bars + 13;     uint64_t* acc_empty = bars + 14;
    uint64_t* do_full = bars + 15;      uint64_t* do_empty = bars + 16;
    uint64_t* q_full = bars + 17;       uint64_t* q_empty = bars + 19;      // [2] each
    uint64_t* sched_full = bars + 21;   uint64_t* sched_empty = bars + 23;  // [2] each
    uint64_t* stat_full = bars + 25;    uint64_t* stat_empty = bars + 27;   // [2] each
    uint64_t* dk_done = bars + 29;      // dK_i complete: the Q_i stage may take the dQ staging
    volatile int* sched_item = reinterpret_cast<volatile int*>(bars + C::kNumBars);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_items = (p.BH / p.G) * p.n_ktiles;       // one item per (batch, K/V head, kv tile)

    if (tid == 0) {
        mbar_init(k_full, 1); mbar_init(v_full, 1); mbar_init(kv_free, 1);
        mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(s_loaded, 256); mbar_init(dp_loaded, 256);
        mbar_init(p_full, 256); mbar_init(ds_full, 256); mbar_init(dq_full, 1); mbar_init(dq_taken, 256);
        mbar_init(dqs_full, 256); mbar_init(dqs_empty, 1); mbar_init(acc_full, 1); mbar_init(acc_empty, 256);
        mbar_init(do_full, 1); mbar_init(do_empty, 1); mbar_init(dk_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2);              // dK_i done (MMA commit) + dQ staging in it read (reducer)
            mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 11);     // MMA, 8 math warps, statistics, reducer
            mbar_init(&stat_full[i], 1); mbar_init(&stat_empty[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    // S^T [0,128) -> P^T (16-bit) [0,64);  dP^T [128,256) -> dS^T (16-bit) [192,256);  dQ partial [64,192);  dV [256,384);  dK [384,512)
    constexpr uint32_t kColST = 0, kColP = 0, kColDQ = 64, kColDPT = 128, kColDS = 192, kColDV = 256, kColDK = 384;

    auto decode = [&](int item, int& bh, int& jt, int& i_start, int& i_end, int& n_it) {
        item_to_head_tile(item, p.BH / p.G, p.n_ktiles, p.hc_dkv, bh, jt);
        i_start = p.causal ? jt : 0;
        i_end = p.n_qtiles;
        n_it = max(i_end - i_start, 0) * p.G;
    };
    auto next_item = [&](uint32_t ix) -> int {                 // whole warp
        const uint32_t slot = ix & 1;
        mbar_wait(&sched_full[slot], (ix >> 1) & 1, 740);
        const int item = __shfl_sync(0xffffffffu, sched_item[slot], 0);
        mbar_arrive_e(&sched_empty[slot]);
        return item;
    };

    if (warp == 11) {
        // ------------------------------ dQ reducer ------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        uint32_t nd = 0;
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            int qtile = i_start, hq = (bh % p.Hk) * p.G;
            const int bq = bh / p.Hk;
            for (int it = 0; it < n_it; ++it, ++nd) {
                const int q0 = qtile * 128, bhq = bq * p.H + hq;
                if (++qtile == i_end) { qtile = i_start; ++hq; }
                mbar_wait(dqs_full, nd & 1, 750);
                if (lane_id() == 0) {
                    if (!(FA_F128_SKIP & 1)) {
                        const uint8_t* sQi = sQ + (nd % C::kQStages) * C::kTileBytes;
                        tma_reduce_add_3d(&mapdQacc, sQi, 0, q0, bhq);
                        tma_reduce_add_3d(&mapdQacc, sQi + 16384, 32, q0, bhq);
                        tma_reduce_add_3d(&mapdQacc, sOutK, 64, q0, bhq);
                        tma_reduce_add_3d(&mapdQacc, sOutK + 16384, 96, q0, bhq);
                        tma_store_commit();
                        tma_store_wait_read0();
                    }
                    mbar_arrive(dqs_empty);
                    mbar_arrive(&q_empty[nd % C::kQStages]);
                }
                __syncwarp();
            }
        }
        if (lane_id() == 0) tma_store_wait_all0();
    } else if (warp == 10) {
        // ------------------------------ statistics loader (as in the dK/dV kernel) ------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        const int lane = lane_id();
        const uint32_t stat_addr = smem_u32(sStat);
        uint32_t gs = 0;
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            int s_qtile = i_start;
            size_t s_row0 = ((size_t)(bh / p.Hk) * p.H + (size_t)(bh % p.Hk) * p.G) * p.Sq;
            const bool vec_ok = (p.Sq & 3) == 0;
            auto fetch = [&](float4& nl, float4& dl) {
                const int q0 = s_qtile * 128 + lane * 4;
                const size_t off = s_row0 + q0;
                if (++s_qtile == i_end) { s_qtile = i_start; s_row0 += p.Sq; }
                float l[4];
                if (vec_ok && q0 + 4 <= p.Sq) {
                    const float4 lv = __ldg(reinterpret_cast<const float4*>(p.lse + off));
                    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.delta + off));
                    dl = make_float4(-dv.x, -dv.y, -dv.z, -dv.w);
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
                mbar_wait(&stat_empty[ss], ((gs / C::kStatStages) & 1) ^ 1, 700);
                sts128(stat_addr + ss * 1024 + lane * 16, __float_as_uint(nl.x), __float_as_uint(nl.y), __float_as_uint(nl.z), __float_as_uint(nl.w));
                sts128(stat_addr + ss * 1024 + 512 + lane * 16, __float_as_uint(dl.x), __float_as_uint(dl.y), __float_as_uint(dl.z), __float_as_uint(dl.w));
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_full[ss]);
                ++gs;
            };
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
        // ----------------------------- TMA producer + scheduler (whole warp, converged) -----------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        if (lane_id() == 0) { tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); tma_prefetch_desc(&mapdO); }
        __syncwarp();
        uint32_t git = 0;
        int item = sched_first(p.sched_dkv, p.dyn_first);
        for (uint32_t ix = 0;; ++ix) {
            const uint32_t slot = ix & 1;
            mbar_wait(&sched_empty[slot], ((ix >> 1) & 1) ^ 1, 741);
            sched_item[slot] = item;
            mbar_arrive_e(&sched_full[slot]);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            mbar_wait(kv_free, (ix & 1) ^ 1, 742);            // K/V of the previous item no longer read
            mbar_arrive_expect_tx_e(k_full, C::kTileBytes);
            #pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_4d_e(sK + c * 16384, &mapK, k_full, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
            int l_qtile = i_start, hq = (bh % p.Hk) * p.G;
            const int bq = bh / p.Hk;
            for (int it = 0; it < n_it; ++it, ++git) {
                const uint32_t st = git % C::kQStages;
                uint8_t* sQi = sQ + st * C::kTileBytes;
                const int q0 = l_qtile * 128, hcur = hq;
                if (++l_qtile == i_end) { l_qtile = i_start; ++hq; }
                mbar_wait(&q_empty[st], ((git / C::kQStages) & 1) ^ 1, 710);
                mbar_arrive_expect_tx_e(&q_full[st], C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < 2; ++c) tma_load_4d_e(sQi + c * 16384, &mapQ, &q_full[st], c * 64, q0, hcur, bq);
                if (it == 0) {
                    mbar_arrive_expect_tx_e(v_full, C::kTileBytes);
                    #pragma unroll
                    for (int c = 0; c < 2; ++c) tma_load_4d_e(sV + c * 16384, &mapV, v_full, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                }
                mbar_wait(do_empty, (git & 1) ^ 1, 711);
                mbar_arrive_expect_tx_e(do_full, C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < 2; ++c) tma_load_4d_e(sdO + c * 16384, &mapdO, do_full, c * 64, q0, hcur, bq);
            }
            if (n_it == 0) {                                   // nothing to stream: V still has to arrive for the protocol (kv parity)
                mbar_arrive_expect_tx_e(v_full, C::kTileBytes);
                #pragma unroll
                for (int c = 0; c < 2; ++c) tma_load_4d_e(sV + c * 16384, &mapV, v_full, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
            }
            item = sched_next(p.sched_dkv, p.dyn_first);
        }
        if (lane_id() == 0) sched_retire(p.sched_dkv);
    } else if (warp == 8) {
        // ---------------------------------- MMA issuer (whole warp, converged) ----------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ0 = smem_u32(sQ), adO = smem_u32(sdO), aDS = smem_u32(sDS);
        uint32_t g = 0;                                          // global iteration counter (all per-iteration barriers flip once per g)
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            mbar_wait(k_full, ix & 1, 720);
            mbar_wait(v_full, ix & 1, 721);
            for (int it = 0; it < n_it; ++it, ++g) {
                const uint32_t aQ = aQ0 + (g % C::kQStages) * C::kTileBytes;
                // S^T(g), dP^T(g): their TMEM runs hold dQ(g-1) until the math warps have it in registers
                if (g > 0) mbar_wait(dq_taken, (g - 1) & 1, 722);
                mbar_wait(&q_full[g % C::kQStages], (g / C::kQStages) & 1, 723); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColST, aK, aQ); tc_commit_e(s_full);
                mbar_wait(do_full, g & 1, 724); tc_fence_after();
                issue_scores<D, kBf16>(tmem + kColDPT, aV, adO); tc_commit_e(dp_full);
                mbar_wait(p_full, g & 1, 725);
                if (it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 726);    // previous item's dV / dK drained from TMEM
                tc_fence_after();
                issue_grad<D, kBf16, false>(tmem + kColDV, tmem + kColP, adO, it > 0);        // dV += P^T dO_i
                tc_commit_e(do_empty);                                                         // dO_i is no longer read
                mbar_wait(ds_full, g & 1, 727); tc_fence_after();
                if (FA_F128_DK_FIRST) {
                    issue_grad<D, kBf16, false>(tmem + kColDK, tmem + kColDS, aQ, it > 0);    // dK += dS^T Q_i
                    tc_commit_e(&q_empty[g % C::kQStages]); tc_commit_e(dk_done);
                    if (!(FA_F128_SKIP & 2)) issue_dq_partial<D, kBf16>(tmem + kColDQ, aDS, aK);   // dQ_i partial = dS K
                    tc_commit_e(dq_full);
                } else {
                    if (!(FA_F128_SKIP & 2)) issue_dq_partial<D, kBf16>(tmem + kColDQ, aDS, aK);
                    tc_commit_e(dq_full);
                    issue_grad<D, kBf16, false>(tmem + kColDK, tmem + kColDS, aQ, it > 0);
                    tc_commit_e(&q_empty[g % C::kQStages]); tc_commit_e(dk_done);
                }
            }
            if (n_it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 726);
            tc_commit_e(acc_full);                       // every MMA of the item is done -> accumulators final
            tc_commit_e(kv_free);                        // ... and K / V are no longer read
        }
    } else {
        // ------------------------------- math warpgroups -------------------------------
        reg_alloc<BwdRegs<D>::kCompute>();
        const int h = warp >> 2;                         // column half (64 of the 128 query columns)
        const int r = tid & 127;                         // kv row in tile == TMEM lane (dQ drain: q row)
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tST = tmem + lane_field + kColST + h * 64;
        const uint32_t tDPT = tmem + lane_field + kColDPT + h * 64;
        const uint32_t tP = tmem + lane_field + kColP + h * 32;
        const uint32_t tDS = tmem + lane_field + kColDS + h * 32;
        const uint32_t tDQ = tmem + lane_field + kColDQ + h * 64;       // my 64 of the 128 dQ columns
        const uint32_t sDSh = smem_u32(sDS) + h * 16384;                // my 64 q columns of the dS^T tile = one swizzled chunk
        const uint32_t sDQded = smem_u32(sOutK);                        // dedicated dQ staging (warpgroup B's two fp32 boxes)
        const float c2 = p.scale_log2;
        uint32_t g = 0;
        bool store_pending = false;
        if (FA_F128_STAGGER && h == 1) named_bar_arrive(3, 256);
        auto staging_free = [&]() {                      // the dV / dK store of the previous item has read the staging
            if (store_pending) {
                if (tid == 0) tma_store_wait_read0();
                named_bar_sync(1, 256);
                store_pending = false;
            }
        };
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            const int kv_g = jt * 128 + r;
            int qtile = i_start;
            for (int it = 0; it < n_it; ++it, ++g) {
                const uint32_t ss = g % C::kStatStages;
                const uint32_t stat = smem_u32(sStat) + ss * 1024 + h * 256;
                const int q0 = qtile * 128 + h * 64;                 // global query index of my column 0
                if (++qtile == i_end) qtile = i_start;
                mbar_wait(&stat_full[ss], (g / C::kStatStages) & 1, 730);
                mbar_wait(s_full, g & 1, 731);
                tc_fence_after();
                float pv[64];
                {
                    uint32_t s[2][32];
                    tmem_ld32(tST, s[0]); tmem_ld32(tST + 32, s[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(s_loaded);
                    if (FA_F128_STAGGER) named_bar_sync(3 + h, 256);
                    const uint64_t c2v = pack_f2(c2, c2);
                    #pragma unroll
                    for (int c = 0; c < 64; c += 4) {
                        const float4 nl = (FA_F128_SKIP & 16) ? make_float4(-8.f, -8.f, -8.f, -8.f) : lds128(stat + c * 4);
                        const uint64_t xa = ffma2(pack_u2(s[c >> 5][c & 31], s[c >> 5][(c & 31) + 1]), c2v, pack_f2(nl.x, nl.y));
                        const uint64_t xb = ffma2(pack_u2(s[c >> 5][(c & 31) + 2], s[c >> 5][(c & 31) + 3]), c2v, pack_f2(nl.z, nl.w));
                        if ((c & 15) < FA_F128_POLY) {
                            ex2_poly2(xa, pv[c], pv[c + 1]); ex2_poly2(xb, pv[c + 2], pv[c + 3]);
                        } else {
                            float x0, x1, x2, x3;
                            unpack_f2(xa, x0, x1); unpack_f2(xb, x2, x3);
                            pv[c] = ex2_approx(x0); pv[c + 1] = ex2_approx(x1); pv[c + 2] = ex2_approx(x2); pv[c + 3] = ex2_approx(x3);
                        }
                    }
                    if (FA_F128_STAGGER) named_bar_arrive(4 - h, 256);
                }
                if (p.causal && q0 < kv_g) {                 // tile straddles the diagonal: query q sees my kv row iff q >= kv_g
                    const int cmin = kv_g - q0;
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) if (c < cmin) pv[c] = 0.f;
                }
                mbar_wait(s_loaded, g & 1, 732);             // the other warpgroup has its half of S^T in registers too
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t pk[16];
                    #pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack2<kBf16>(pv[q * 32 + 2 * i], pv[q * 32 + 2 * i + 1]);
                    tmem_st16(tP + q * 16, pk);
                }
                tc_wait_st(); tc_fence_before();
                mbar_arrive(p_full);
                mbar_wait(dp_full, g & 1, 733);
                tc_fence_after();
                {
                    uint32_t dp[2][32];
                    tmem_ld32(tDPT, dp[0]); tmem_ld32(tDPT + 32, dp[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(dp_loaded);
                    if (it == 0) staging_free();
                    uint32_t pk[2][16];
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        #pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const int c = q * 32 + 2 * i;
                            const float4 dl = (FA_F128_SKIP & 16) ? make_float4(0.f, 0.f, 0.f, 0.f) : lds128(stat + 512 + c * 4);
                            float d0, d1, d2, d3;        // dS = P o (dP - delta), packed: FADD2 + FMUL2
                            const uint64_t dpa = pack_u2(dp[q][2 * i], dp[q][2 * i + 1]), dpb = pack_u2(dp[q][2 * i + 2], dp[q][2 * i + 3]);
                            unpack_f2(fmul2(pack_f2(pv[c], pv[c + 1]), fadd2(dpa, pack_f2(dl.x, dl.y))), d0, d1);
                            unpack_f2(fmul2(pack_f2(pv[c + 2], pv[c + 3]), fadd2(dpb, pack_f2(dl.z, dl.w))), d2, d3);
                            pk[q][i] = pack2<kBf16>(d0, d1); pk[q][i + 1] = pack2<kBf16>(d2, d3);
                        }
                        if (!(FA_F128_SKIP & 4))
                        #pragma unroll
                        for (int j = 0; j < 4; ++j)                          // the smem tile dQ reads as dS
                            sts128(sDSh + sw128_offset(r, q * 4 + j), pk[q][4 * j], pk[q][4 * j + 1], pk[q][4 * j + 2], pk[q][4 * j + 3]);
                    }
                    mbar_wait(dp_loaded, g & 1, 735);        // the other warpgroup has its half of dP^T in registers too
                    tmem_st16(tDS, pk[0]); tmem_st16(tDS + 16, pk[1]);        // A operand of dK, in the high half of dP^T
                }
                tc_wait_st(); tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(ds_full);
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&stat_empty[ss]);
                // ---- drain dQ(g): TMEM -> registers (frees S^T / dP^T for tile g+1) -> fp32 staging -> reducer warp
                mbar_wait(dq_full, g & 1, 736); tc_fence_after();
                {
                    uint32_t v[2][32];
                    tmem_ld32(tDQ, v[0]); tmem_ld32(tDQ + 32, v[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(dq_taken);
                    // warpgroup A stages into the Q_i buffer (dead after dK_i), warpgroup B into the dedicated 32 KB (free once the
                    // previous reduce has read it — a whole iteration ago)
                    uint32_t sDQh;
                    if (h == 0) { mbar_wait(dk_done, g & 1, 739); sDQh = smem_u32(sQ) + (g % C::kQStages) * C::kTileBytes; }
                    else { if (g > 0) mbar_wait(dqs_empty, (g - 1) & 1, 734); sDQh = sDQded; }
                    if (!(FA_F128_SKIP & 8))
                    #pragma unroll
                    for (int b = 0; b < 2; ++b)
                        #pragma unroll
                        for (int j = 0; j < 8; ++j)
                            sts128(sDQh + b * 16384 + sw128_offset(r, j), v[b][4 * j], v[b][4 * j + 1], v[b][4 * j + 2], v[b][4 * j + 3]);
                    fence_proxy_async_smem();
                    mbar_arrive(dqs_full);
                }
            }
            // ---- epilogue: dV, dK*scale -> 16-bit -> smem staging (over the dS^T tile and the dQ staging) -> TMA store
            mbar_wait(acc_full, ix & 1, 737); tc_fence_after();
            staging_free();                                  // (n_it == 0 items back to back)
            if (g > 0) mbar_wait(dqs_empty, (g - 1) & 1, 738);              // the last reduce has read the staging
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDV, sOutV, r, h, 1.0f, n_it == 0);
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDK, sOutK, r, h, p.scale, n_it == 0);
            tc_fence_before();
            mbar_arrive(acc_empty);                          // TMEM accumulators drained
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (tid == 0) {
                #pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tma_store_4d(&mapdV, sOutV + c * 16384, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                    tma_store_4d(&mapdK, sOutK + c * 16384, c * 64, jt * 128, bh % p.Hk, bh / p.Hk);
                }
                tma_store_commit();
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

template <bool kBf16>
int launch_bwd_fused128_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                          const CUtensorMap& mdk, const CUtensorMap& mdv, const CUtensorMap& macc, const BwdParams& p,
                          const float* acc, void* dq, RowStrides s_dq, cudaStream_t st, int parts) {
    constexpr int D = 128;
    {
        cudaError_t e = ensure_smem<fa_bwd_fused128_kernel<kBf16>>(Fused128Cfg::kSmemBytes, p.dev);
        if (e != cudaSuccess) return (int)e;
    }
    const int items = (p.BH / p.G) * p.n_ktiles;
    const int grid = items < p.sms ? items : p.sms;
    cudaError_t e = cudaSuccess;
    if (parts & 8) e = launch_pdl(fa_bwd_fused128_kernel<kBf16>, grid, kBwdThreads, Fused128Cfg::kSmemBytes, st, mq, mk, mv, mdo, mdk, mdv, macc, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess || !(parts & 16)) return (int)e;
    const long long rows = (long long)p.BH * p.Sq;
    const int rpb = 256 / (D / 8) * 4;
    long long blocks = (rows + rpb - 1) / rpb;
    const long long cap = (long long)p.sms * 8;
    if (blocks > cap) blocks = cap;
    e = launch_pdl(fa_dq_convert_kernel<D, kBf16>, (int)blocks, 256, 0, st, (const float4*)acc, (uint4*)dq, rows, p.H, p.Sq, s_dq, p.scale);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

}  // namespace fa
