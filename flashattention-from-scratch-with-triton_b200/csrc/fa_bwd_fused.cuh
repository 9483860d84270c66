// Fused single-pass backward for sm_100a, D = 64.  One kernel replaces flash_attention_dQ_kernel AND
// flash_attention_dKV_kernel (reference code/_flash_attention_kernel_optimized.py:164-258, :291-386): 5 GEMMs per
// (kv tile, q tile) pair instead of the reference split's 3 + 4, and P = exp2(S c - LSE log2e) is evaluated once instead
// of twice — at D = 64 the exp unit, not the tensor pipe, bounds the backward.
//
// Structure = the dK/dV kernel of fa_bwd.cuh (one CTA walks (batch, K/V head, kv tile) items, transposed scores so kv
// rows are TMEM lanes) plus, per q tile i:
//   dS^T (16-bit) is written twice: into TMEM in place of dP^T (A operand of dK += dS^T Q_i, as before) and into a
//   SWIZZLE_128B shared-memory tile [128 kv rows][128 q]; read as an MN-major A operand that tile is dS[q, kv], so
//   dQ_i(partial) = dS K   (B = the resident K tile, MN-major)  -> TMEM [128 q lanes x 64]
//   the math warps move the fp32 partial to a swizzled staging tile and the reducer warp adds it into the fp32 dQ
//   accumulator in global memory with cp.reduce.async.bulk.tensor (.add) — the sum over kv tiles happens in L2.
// A small conversion kernel (dq = scale * accumulator -> 16-bit, any strides) follows; the accumulator is zeroed by
// the delta preprocess kernel.  Summation order over kv tiles depends on scheduling, so dQ is not bitwise
// reproducible run to run (dK, dV are); the two-kernel path stays available as the deterministic mode.
//
// TMEM (512 columns): S^T [0,128)  dP^T / dS^T [128,256)  dV [256,320)  dK [320,384)  dQ partial | P^T [384,448) (they take
// turns)  K [448,480)  V [480,512) (16-bit copies of the resident tiles: A operands of the score MMAs).  P^T does not
// overwrite S^T, so S^T is released as soon as it is in registers and S^T(i+1) runs under the exp of tile i; dQ(i) is
// drained by the math warps in iteration i+1 right before they store P^T(i+1).
// Warps: 0-7 math (two warpgroups, 64 score columns each), 8 MMA issuer, 9 TMA producer + scheduler,
// 10 statistics loader, 11 dQ reducer.
#pragma once
#include "fa_bwd.cuh"
#include "fa_aux.cuh"

namespace fa {

// timing experiments only (wrong results): 1 = no dQ staging / reduce, 2 = no dQ MMA, 4 = no dS^T smem copy
#ifndef FA_FUSED_SKIP
#define FA_FUSED_SKIP 0
#endif
// of every 16 score columns, this many (0, 4, 8) get their exp2 from the FMA-pipe polynomial instead of MUFU: the exp
// phase of the loop is bound by the MUFU queue (ncu: stall_mio) while the FMA pipe idles; 4 measured best (-2..-7 %)
#ifndef FA_FUSED_POLY
#define FA_FUSED_POLY 4
#endif
// the two math warpgroups take turns on the exp phase (named barriers 3/4): while one is on the MUFU unit the other runs
// its FMA / LDS / TMEM-bound dS phase
#ifndef FA_FUSED_STAGGER
#define FA_FUSED_STAGGER 1      // measured: -3..-4 %
#endif
// MMA issue order after dS(i): dK(i), dP^T(i+1), dQ(i) instead of dK(i), dQ(i), dP^T(i+1)
#ifndef FA_FUSED_DP_FIRST
#define FA_FUSED_DP_FIRST 0
#endif

template <int D> struct FusedCfg {
    static_assert(D == 64, "the fused backward needs 3*D + 320 <= 512 TMEM columns: D = 64 only");
    static constexpr int kChunks = D / 64;
    static constexpr int kTileBytes = 128 * D * 2;
    static constexpr int kStages = 2;
    static constexpr int kKVBufs = 2;                                  // K/V double-buffered: the next item's K/V land under the current item
    static constexpr int kStatStages = 8;
    static constexpr int kOffRes = 0;                                  // K, V
    static constexpr int kOffStage = kKVBufs * 2 * kTileBytes;         // per stage: Q_i, dO_i
    static constexpr int kStageBytes = 2 * kTileBytes;
    static constexpr int kOffStat = kOffStage + kStages * kStageBytes;
    static constexpr int kOffDS = kOffStat + kStatStages * 1024;       // dS^T [128 kv][128 q] 16-bit = 2 chunks x 16 KB; the
                                                                       // dV / dK store staging aliases it at item end
    static constexpr int kOffDQ = kOffDS + 32768;                      // dQ partial fp32: 2 boxes [128 q][32 d] x 16 KB
    static constexpr int kOffBar = kOffDQ + 128 * D * 4;
    static constexpr int kNumBars = 24 + 3 * kStages + 2 * kStatStages;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 32 + 1024;
};

// dQ partial: D[tmem 128 q x D] = A[smem dS^T tile read MN-major: 128 q x 128 kv] * B[smem K tile, MN-major: 128 kv x D]
template <int D, bool kBf16>
__device__ __forceinline__ void issue_dq_partial(uint32_t d_tmem, uint32_t ds_addr, uint32_t k_addr) {
    constexpr uint32_t idesc = make_idesc(kBf16, true, true, 128, D);
    #pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_ss_e(d_tmem, make_smem_desc(ds_addr + k * 2048, 16384, 1024), make_smem_desc(k_addr + k * 2048, 16384, 1024),
                  idesc, k > 0);
}

// score-tile MMA with the resident operand in TMEM: D[tmem 128x128] = A[tmem: 128 lanes x D 16-bit = D/2 columns] * B[smem 128 x D, K-major]^T
template <int D, bool kBf16>
__device__ __forceinline__ void issue_scores_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr) {
    constexpr uint32_t idesc = make_idesc(kBf16, false, false, 128, 128);
    #pragma unroll
    for (int k = 0; k < D / 16; ++k)
        umma_ts_e(d_tmem, a_tmem + k * 8, make_smem_desc(b_addr + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), idesc, k > 0);
}

// kExt: 0 = the plain kernel; 1 = range masks (BwdParams::col_lo/col_hi, if set); 2 = dropout (and range masks, if set) — exactly
// as in the dK/dV kernel of fa_bwd.cuh.  Separate instantiations: the plain kernel is unchanged and the range-masked one does not
// carry the dropout generator.
template <int D, bool kBf16, int kExt = 0>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_fused_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                    const __grid_constant__ CUtensorMap mapdK, const __grid_constant__ CUtensorMap mapdV,
                    const __grid_constant__ CUtensorMap mapdQacc, const BwdParams p) {
    using C = FusedCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem + C::kOffRes;           // buffer b: K at b * 2 * kTileBytes, V right behind it
    uint8_t* sStage = smem + C::kOffStage;
    float* sStat = reinterpret_cast<float*>(smem + C::kOffStat);
    uint8_t* sDS = smem + C::kOffDS;
    uint8_t* sDQ = smem + C::kOffDQ;
    uint8_t* sOutV = sDS;                      // 16 KB each at D = 64
    uint8_t* sOutK = sDS + 16384;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* k_full = bars + 18;       uint64_t* v_full = bars + 20;        // [2] each (per K/V buffer)
    uint64_t* s_full = bars + 2;        uint64_t* dp_full = bars + 3;
    uint64_t* p_full = bars + 4;        uint64_t* ds_full = bars + 5;
    uint64_t* acc_full = bars + 6;      uint64_t* s_taken = bars + 7;     // S^T(i) is in registers
    uint64_t* acc_empty = bars + 8;     uint64_t* kv_free = bars + 22;       // [2]
    uint64_t* sched_full = bars + 10;   uint64_t* sched_empty = bars + 12;   // [2] each
    uint64_t* kvt_full = bars + 14;     // K (warpgroup A) and V (warpgroup B) of the item are in TMEM (256 math threads)
    uint64_t* dq_full = bars + 15;      // dQ partial of tile i is in TMEM (and dS^T(i) in smem is no longer read)
    uint64_t* dqs_full = bars + 16;     // dQ partial staged in smem (256 math threads)
    uint64_t* dqs_empty = bars + 17;    // the reduce has read the staging
    uint64_t* q_full = bars + 24;                       // [kStages]
    uint64_t* do_full = q_full + C::kStages;            // [kStages]
    uint64_t* stage_empty = do_full + C::kStages;       // [kStages]
    uint64_t* stat_full = stage_empty + C::kStages;     // [kStatStages]
    uint64_t* stat_empty = stat_full + C::kStatStages;  // [kStatStages]
    volatile int* sched_item = reinterpret_cast<volatile int*>(bars + C::kNumBars);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_items = (p.BH / p.G) * p.n_ktiles;       // one item per (batch, K/V head, kv tile)

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_free[i], 1); }
        mbar_init(s_full, 1); mbar_init(dp_full, 1);
        mbar_init(p_full, 256); mbar_init(ds_full, 256); mbar_init(acc_full, 1); mbar_init(acc_empty, 256);
        mbar_init(s_taken, 256); mbar_init(kvt_full, 256); mbar_init(dq_full, 1);
        mbar_init(dqs_full, 256); mbar_init(dqs_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 11); }
        for (int i = 0; i < C::kStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&do_full[i], 1); mbar_init(&stage_empty[i], 1); }
        for (int i = 0; i < C::kStatStages; ++i) { mbar_init(&stat_full[i], 1); mbar_init(&stat_empty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    // dQ partial (fp32, q lanes) and P^T (16-bit, kv lanes) take turns in ONE region: dQ(i-1) is drained right before P^T(i)
    // is written, dV(i) has consumed P^T(i) before the in-order MMA pipe reaches dQ(i).  The 64 columns this saves hold K and V
    // (16-bit, kv lanes) as the A operands of the two score MMAs: 32 KB less shared-memory operand traffic per tile — the
    // kernel is bound by the shared-memory pipe (ncu: LSU + tensor-core wavefronts ~ 85 % of peak).
    constexpr uint32_t kColST = 0, kColDPT = 128, kColDV = 256, kColDK = 256 + D, kColDQ = 256 + 2 * D, kColP = kColDQ;
    constexpr uint32_t kColKT = 256 + 3 * D, kColVT = kColKT + D / 2;
    static_assert(kColVT + D / 2 <= 512, "TMEM budget");

    // item -> (batch*Hk + kv head, kv tile, first q tile, iterations); with GQA the item walks the q tiles of every
    // query head of the group (dK/dV reduce over the group in TMEM)
    auto decode = [&](int item, int& bh, int& jt, int& i_start, int& i_end, int& n_it) {
        item_to_head_tile(item, p.BH / p.G, p.n_ktiles, p.hc_dkv, bh, jt);
        i_start = p.causal ? jt : 0;
        i_end = p.n_qtiles;
        if constexpr (kExt) {
            if (p.col_lo) {                                // q tiles [first query of the first kv row, last query of the last kv row)
                const size_t cb = (size_t)(bh / p.Hk) * p.Sk;
                i_start = max(i_start, __ldg(p.col_lo + cb + min(jt * 128, p.Sk - 1)) >> 7);
                i_end = min(i_end, ((min(__ldg(p.col_hi + cb + min(jt * 128 + 127, p.Sk - 1)), p.Sq) - 1) >> 7) + 1);
            }
        }
        n_it = max(i_end - i_start, 0) * p.G;
    };
    auto next_item = [&](uint32_t ix) -> int {                 // whole warp
        const uint32_t slot = ix & 1;
        mbar_wait(&sched_full[slot], (ix >> 1) & 1, 640);
        const int item = __shfl_sync(0xffffffffu, sched_item[slot], 0);
        mbar_arrive_e(&sched_empty[slot]);
        return item;
    };

    if (warp == 11) {
        // ------------------------------ dQ reducer ------------------------------
        // staged fp32 partial -> += into the accumulator [B*H, Sq, D] (rows past Sq are dropped by the tensor map)
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
                mbar_wait(dqs_full, nd & 1, 650);
                if (lane_id() == 0) {
                    if (!(FA_FUSED_SKIP & 1)) {
                        tma_reduce_add_3d(&mapdQacc, sDQ, 0, q0, bhq);
                        tma_reduce_add_3d(&mapdQacc, sDQ + 16384, 32, q0, bhq);
                        tma_store_commit();
                        tma_store_wait_read0();
                    }
                    mbar_arrive(dqs_empty);
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
                mbar_wait(&stat_empty[ss], ((gs / C::kStatStages) & 1) ^ 1, 600);
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
            mbar_wait(&sched_empty[slot], ((ix >> 1) & 1) ^ 1, 641);
            sched_item[slot] = item;
            mbar_arrive_e(&sched_full[slot]);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            const uint32_t kb = ix & 1;                   // K/V buffer of this item; its previous user was item ix - 2
            uint8_t* sK = sKV + kb * 2 * C::kTileBytes;
            uint8_t* sV = sK + C::kTileBytes;
            mbar_wait(&kv_free[kb], ((ix >> 1) & 1) ^ 1, 642);
            mbar_arrive_expect_tx_e(&k_full[kb], C::kTileBytes);
            tma_load_4d_e(sK, &mapK, &k_full[kb], 0, jt * 128, bh % p.Hk, bh / p.Hk);
            if (n_it == 0) {
                mbar_arrive_expect_tx_e(&v_full[kb], C::kTileBytes);
                tma_load_4d_e(sV, &mapV, &v_full[kb], 0, jt * 128, bh % p.Hk, bh / p.Hk);
            }
            if (n_it > 0) {
                int l_qtile = i_start, hq = (bh % p.Hk) * p.G;
                const int bq = bh / p.Hk;
                for (int it = 0; it < n_it; ++it, ++git) {
                    const uint32_t st = git % C::kStages;
                    uint8_t* sQi = sStage + st * C::kStageBytes;
                    uint8_t* sdOi = sQi + C::kTileBytes;
                    const int q0 = l_qtile * 128, hcur = hq;
                    if (++l_qtile == i_end) { l_qtile = i_start; ++hq; }
                    mbar_wait(&stage_empty[st], ((git / C::kStages) & 1) ^ 1, 610);
                    mbar_arrive_expect_tx_e(&q_full[st], C::kTileBytes);
                    tma_load_4d_e(sQi, &mapQ, &q_full[st], 0, q0, hcur, bq);
                    if (it == 0) {
                        mbar_arrive_expect_tx_e(&v_full[kb], C::kTileBytes);
                        tma_load_4d_e(sV, &mapV, &v_full[kb], 0, jt * 128, bh % p.Hk, bh / p.Hk);
                    }
                    mbar_arrive_expect_tx_e(&do_full[st], C::kTileBytes);
                    tma_load_4d_e(sdOi, &mapdO, &do_full[st], 0, q0, hcur, bq);
                }
            }
            item = sched_next(p.sched_dkv, p.dyn_first);
        }
        if (lane_id() == 0) sched_retire(p.sched_dkv);
    } else if (warp == 8) {
        // ---------------------------------- MMA issuer (whole warp, converged) ----------------------------------
        reg_dealloc<BwdRegs<D>::kOther>();
        const uint32_t aKV = smem_u32(sKV), aSt = smem_u32(sStage), aDS = smem_u32(sDS);
        uint32_t git = 0, gi = 0;
        for (uint32_t ix = 0;; ++ix) {
            const int item = next_item(ix);
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            auto qfull = [&](uint32_t g) { mbar_wait(&q_full[g % C::kStages], (g / C::kStages) & 1, 621); };
            auto dofull = [&](uint32_t g) { mbar_wait(&do_full[g % C::kStages], (g / C::kStages) & 1, 623); };
            auto stage_addr = [&](uint32_t g) { return aSt + (g % C::kStages) * C::kStageBytes; };
            const uint32_t kb = ix & 1;
            const uint32_t aK = aKV + kb * 2 * C::kTileBytes;         // smem K: B operand of the dQ MMA
            mbar_wait(kvt_full, ix & 1, 620);                         // K, V of this item copied into TMEM (implies k_full, v_full)
            tc_fence_after();
            if (n_it > 0) {
                if (gi > 0) mbar_wait(s_taken, (gi - 1) & 1, 627);      // the previous item's last S^T is in registers
                qfull(git); tc_fence_after();
                issue_scores_ts<D, kBf16>(tmem + kColST, tmem + kColKT, stage_addr(git)); tc_commit_e(s_full);
                dofull(git); tc_fence_after();
                issue_scores_ts<D, kBf16>(tmem + kColDPT, tmem + kColVT, stage_addr(git) + C::kTileBytes); tc_commit_e(dp_full);
            }
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it, gt = git + it;
                const uint32_t aQ = stage_addr(gt), adO = aQ + C::kTileBytes;
                const bool more = it + 1 < n_it;
                if (more) {                                              // S^T(i+1) under the exp of tile i
                    mbar_wait(s_taken, g & 1, 627);
                    qfull(gt + 1); tc_fence_after();
                    issue_scores_ts<D, kBf16>(tmem + kColST, tmem + kColKT, stage_addr(gt + 1)); tc_commit_e(s_full);
                }
                mbar_wait(p_full, g & 1, 624);
                if (it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 629);    // previous item's dV/dK/dQ drained from TMEM
                tc_fence_after();
                issue_grad<D, kBf16, false>(tmem + kColDV, tmem + kColP, adO, it > 0);      // dV += P^T dO_i
                mbar_wait(ds_full, g & 1, 626); tc_fence_after();
                issue_grad<D, kBf16>(tmem + kColDK, tmem + kColDPT, aQ, it > 0);            // dK += dS^T Q_i
                tc_commit_e(&stage_empty[gt % C::kStages]);
                if (FA_FUSED_DP_FIRST && more) {
                    dofull(gt + 1); tc_fence_after();
                    issue_scores_ts<D, kBf16>(tmem + kColDPT, tmem + kColVT, stage_addr(gt + 1) + C::kTileBytes); tc_commit_e(dp_full);   // dP^T(i+1)
                }
                if (!(FA_FUSED_SKIP & 2)) issue_dq_partial<D, kBf16>(tmem + kColDQ, aDS, aK);    // dQ_i partial = dS K
                tc_commit_e(dq_full);
                if (!FA_FUSED_DP_FIRST && more) {
                    dofull(gt + 1); tc_fence_after();
                    issue_scores_ts<D, kBf16>(tmem + kColDPT, tmem + kColVT, stage_addr(gt + 1) + C::kTileBytes); tc_commit_e(dp_full);   // dP^T(i+1)
                }
            }
            if (n_it == 0) mbar_wait(acc_empty, (ix & 1) ^ 1, 629);
            tc_commit_e(acc_full);                       // every MMA of the item is done -> accumulators final
            tc_commit_e(&kv_free[kb]);                   // ... and this K/V buffer is no longer read
            gi += n_it; git += n_it;
        }
    } else {
        // ------------------------------- math warpgroups -------------------------------
        reg_alloc<BwdRegs<D>::kCompute>();
        const int h = warp >> 2;                         // column half
        const int r = tid & 127;                         // kv row in tile == TMEM lane
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tST = tmem + lane_field + kColST + h * 64;
        const uint32_t tDPT = tmem + lane_field + kColDPT + h * 64;
        const uint32_t tP = tmem + lane_field + kColP + h * 32;
        const uint32_t tDQ = tmem + lane_field + kColDQ + h * 32;       // my 32 of the 64 dQ columns (lane = q row)
        const uint32_t sDSh = smem_u32(sDS) + h * 16384;                // my 64 q columns = one swizzled chunk
        const uint32_t sDQh = smem_u32(sDQ) + h * 16384;                // my fp32 box
        const uint32_t tKV = tmem + lane_field + (h ? kColVT : kColKT); // warpgroup A keeps K in TMEM, warpgroup B keeps V
        const float c2 = p.scale_log2;
        uint32_t gi = 0;
        bool store_pending = false;
        if (FA_FUSED_STAGGER && h == 1) named_bar_arrive(3, 256);       // warpgroup A takes the first turn
        // the dV/dK store of the previous item reads the staging that aliases the dS^T tile
        auto staging_free = [&]() {
            if (store_pending) {
                if (tid == 0) tma_store_wait_read0();
                named_bar_sync(1, 256);
                store_pending = false;
            }
        };
        // dQ partial of global iteration g: TMEM -> fp32 staging; the reducer warp adds it into global memory
        auto drain_dq = [&](uint32_t g) {
            mbar_wait(dq_full, g & 1, 634); tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tDQ, v); tc_wait_ld();
            if (g > 0) mbar_wait(dqs_empty, (g - 1) & 1, 635);          // the previous reduce has read the staging
            if (!(FA_FUSED_SKIP & 1)) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) sts128(sDQh + sw128_offset(r, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(dqs_full);
        };
        // K / V tile of item ix: shared memory (row r = 128 swizzled bytes) -> 32 TMEM columns of lane r.  Called once the
        // previous item's MMAs are complete (acc_full), i.e. nothing reads the TMEM copies any more.
        auto copy_kv = [&](uint32_t ix) {
            const uint32_t kb = ix & 1;
            mbar_wait(h ? &v_full[kb] : &k_full[kb], (ix >> 1) & 1, 637);
            const uint32_t src = smem_u32(sKV) + kb * 2 * C::kTileBytes + h * C::kTileBytes;
            #pragma unroll
            for (int q = 0; q < D / 32; ++q) {
                uint32_t w[16];
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = lds128(src + sw128_offset(r, q * 4 + j));
                    w[4 * j] = __float_as_uint(v.x); w[4 * j + 1] = __float_as_uint(v.y);
                    w[4 * j + 2] = __float_as_uint(v.z); w[4 * j + 3] = __float_as_uint(v.w);
                }
                tmem_st16(tKV + q * 16, w);
            }
            tc_wait_st(); tc_fence_before();
            mbar_arrive(kvt_full);
        };
        int item = next_item(0);
        if (item < n_items) copy_kv(0);
        for (uint32_t ix = 0;; ++ix) {
            if (item >= n_items) break;
            int bh, jt, i_start, i_end, n_it; decode(item, bh, jt, i_start, i_end, n_it);
            const int kv_g = jt * 128 + r;
            int q_lo = p.causal ? kv_g : 0, q_hi = p.Sq;  // queries that see my kv row
            if constexpr (kExt) {
                if (p.col_lo) {
                    const size_t ci = (size_t)(bh / p.Hk) * p.Sk + min(kv_g, p.Sk - 1);
                    q_lo = max(q_lo, __ldg(p.col_lo + ci)); q_hi = min(q_hi, __ldg(p.col_hi + ci));
                }
            }
            int qtile = i_start;
            uint32_t bhq = (uint32_t)((bh / p.Hk) * p.H + (bh % p.Hk) * p.G);   // batch*H + query head of the current iteration
            for (int it = 0; it < n_it; ++it) {
                const uint32_t g = gi + it;
                const uint32_t ss = g % C::kStatStages;
                const uint32_t stat = smem_u32(sStat) + ss * 1024 + h * 256;
                const int q0 = qtile * 128 + h * 64;                 // global query index of my column 0
                const uint32_t bhq_it = bhq;
                if (++qtile == i_end) { qtile = i_start; ++bhq; }
                mbar_wait(&stat_full[ss], (g / C::kStatStages) & 1, 630);
                mbar_wait(s_full, g & 1, 631);
                tc_fence_after();
                float pv[64];
                {
                    uint32_t s[2][32];
                    tmem_ld32(tST, s[0]); tmem_ld32(tST + 32, s[1]);
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(s_taken);
                    if (FA_FUSED_STAGGER) named_bar_sync(3 + h, 256);
                    const uint64_t c2v = pack_f2(c2, c2);
                    #pragma unroll
                    for (int c = 0; c < 64; c += 4) {
                        const float4 nl = lds128(stat + c * 4);
                        const uint64_t xa = ffma2(pack_u2(s[c >> 5][c & 31], s[c >> 5][(c & 31) + 1]), c2v, pack_f2(nl.x, nl.y));
                        const uint64_t xb = ffma2(pack_u2(s[c >> 5][(c & 31) + 2], s[c >> 5][(c & 31) + 3]), c2v, pack_f2(nl.z, nl.w));
                        if ((c & 15) < FA_FUSED_POLY) {
                            ex2_poly2(xa, pv[c], pv[c + 1]); ex2_poly2(xb, pv[c + 2], pv[c + 3]);
                        } else {
                            float x0, x1, x2, x3;
                            unpack_f2(xa, x0, x1); unpack_f2(xb, x2, x3);
                            if (FA_FUSED_STAGGER) {
                                pv[c] = ex2_approx_ordered(x0); pv[c + 1] = ex2_approx_ordered(x1);
                                pv[c + 2] = ex2_approx_ordered(x2); pv[c + 3] = ex2_approx_ordered(x3);
                            } else {
                                pv[c] = ex2_approx(x0); pv[c + 1] = ex2_approx(x1); pv[c + 2] = ex2_approx(x2); pv[c + 3] = ex2_approx(x3);
                            }
                        }
                    }
                    if (FA_FUSED_STAGGER) named_bar_arrive(4 - h, 256);
                }
                if (kExt ? (q0 < q_lo || q0 + 64 > q_hi) : (p.causal && q0 < kv_g)) {   // tile straddles the diagonal / a range end
                    const int cmin = q_lo - q0, cmax = kExt ? q_hi - 1 - q0 : 63;
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) if (c < cmin || c > cmax) pv[c] = 0.f;
                }
                uint64_t keep = ~0ull;                       // dropout keep bit per column (query) of my kv row
                constexpr bool drop = (kExt == 2);
                if constexpr (drop) {
                    keep = 0ull;
                    #pragma unroll
                    for (int c = 0; c < 64; ++c) {
                        const uint32_t w = dropout_word(dropout_row_key(p.drop.seed0, bhq_it, (uint32_t)(q0 + c)), p.drop.seed1, (uint32_t)kv_g >> 2);
                        keep |= (uint64_t)dropout_keep(w, (uint32_t)kv_g, p.drop.thresh) << c;
                    }
                }
                const float dscale = drop ? p.drop.scale : 1.f;
                if (it > 0) drain_dq(g - 1);                 // frees the shared TMEM region (its MMAs finished during the exp)
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t pk[16];
                    #pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = q * 32 + 2 * i;
                        if constexpr (drop)                    // dV sees the dropped-out, rescaled P^T; pv keeps P for dS
                            pk[i] = pack2<kBf16>(((keep >> c) & 1) ? pv[c] * dscale : 0.f, ((keep >> (c + 1)) & 1) ? pv[c + 1] * dscale : 0.f);
                        else
                            pk[i] = pack2<kBf16>(pv[c], pv[c + 1]);
                    }
                    tmem_st16(tP + q * 16, pk);
                }
                tc_wait_st(); tc_fence_before();
                mbar_arrive(p_full);
                mbar_wait(dp_full, g & 1, 632);
                tc_fence_after();
                {
                    uint32_t dp[2][32];
                    tmem_ld32(tDPT, dp[0]); tmem_ld32(tDPT + 32, dp[1]);
                    tc_wait_ld();
                    if (it == 0) staging_free();
                    #pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint32_t pk[16];
                        #pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const int c = q * 32 + 2 * i;
                            const float4 dl = lds128(stat + 512 + c * 4);
                            float d0, d1, d2, d3;        // dS = P o (dP - delta), packed: FADD2 + FMUL2
                            uint64_t dpa = pack_u2(dp[q][2 * i], dp[q][2 * i + 1]), dpb = pack_u2(dp[q][2 * i + 2], dp[q][2 * i + 3]);
                            if constexpr (drop) {
                                float e0, e1, e2, e3; unpack_f2(dpa, e0, e1); unpack_f2(dpb, e2, e3);
                                dpa = pack_f2(((keep >> c) & 1) ? e0 * dscale : 0.f, ((keep >> (c + 1)) & 1) ? e1 * dscale : 0.f);
                                dpb = pack_f2(((keep >> (c + 2)) & 1) ? e2 * dscale : 0.f, ((keep >> (c + 3)) & 1) ? e3 * dscale : 0.f);
                            }
                            unpack_f2(fmul2(pack_f2(pv[c], pv[c + 1]), fadd2(dpa, pack_f2(dl.x, dl.y))), d0, d1);
                            unpack_f2(fmul2(pack_f2(pv[c + 2], pv[c + 3]), fadd2(dpb, pack_f2(dl.z, dl.w))), d2, d3);
                            pk[i] = pack2<kBf16>(d0, d1); pk[i + 1] = pack2<kBf16>(d2, d3);
                        }
                        tmem_st16(tDPT + q * 16, pk);                        // A operand of dK (in place of dP^T)
                        if (!(FA_FUSED_SKIP & 4))
                        #pragma unroll
                        for (int j = 0; j < 4; ++j)                          // and the smem tile dQ reads as dS
                            sts128(sDSh + sw128_offset(r, q * 4 + j), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    }
                }
                tc_wait_st(); tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(ds_full);
                __syncwarp();
                if (lane_id() == 0) mbar_arrive(&stat_empty[ss]);
            }
            gi += n_it;
            if (n_it > 0) drain_dq(gi - 1);
            // ---- epilogue: dV, dK*scale -> 16-bit -> smem staging (over the dS^T tile) -> TMA store
            mbar_wait(acc_full, ix & 1, 633); tc_fence_after();
            const int item_next = next_item(ix + 1);         // before the epilogue: the MMA warp starts the next item's score
            if (item_next < n_items) copy_kv(ix + 1);        // MMAs under it
            staging_free();
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDV, sOutV, r, h, 1.0f, n_it == 0);
            stage_grad_half<D, kBf16>(tmem + lane_field + kColDK, sOutK, r, h, p.scale, n_it == 0);
            tc_fence_before();
            mbar_arrive(acc_empty);                          // TMEM accumulators drained
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (tid == 0) {
                tma_store_4d(&mapdV, sOutV, 0, jt * 128, bh % p.Hk, bh / p.Hk);
                tma_store_4d(&mapdK, sOutK, 0, jt * 128, bh % p.Hk, bh / p.Hk);
                tma_store_commit();
            }
            store_pending = true;
            item = item_next;
        }
        if (tid == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
    hang_trap_if_set();
}

// dq[b,h,s,:] = scale * accumulator[b,h,s,:]  (fp32 contiguous [B*H, Sq, D] -> 16-bit, any strides); four rows per thread,
// loads first (the accumulator usually still sits in L2 after the reduce-adds)
template <int D, bool kBf16>
__global__ void __launch_bounds__(256) fa_dq_convert_kernel(const float4* __restrict__ acc, uint4* __restrict__ dq, long long rows,
                                                            int H, int Sq, RowStrides sd, float scale) {
    constexpr int TPR = D / 8;
    constexpr int RPB = 256 / TPR;
    constexpr int U = 4;
    const int sub = threadIdx.x % TPR;
    const long long stride = (long long)gridDim.x * RPB;
    pdl_wait();
    for (long long row0 = (long long)blockIdx.x * RPB + threadIdx.x / TPR; row0 < rows; row0 += stride * U) {
        float4 a[U], b[U];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            if (row < rows) { a[u] = __ldg(acc + row * (D / 4) + sub * 2); b[u] = __ldg(acc + row * (D / 4) + sub * 2 + 1); }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = row0 + u * stride;
            if (row < rows) {
                const long long bh = row / Sq, sq = row % Sq, bb = bh / H, hh = bh % H;
                uint4 o;
                o.x = pack2<kBf16>(a[u].x * scale, a[u].y * scale); o.y = pack2<kBf16>(a[u].z * scale, a[u].w * scale);
                o.z = pack2<kBf16>(b[u].x * scale, b[u].y * scale); o.w = pack2<kBf16>(b[u].z * scale, b[u].w * scale);
                dq[((bb * sd.b + hh * sd.h + sq * sd.r) >> 3) + sub] = o;
            }
        }
    }
}

template <bool kBf16, int kExt>
int launch_bwd_fused_te(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                        const CUtensorMap& mdk, const CUtensorMap& mdv, const CUtensorMap& macc, const BwdParams& p,
                        const float* acc, void* dq, RowStrides s_dq, cudaStream_t st, int parts) {
    constexpr int D = 64;
    {
        cudaError_t e = ensure_smem<fa_bwd_fused_kernel<D, kBf16, kExt>>(FusedCfg<D>::kSmemBytes, p.dev);
        if (e != cudaSuccess) return (int)e;
    }
    const int items = (p.BH / p.G) * p.n_ktiles;
    const int grid = items < p.sms ? items : p.sms;
    cudaError_t e = cudaSuccess;
    if (parts & 8) e = launch_pdl(fa_bwd_fused_kernel<D, kBf16, kExt>, grid, kBwdThreads, FusedCfg<D>::kSmemBytes, st, mq, mk, mv, mdo, mdk, mdv, macc, p);
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

template <bool kBf16>
int launch_bwd_fused_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                       const CUtensorMap& mdk, const CUtensorMap& mdv, const CUtensorMap& macc, const BwdParams& p,
                       const float* acc, void* dq, RowStrides s_dq, cudaStream_t st, int parts) {
    if (p.drop.thresh) return launch_bwd_fused_te<kBf16, 2>(mq, mk, mv, mdo, mdk, mdv, macc, p, acc, dq, s_dq, st, parts);
    return p.col_lo ? launch_bwd_fused_te<kBf16, 1>(mq, mk, mv, mdo, mdk, mdv, macc, p, acc, dq, s_dq, st, parts)
                    : launch_bwd_fused_te<kBf16, 0>(mq, mk, mv, mdo, mdk, mdv, macc, p, acc, dq, s_dq, st, parts);
}

}  // namespace fa
