// Forward attention for sm_100a:  O = softmax(Q K^T * scale [+causal]) V,  LSE = m + ln(l).
// Replaces flash_attention_forward_kernel (reference code/_flash_attention_kernel_optimized.py:34-129)
// with a persistent, warp-specialized tcgen05 / TMA kernel.
//
// CTA = 384 threads (warp 11 idle; the third warpgroup donates registers via setmaxnreg), one CTA per SM, each work item = 256 query rows of one (batch, head):
//   warps 0-3  softmax warpgroup for Q tile 0 (rows 0..127 of the item; thread r <-> TMEM lane r)
//   warps 4-7  softmax warpgroup for Q tile 1
//   warp  8    MMA issuer (one thread): S_t = Q_t K^T, O_t += P_t V on tcgen05, accumulators in TMEM
//              (at D=64, where P has its own TMEM region, warp 10 issues for tile 1 and warp 8 for tile 0)
//   warp  9    TMA producer (one thread) + dynamic tile scheduler
// TMEM (512 cols): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D); P_t (16-bit) overwrites
// the first 64 columns of S_t and is consumed as the A operand straight from TMEM.
// The two tiles ping-pong: while warpgroup 0 runs softmax on S0(j+1), the tensor core runs
// P1(j) V(j) and S1(j+1).
// D=64 leaves 128 TMEM columns free: there P_t gets its own region (kSepP), so S_t(j+1) is issued as soon
// as S_t(j) sits in the softmax registers instead of after P_t(j) V(j) — the softmax warpgroups then
// never wait for the tensor pipe (the exp2 unit is the bound at D=64).
#pragma once
#include "fa_ptx.cuh"

namespace fa {

struct FwdParams {
    int BH, H, Sq, Sk;     // tensor maps are 4-D [B, H, S, D] with explicit strides: coordinates (col, row, h, b)
    int G;                 // query heads per K/V head (GQA/MQA; 1 = the reference's layout): kv head = h / G
    int n_qblk;            // ceil(Sq / 256)
    int n_items;           // BH * n_qblk
    int hc;                // heads per scheduling chunk (item_to_head_tile)
    int causal;
    float scale;           // softmax scale (1/sqrt(D) by default)
    float scale_log2;      // scale * log2(e)
    float* lse;            // [BH, Sq] fp32
    unsigned int* sched;   // work counter, zeroed before launch
    int dyn_first;         // draw the first item from the counter too (shared SMs, see sched_first)
    // Optional per-row key ranges [B, Sq] (var-len packing, key padding, windows): query row i of batch b sees keys
    // [row_lo, row_hi) (and, if causal, only keys <= i).  Both arrays must be non-decreasing in i.  NULL = [0, Sk).
    const int* row_lo;
    const int* row_hi;
    DropoutParams drop;    // thresh = 0: no dropout (extended instantiation only)
};

template <int D> struct FwdCfg {
    static constexpr int kChunks = D / 64;                 // 64-element (128-byte) column chunks
    static constexpr int kTileBytes = 128 * D * 2;         // one 128-row Q / K / V tile
    static constexpr int kStages = (D == 128) ? 4 : 6;     // K/V ring slots
    static constexpr int kOStageBytes = 128 * 128;         // [128 rows][64 cols] staging per warpgroup
    static constexpr int kOffQ = 0;
    static constexpr int kOffKV = 2 * kTileBytes;
    static constexpr int kOffO = kOffKV + kStages * kTileBytes;
    static constexpr int kOffBar = kOffO + 2 * kOStageBytes;
    static constexpr bool kSepP = (D == 64);               // P in its own TMEM columns [384,512)
    static constexpr int kNumBars = 2 + 2 + 2 * kStages + 2 + 2 + 2 + 2 + 2 + 2 + 2 + 2 + 2;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 64 + 1024;   // +1024: manual alignment slack
};

constexpr int kFwdThreads = 384;    // 3 warpgroups: softmax0, softmax1, {MMA, TMA, 2 idle warps}
// setmaxnreg split of the launch pool: the CTA owns 384 x 168 = 64512 registers, so 2 * softmax + other <= 504 per thread triple
// (a split that needs more than the pool makes setmaxnreg.inc wait forever).
// Measured: 216 / 72 at D=64 removes every spill but is 25-30 % SLOWER (the MMA / TMA warps' code degrades at 72), 208 / 88 at
// D=128 changes nothing -> 208 / 80 everywhere.
template <int D> struct FwdRegs { static constexpr int kSoftmax = 208, kOther = 80; };
static_assert(2 * FwdRegs<64>::kSoftmax + FwdRegs<64>::kOther <= 504 && 2 * FwdRegs<128>::kSoftmax + FwdRegs<128>::kOther <= 504, "register pool");
// Of every FA_FWD_POLY_DEN pairs of exponentials, FA_FWD_POLY_NUM are evaluated on the FMA pipe (ex2_poly2),
// the rest on the MUFU unit: the exp loop is XU-saturated (profiles/), the polynomial shifts load to FFMA2.
// Measured (profiles/r01_ab_fwd_poly_recheck.txt): 1/4 and more is slower (the loop turns issue-bound), 1/8 is -3 % on long D=64
// sequences and neutral elsewhere.
#ifndef FA_FWD_POLY_NUM
#define FA_FWD_POLY_NUM 1
#endif
#ifndef FA_FWD_POLY_DEN
#define FA_FWD_POLY_DEN 8
#endif
// The two softmax warpgroups do identical work and would otherwise run in lockstep, hitting the MUFU unit
// at the same time and leaving it idle at the same time.  Turn-taking around the exp loop (named barriers
// 3 and 4, the FlashAttention-3 "warp scheduler barrier" idea) staggers them: while one warpgroup
// exponentiates, the other loads / reduces / stores.  Rounds are counted per item as max(n0, n1) so that
// unequal iteration counts (causal, ragged) cannot deadlock.
// Measured (A/B on one box): D=64 -7 % time, D=128 +2 % (there the single in-order MMA issuer already offsets the
// two tiles by one S MMA) -> on only where P has its own region.
#ifndef FA_FWD_STAGGER
#define FA_FWD_STAGGER (C::kSepP)
#endif
// Three knobs on the softmax critical path (per K/V tile a warpgroup's chain is: S ready -> load -> row max -> exp -> P stored ->
// P V and the next S on the tensor pipe; at D=128 that chain, not a pipe, sets the period — profiles/r01: tensor 61 %, XU 54 %):
//  FA_FWD_MAX3      row max with the 3-input FMNMX3 (64 instead of 128 instructions ahead of the first exponential)
//  FA_FWD_SPLIT_LD  S is fetched in two halves; the max of the first runs under the TMEM load of the second
//  FA_FWD_SPLIT_P   (P aliases S, D=128) P is published in two halves: the first four K-steps of P V are issued while the
//                   second half of the exponentials is still being computed
#ifndef FA_FWD_MAX3
#define FA_FWD_MAX3 1
#endif
#ifndef FA_FWD_SPLIT_LD
#define FA_FWD_SPLIT_LD 1
#endif
#ifndef FA_FWD_SPLIT_P
#define FA_FWD_SPLIT_P 0        // measured: no change (profiles/r02_ab_fwd_knobs_negative_result.jsonl); excluded by FA_FWD_SPEC
#endif
//  FA_FWD_SPEC      speculative exponentials against the stale row maximum (see the softmax loop): 1 = where P aliases S (D = 128),
//                   2 = every head dim, 0 = off
// Measured (profiles/r02_ab_fwd_speculative_max_negative_result.jsonl): correct (102 GPU tests green) but 6-7 % SLOWER at D = 128
// (C3 fwd 0.443 -> 0.473 ms) and 14-18 % slower at D = 64 — the load + max were not the exposed part of the chain -> off.
#ifndef FA_FWD_SPEC
#define FA_FWD_SPEC 0
#endif
static_assert(!(FA_FWD_SPEC && FA_FWD_SPLIT_P), "a speculative tile may be redone: P cannot be published in halves");
constexpr float kLazyRescaleLog2 = 8.0f;   // rescale O only when the row max grows by > 2^8 in exp2 units

// iterations (128-wide K/V tiles) that tile `t` of the item starting at row q0 must visit
__device__ __forceinline__ int fwd_tile_iters(int q0, int t, int Sq, int Sk, int causal) {
    const int r0 = q0 + t * 128;
    if (r0 >= Sq) return 0;
    const int nkv = (Sk + 127) >> 7;
    if (!causal) return nkv;
    const int last_row = min(r0 + 127, Sq - 1);          // top-left aligned: row i sees cols <= i
    const int n = (min(last_row, Sk - 1) >> 7) + 1;
    return min(n, nkv);
}

// Same with per-row key ranges: `jb` = first K/V tile of the item (from the range of the item's first row; ranges are
// monotone, so it is the minimum), return = number of tiles from jb that tile `t` must visit (up to the range end of its last row).
// Rows whose own range starts later / ends earlier are handled by the element mask.
__device__ __forceinline__ int fwd_item_iters(const int* row_lo, const int* row_hi, int b, int q0, int t, int Sq, int Sk,
                                              int causal, int& jb) {
    if (!row_lo) { jb = 0; return fwd_tile_iters(q0, t, Sq, Sk, causal); }
    jb = __ldg(row_lo + (size_t)b * Sq + min(q0, Sq - 1)) >> 7;
    const int r0 = q0 + t * 128;
    if (r0 >= Sq) return 0;
    const int last_row = min(r0 + 127, Sq - 1);
    int hi = min(__ldg(row_hi + (size_t)b * Sq + last_row), Sk);
    if (causal) hi = min(hi, last_row + 1);
    return max(((hi - 1) >> 7) - jb + 1, 0);
}

// kRanges / kDropout: separate instantiations for the per-row key ranges and for dropout.  The plain operator keeps its own code and
// register budget (the softmax loop sits at the 208-register ceiling and three more live values spill), and the range-masked one
// must not carry the dropout generator: as a run-time branch inside the exp loop it was if-converted and made the packed
// variable-length forward 2.2x slower with dropout off.
template <int D, bool kBf16, bool kRanges = false, bool kDropout = false>
__global__ void __launch_bounds__(kFwdThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapO,
              const FwdParams p) {
    using C = FwdCfg<D>;
    const int* const rlo = kRanges ? p.row_lo : nullptr;       // compile-time NULL in the plain instantiation
    const int* const rhi = kRanges ? p.row_hi : nullptr;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + C::kOffQ;
    uint8_t* sKV = smem + C::kOffKV;
    uint8_t* sO = smem + C::kOffO;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* q_full = bars;                    // [2]  TMA -> MMA
    uint64_t* q_empty = bars + 2;               // [2]  MMA -> TMA
    uint64_t* kv_full = bars + 4;               // [kStages]
    uint64_t* kv_empty = kv_full + C::kStages;  // [kStages]
    uint64_t* s_full = kv_empty + C::kStages;   // [2]  MMA -> softmax (S_t ready; also: all earlier MMAs done)
    uint64_t* p_full = s_full + 2;              // [2]  softmax -> MMA (P_t in TMEM, O_t rescaled)
    uint64_t* o_full = p_full + 2;              // [2]  MMA -> softmax (last P V of the item done)
    uint64_t* o_empty = o_full + 2;             // [2]  softmax -> MMA (O_t drained from TMEM)
    uint64_t* s_empty = o_empty + 2;            // [2]  softmax -> MMA (S_t is in registers)           (kSepP)
    uint64_t* pv_done = s_empty + 2;            // [2]  MMA -> softmax (P_t V of this iteration done)  (kSepP)
    uint64_t* p_half = pv_done + 2;             // [2]  softmax -> MMA (first half of P_t in TMEM)        (FA_FWD_SPLIT_P)
    uint64_t* sched_full = p_half + 2;          // [2]
    uint64_t* sched_empty = sched_full + 2;     // [2]
    volatile int* sched_item = reinterpret_cast<volatile int*>(sched_empty + 2);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_item) + 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 128);
            mbar_init(&s_empty[i], 128); mbar_init(&pv_done[i], 1); mbar_init(&p_half[i], 128);
            mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], C::kSepP ? 10 : 9);  // MMA thread(s) + 8 softmax warps
        }
        for (int i = 0; i < C::kStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], C::kSepP ? 2 : 1); }   // released by every MMA thread
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 11) {
        reg_dealloc<FwdRegs<D>::kOther>();     // idle warp (register donor)
    } else if (warp == 9) {
        // ================================ TMA producer + scheduler ================================
        reg_dealloc<FwdRegs<D>::kOther>();
        {   // whole warp, converged; single-lane instructions elect their leader (fa_ptx.cuh)
            if (lane_id() == 0) { tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); }
            __syncwarp();
            uint32_t kv_cnt = 0;        // K/V tiles produced so far
            uint32_t q_cnt0 = 0, q_cnt1 = 0;
            int item = sched_first(p.sched, p.dyn_first);
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it & 1;
                mbar_wait(&sched_empty[slot], ((it >> 1) & 1) ^ 1, 100);
                sched_item[slot] = item;
                mbar_arrive_e(&sched_full[slot]);
                if (item >= p.n_items) break;
                int bh, qt; item_to_head_tile(item, p.BH, p.n_qblk, p.hc, bh, qt);
                const int q0 = (p.n_qblk - 1 - qt) * 256;                   // heavy (late) query blocks first
                int jb;
                const int n0 = fwd_item_iters(rlo, rhi, bh / p.H, q0, 0, p.Sq, p.Sk, p.causal, jb);
                const int n1 = fwd_item_iters(rlo, rhi, bh / p.H, q0, 1, p.Sq, p.Sk, p.causal, jb);
                const int n = max(n0, n1);
                auto load_q = [&](int t, uint32_t& cnt) {
                    mbar_wait(&q_empty[t], (cnt & 1) ^ 1, 101 + t);
                    mbar_arrive_expect_tx_e(&q_full[t], C::kTileBytes);
                    #pragma unroll
                    for (int c = 0; c < C::kChunks; ++c)
                        tma_load_4d_e(sQ + t * C::kTileBytes + c * 16384, &mapQ, &q_full[t], c * 64, q0 + t * 128, bh % p.H, bh / p.H);
                    ++cnt;
                };
                auto load_kv = [&](const CUtensorMap* m, int j) {
                    const uint32_t st = kv_cnt % C::kStages;
                    mbar_wait(&kv_empty[st], ((kv_cnt / C::kStages) & 1) ^ 1, 110);
                    mbar_arrive_expect_tx_e(&kv_full[st], C::kTileBytes);
                    #pragma unroll
                    for (int c = 0; c < C::kChunks; ++c)
                        tma_load_4d_e(sKV + st * C::kTileBytes + c * 16384, m, &kv_full[st], c * 64, (jb + j) * 128, (bh % p.H) / p.G, bh / p.H);
                    ++kv_cnt;
                };
                if (n > 0) {                              // n == 0: no row of the item sees a key (range masks) -> nothing to load
                    if (n0 > 0) load_q(0, q_cnt0);
                    load_kv(&mapK, 0);
                    if (n1 > 0) load_q(1, q_cnt1);
                    load_kv(&mapV, 0);
                    for (int j = 1; j < n; ++j) { load_kv(&mapK, j); load_kv(&mapV, j); }
                }
                item = sched_next(p.sched, p.dyn_first);
            }
            if (lane_id() == 0) sched_retire(p.sched);
        }
    } else if (warp == 8 || warp == 10) {
        // ===================================== MMA issuers =====================================
        // One issuing thread PER Q tile (warp 8 -> tile 0, warp 10 -> tile 1).  The two tiles' MMA chains
        // are independent (own TMEM regions, read-only K/V), so neither tile ever waits behind a barrier
        // that belongs to the other.  Each K/V ring slot is released by both threads (kv_empty count 2).
        reg_dealloc<FwdRegs<D>::kOther>();
        if constexpr (C::kSepP) {
        {   // whole warp, converged
            const int t = (warp == 8) ? 0 : 1;
            constexpr uint32_t idesc_s = make_idesc(kBf16, false, false, 128, 128);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, false, true, 128, D);
            const uint32_t sq_addr = smem_u32(sQ) + t * C::kTileBytes, skv_addr = smem_u32(sKV);
            const uint32_t tS = tmem + t * 128, tO = tmem + 256 + t * D;
            const uint32_t tP = C::kSepP ? tmem + 384 + t * 64 : tS;
            uint32_t kv_cnt = 0;                           // ring element index of K(0) of the current item
            uint32_t ph_q = 0, ph_p = 0, ph_oe = 0, ph_se = 0;
            auto kv_wait = [&](uint32_t cnt) {
                mbar_wait(&kv_full[cnt % C::kStages], (cnt / C::kStages) & 1, 200);
            };
            auto issue_s = [&](uint32_t st) {              // S_t = Q_t K^T
                const uint32_t b = skv_addr + st * C::kTileBytes;
                #pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
                    umma_ss_e(tS, make_smem_desc(sq_addr + off, 0, 1024), make_smem_desc(b + off, 0, 1024), idesc_s, k > 0);
                }
            };
            auto issue_pv = [&](uint32_t st, bool acc) {   // O_t (+)= P_t V
                const uint32_t b = skv_addr + st * C::kTileBytes;
                #pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_ts_e(tO, tP + k * 8, make_smem_desc(b + k * 2048, 16384, 1024), idesc_pv, acc || k > 0);
            };
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it & 1;
                mbar_wait(&sched_full[slot], (it >> 1) & 1, 201);
                const int item = __shfl_sync(0xffffffffu, sched_item[slot], 0);   // warp-uniform for the compiler
                mbar_arrive_e(&sched_empty[slot]);
                if (item >= p.n_items) break;
                int bh_, qt, jb_; item_to_head_tile(item, p.BH, p.n_qblk, p.hc, bh_, qt);
                const int q0 = (p.n_qblk - 1 - qt) * 256;
                const int n0 = fwd_item_iters(rlo, rhi, bh_ / p.H, q0, 0, p.Sq, p.Sk, p.causal, jb_);
                const int n1 = fwd_item_iters(rlo, rhi, bh_ / p.H, q0, 1, p.Sq, p.Sk, p.causal, jb_);
                const int n = max(n0, n1), nt = t ? n1 : n0;
                // ring elements of this item: K(j) = kv_cnt + 2j, V(j) = kv_cnt + 2j + 1
                auto do_s = [&](int j) {                   // consume K(j): S_t(j) if this tile needs it
                    const uint32_t cnt = kv_cnt + 2 * j, st = cnt % C::kStages;
                    kv_wait(cnt);
                    if (j < nt) {
                        if (j == 0) { mbar_wait(&q_full[t], ph_q, 202); ph_q ^= 1; }
                        else if (C::kSepP) { mbar_wait(&s_empty[t], ph_se, 208); ph_se ^= 1; }   // S_t(j-1) is in registers
                        tc_fence_after();
                        issue_s(st); tc_commit_e(&s_full[t]);
                        if (j == nt - 1) tc_commit_e(&q_empty[t]);
                        tc_commit_e(&kv_empty[st]);
                    } else {
                        mbar_arrive_e(&kv_empty[st]);
                    }
                };
                auto do_pv = [&](int j) {                  // consume V(j): O_t += P_t(j) V(j)
                    const uint32_t cnt = kv_cnt + 2 * j + 1, st = cnt % C::kStages;
                    kv_wait(cnt);
                    if (j < nt) {
                        mbar_wait(&p_full[t], ph_p, 204); ph_p ^= 1;
                        if (j == 0) { mbar_wait(&o_empty[t], ph_oe ^ 1, 205); ph_oe ^= 1; }
                        tc_fence_after();
                        issue_pv(st, j > 0);
                        if (C::kSepP) tc_commit_e(&pv_done[t]);
                        else if (j == nt - 1) tc_commit_e(&o_full[t]);
                        tc_commit_e(&kv_empty[st]);
                    } else {
                        mbar_arrive_e(&kv_empty[st]);
                    }
                };
                if (n > 0) do_s(0);
                for (int j = 0; j < n; ++j) {
                    if (C::kSepP) {                        // S_t(j+1) does not depend on P_t(j) V(j): issue it first
                        if (j + 1 < n) do_s(j + 1);
                        do_pv(j);
                    } else {                               // P_t aliases S_t: in-order after P_t(j) V(j)
                        do_pv(j);
                        if (j + 1 < n) do_s(j + 1);
                    }
                }
                if (C::kSepP && nt > 0) ph_se ^= 1;        // the arrival for the item's last S_t is never waited on
                kv_cnt += 2 * n;
            }
        }
        } else {
        // P_t aliases S_t (TMEM is full at D=128): ONE issuer keeps the strict order P0V, S0, P1V, S1 — interleaving
        // two tiles' long MMAs at instruction granularity costs tensor throughput (measured 1105 -> 855 TFLOP/s)
        if (warp == 8) {
        {   // whole warp, converged
            constexpr uint32_t idesc_s = make_idesc(kBf16, false, false, 128, 128);
            constexpr uint32_t idesc_pv = make_idesc(kBf16, false, true, 128, D);
            const uint32_t sq_addr = smem_u32(sQ), skv_addr = smem_u32(sKV);
            uint32_t kv_cnt = 0;
            uint32_t ph_q = 0, ph_p = 0, ph_oe = 0;      // bit t = parity to wait for next
            auto kv_wait = [&](uint32_t cnt) {
                mbar_wait(&kv_full[cnt % C::kStages], (cnt / C::kStages) & 1, 200);
            };
            auto issue_s = [&](int t, uint32_t st) {   // S_t = Q_t K^T
                const uint32_t a = sq_addr + t * C::kTileBytes, b = skv_addr + st * C::kTileBytes;
                #pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
                    umma_ss_e(tmem + t * 128, make_smem_desc(a + off, 0, 1024), make_smem_desc(b + off, 0, 1024),
                            idesc_s, k > 0);
                }
            };
            auto issue_pv_k = [&](int t, uint32_t st, bool acc, int k0, int k1) {   // O_t (+)= P_t V, K-steps [k0, k1)
                const uint32_t b = skv_addr + st * C::kTileBytes;
                #pragma unroll
                for (int k = k0; k < k1; ++k)
                    umma_ts_e(tmem + 256 + t * D, tmem + t * 128 + k * 8, make_smem_desc(b + k * 2048, 16384, 1024),
                            idesc_pv, acc || k > 0);
            };
            uint32_t ph_h = 0;                           // bit t = parity of p_half[t] to wait for next
            // P_t(j) V(j): with FA_FWD_SPLIT_P the first four K-steps go out as soon as the first half of P_t is in TMEM
            auto do_pv = [&](int t, uint32_t st, int j, uint32_t& ph_pp, uint32_t& ph_oo) {
                const uint32_t bit = 1u << t;
                if (FA_FWD_SPLIT_P) {
                    mbar_wait(&p_half[t], (ph_h >> t) & 1, 209 + t); ph_h ^= bit;
                    if (j == 0) { mbar_wait(&o_empty[t], ((ph_oo >> t) & 1) ^ 1, 205 + 2 * t); ph_oo ^= bit; }
                    tc_fence_after();
                    issue_pv_k(t, st, j > 0, 0, 4);
                    mbar_wait(&p_full[t], (ph_pp >> t) & 1, 204 + 2 * t); ph_pp ^= bit;
                    tc_fence_after();
                    issue_pv_k(t, st, true, 4, 8);
                } else {
                    mbar_wait(&p_full[t], (ph_pp >> t) & 1, 204 + 2 * t); ph_pp ^= bit;
                    if (j == 0) { mbar_wait(&o_empty[t], ((ph_oo >> t) & 1) ^ 1, 205 + 2 * t); ph_oo ^= bit; }
                    tc_fence_after();
                    issue_pv_k(t, st, j > 0, 0, 8);
                }
            };
            for (uint32_t it = 0;; ++it) {
                const uint32_t slot = it & 1;
                mbar_wait(&sched_full[slot], (it >> 1) & 1, 201);
                const int item = __shfl_sync(0xffffffffu, sched_item[slot], 0);   // warp-uniform for the compiler
                mbar_arrive_e(&sched_empty[slot]);
                if (item >= p.n_items) break;
                int bh_, qt, jb_; item_to_head_tile(item, p.BH, p.n_qblk, p.hc, bh_, qt);
                const int q0 = (p.n_qblk - 1 - qt) * 256;
                const int n0 = fwd_item_iters(rlo, rhi, bh_ / p.H, q0, 0, p.Sq, p.Sk, p.causal, jb_);
                const int n1 = fwd_item_iters(rlo, rhi, bh_ / p.H, q0, 1, p.Sq, p.Sk, p.causal, jb_);
                const int n = max(n0, n1);
                // ---- S_t(0)
                if (n > 0) {
                    const uint32_t st = kv_cnt % C::kStages;
                    kv_wait(kv_cnt);
                    if (n0 > 0) {
                        mbar_wait(&q_full[0], ph_q & 1, 202); ph_q ^= 1; tc_fence_after();
                        issue_s(0, st); tc_commit_e(&s_full[0]);
                        if (n0 == 1) tc_commit_e(&q_empty[0]);
                    }
                    if (n1 > 0) {
                        mbar_wait(&q_full[1], (ph_q >> 1) & 1, 203); ph_q ^= 2; tc_fence_after();
                        issue_s(1, st); tc_commit_e(&s_full[1]);
                        if (n1 == 1) tc_commit_e(&q_empty[1]);
                    }
                    tc_commit_e(&kv_empty[st]); ++kv_cnt;
                }
                for (int j = 0; j < n; ++j) {
                        const uint32_t vst = kv_cnt % C::kStages;
                        const uint32_t kst = (kv_cnt + 1) % C::kStages;
                        const bool more = (j + 1 < n);
                        kv_wait(kv_cnt);                       // V(j)
                        if (j < n0) {
                            do_pv(0, vst, j, ph_p, ph_oe);
                            if (j == n0 - 1) tc_commit_e(&o_full[0]);
                        }
                        if (more) kv_wait(kv_cnt + 1);         // K(j+1)
                        if (j + 1 < n0) {
                            tc_fence_after();
                            issue_s(0, kst); tc_commit_e(&s_full[0]);
                            if (j + 1 == n0 - 1) tc_commit_e(&q_empty[0]);
                        }
                        if (j < n1) {
                            do_pv(1, vst, j, ph_p, ph_oe);
                            if (j == n1 - 1) tc_commit_e(&o_full[1]);
                        }
                        tc_commit_e(&kv_empty[vst]); ++kv_cnt;
                        if (more) {
                            if (j + 1 < n1) {
                                issue_s(1, kst); tc_commit_e(&s_full[1]);
                                if (j + 1 == n1 - 1) tc_commit_e(&q_empty[1]);
                            }
                            tc_commit_e(&kv_empty[kst]); ++kv_cnt;
                        }
                    }
            }
        }
        }
        }
    } else {
        // ================================= softmax warpgroups (0,1) ================================
        reg_alloc<FwdRegs<D>::kSoftmax>();
        const int t = warp >> 2;                       // Q tile handled by this warpgroup
        const int r = tid & 127;                       // row inside the tile == TMEM lane
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + lane_field + t * 128;
        const uint32_t tO = tmem + lane_field + 256 + t * D;
        const uint32_t tP = C::kSepP ? tmem + lane_field + 384 + t * 64 : tS;    // 16-bit P (A operand of P V)
        uint8_t* sOt = sO + t * C::kOStageBytes;
        uint32_t ph_s = 0, ph_o = 0, ph_pv = 0;
        const float c2 = p.scale_log2;
        constexpr bool kSpec = (FA_FWD_SPEC == 2) || (FA_FWD_SPEC == 1 && !C::kSepP);
        if (FA_FWD_STAGGER && t == 1) named_bar_arrive(3, 256);        // warpgroup 0 takes the first turn
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it & 1;
            mbar_wait(&sched_full[slot], (it >> 1) & 1, 300);
            const int item = sched_item[slot];
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&sched_empty[slot]);
            if (item >= p.n_items) break;
            int bh, qt; item_to_head_tile(item, p.BH, p.n_qblk, p.hc, bh, qt);
            const int q0 = (p.n_qblk - 1 - qt) * 256;
            int jb;
            const int nt = fwd_item_iters(rlo, rhi, bh / p.H, q0, t, p.Sq, p.Sk, p.causal, jb);
            const int n_rounds = max(nt, fwd_item_iters(rlo, rhi, bh / p.H, q0, 1 - t, p.Sq, p.Sk, p.causal, jb));
            const int row_g = q0 + t * 128 + r;
            if (nt == 0) {
                if (FA_FWD_STAGGER) for (int j = 0; j < n_rounds; ++j) { named_bar_sync(3 + t, 256); named_bar_arrive(4 - t, 256); }
                if constexpr (kRanges) {
                    if (q0 + t * 128 < p.Sq) {             // rows whose key range is empty: O = 0, LSE = -inf (the outputs are torch.empty)
                        #pragma unroll
                        for (int g = 0; g < 8; ++g) sts128(smem_u32(sOt) + sw128_offset(r, g), 0u, 0u, 0u, 0u);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + t, 128);
                        if (r == 0) {
                            #pragma unroll
                            for (int c = 0; c < C::kChunks; ++c) tma_store_4d(&mapO, sOt, c * 64, q0 + t * 128, bh % p.H, bh / p.H);
                            tma_store_commit();
                            tma_store_wait_read0();
                        }
                        named_bar_sync(1 + t, 128);
                        if (row_g < p.Sq) p.lse[(size_t)bh * p.Sq + row_g] = -INFINITY;
                    }
                }
                continue;
            }
            int k_lo = 0, k_hi = p.Sk;                    // keys this row may see (before the causal clip)
            if constexpr (kRanges) {
                const size_t ri = (size_t)(bh / p.H) * p.Sq + min(row_g, p.Sq - 1);
                k_lo = __ldg(p.row_lo + ri); k_hi = min(__ldg(p.row_hi + ri), p.Sk);
            }
            const uint32_t drop_key = kDropout ? dropout_row_key(p.drop.seed0, (uint32_t)bh, (uint32_t)row_g) : 0u;
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < nt; ++j) {
                mbar_wait(&s_full[t], ph_s, 301); ph_s ^= 1;
                tc_fence_after();
                uint32_t s[4][32];
                // element mask only on tiles that straddle the diagonal or the end of K
                const int kbase = (kRanges ? jb + j : j) * 128;
                int cmax = (kRanges ? k_hi : p.Sk) - 1 - kbase;             // last valid column in this tile
                if (p.causal) cmax = min(cmax, row_g - kbase);
                const int cmin = kRanges ? k_lo - kbase : 0;                // first valid column (> 0 only with row ranges)
                auto mask_half = [&](int q0) {
                    if (cmax < 127 || cmin > 0) {
                        #pragma unroll
                        for (int q = q0; q < q0 + 2; ++q)
                            #pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (q * 32 + i > cmax || q * 32 + i < cmin) s[q][i] = 0xff800000u;  // -inf
                    }
                };
                auto max_half = [&](int q0, float& a, float& b) {
                    a = -INFINITY; b = -INFINITY;
                    if (FA_FWD_MAX3) {
                        #pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            a = fmax3(a, __uint_as_float(s[q0][i]), __uint_as_float(s[q0][i + 1]));
                            b = fmax3(b, __uint_as_float(s[q0 + 1][i]), __uint_as_float(s[q0 + 1][i + 1]));
                        }
                    } else {
                        #pragma unroll
                        for (int i = 0; i < 32; ++i) { a = fmaxf(a, __uint_as_float(s[q0][i])); b = fmaxf(b, __uint_as_float(s[q0 + 1][i])); }
                    }
                };
                // exponentials of quarters [q_lo, q_hi) of the tile against the row maximum `mm`: p -> 16-bit -> TMEM, row sum into
                // the two pair accumulators (fp32 p before rounding, ref :111)
                const uint64_t c2v = pack_f2(c2, c2);
                uint64_t lA = pack_f2(0.f, 0.f), lB = lA;
                auto exp_quarters = [&](int q_lo, int q_hi, float mm) {
                    const float neg_mc = (mm == -INFINITY) ? 0.f : -mm * c2;
                    const uint64_t nmv = pack_f2(neg_mc, neg_mc);
                    #pragma unroll
                    for (int q = q_lo; q < q_hi; ++q) {
                        uint32_t pk[16];
                        #pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const uint64_t xv = ffma2(pack_u2(s[q][2 * i], s[q][2 * i + 1]), c2v, nmv);
                            float p0, p1;
                            if ((i % FA_FWD_POLY_DEN) < FA_FWD_POLY_NUM) {
                                ex2_poly2(xv, p0, p1);
                            } else {
                                float x0, x1; unpack_f2(xv, x0, x1);
                                p0 = ex2_approx(x0); p1 = ex2_approx(x1);
                            }
                            if (i & 1) lB = fadd2(lB, pack_f2(p0, p1)); else lA = fadd2(lA, pack_f2(p0, p1));   // l: before dropout
                            if constexpr (kDropout) {              // O accumulates the dropped-out, rescaled P; LSE / l do not
                                const uint32_t kcol = (uint32_t)(kbase + q * 32 + 2 * i);
                                const uint32_t w = dropout_word(drop_key, p.drop.seed1, kcol >> 2);
                                p0 = dropout_keep(w, kcol, p.drop.thresh) ? p0 * p.drop.scale : 0.f;
                                p1 = dropout_keep(w, kcol + 1, p.drop.thresh) ? p1 * p.drop.scale : 0.f;
                            }
                            pk[i] = pack2<kBf16>(p0, p1);
                        }
                        if (FA_FWD_SPLIT_P && !C::kSepP && q == 2) {   // columns 0..63 of P_t were stored a quarter of the exponentials ago:
                            tc_wait_st(); tc_fence_before();           // release the first four K-steps of P_t V while the rest is computed
                            mbar_arrive(&p_half[t]);
                        }
                        tmem_st16(tP + q * 16, pk);
                    }
                };
                // O_t and l rescaled from row maximum m to m_new (warp-collective TMEM traffic)
                auto rescale_to = [&](float m_new) {
                    // m = -inf -> 0.  A row that is STILL fully masked (range masks: its keys start tiles later) is rescaled with
                    // its warp (the decision is warp-uniform): -inf - -inf would poison l and O with NaN
                    const float corr = (m_new == -INFINITY) ? 1.f : ex2_approx((m - m_new) * c2);
                    l *= corr;
                    #pragma unroll
                    for (int q = 0; q < D / 32; ++q) {
                        uint32_t o[32];
                        tmem_ld32(tO + q * 32, o);
                        tc_wait_ld();
                        #pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st32(tO + q * 32, o);
                    }
                    m = m_new;
                };
                float mx0, mx1, mx2, mx3;
                tmem_ld32(tS, s[0]); tmem_ld32(tS + 32, s[1]);
                if (FA_FWD_SPLIT_LD || kSpec) tc_wait_ld();
                tmem_ld32(tS + 64, s[2]); tmem_ld32(tS + 96, s[3]);
                if (kSpec && j > 0) {
                    // Speculative tile (every tile but the first): the exponentials start against the STALE row maximum as soon as the
                    // first half of S_t is in registers — the TMEM load of the second half and the whole max reduction leave the
                    // S -> P critical path (it, not a pipe, sets the period at D = 128: the softmax warps spend 37 % of their time
                    // waiting for S_t, profiles/r02).  The tile's true maximum is reduced alongside; if it exceeds the stale one by
                    // more than the lazy-rescale threshold (rare after the first tiles), O_t / l are rescaled and the tile is redone
                    // with the new maximum, so P never exceeds 2^threshold (fp16-safe) — the same bound as before.
                    if constexpr (C::kSepP) { mbar_wait(&pv_done[t], ph_pv, 303); ph_pv ^= 1; tc_fence_after(); }
                    mask_half(0);
                    if (FA_FWD_STAGGER) named_bar_sync(3 + t, 256);
                    exp_quarters(0, 2, m);
                    max_half(0, mx0, mx1);
                    tc_wait_ld();
                    if constexpr (C::kSepP) { tc_fence_before(); mbar_arrive(&s_empty[t]); }
                    mask_half(2);
                    exp_quarters(2, 4, m);
                    max_half(2, mx2, mx3);
                    const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                    const bool need = (m_tile - m) * c2 > kLazyRescaleLog2 || (m == -INFINITY && m_tile != -INFINITY);
                    if (__any_sync(0xffffffffu, need)) {
                        tc_wait_st();                              // the speculative P stores are ordered before the redo's
                        rescale_to(fmaxf(m, m_tile));
                        lA = pack_f2(0.f, 0.f); lB = lA;
                        exp_quarters(0, 4, m);
                    }
                    if (FA_FWD_STAGGER) named_bar_arrive(4 - t, 256);
                } else {
                    if (FA_FWD_SPLIT_LD && !kSpec) { mask_half(0); max_half(0, mx0, mx1); }    // under the TMEM load of the second half
                    tc_wait_ld();
                    if constexpr (C::kSepP) { tc_fence_before(); mbar_arrive(&s_empty[t]); }   // S_t(j+1) may overwrite S_t now
                    if (!FA_FWD_SPLIT_LD || kSpec) { mask_half(0); max_half(0, mx0, mx1); }
                    mask_half(2); max_half(2, mx2, mx3);
                    const float m_new = FA_FWD_MAX3 ? fmaxf(fmax3(m, mx0, mx1), fmaxf(mx2, mx3)) : fmaxf(fmaxf(m, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
                    if (j == 0) {
                        m = m_new;
                    } else {
                        if constexpr (C::kSepP) {      // P_t(j-1) V(j-1) must be complete before O_t or P_t is touched
                            mbar_wait(&pv_done[t], ph_pv, 303); ph_pv ^= 1; tc_fence_after();
                        }
                        // lazy rescale (warp-uniform decision: tcgen05.ld/st are warp-collective)
                        const bool need = (m_new - m) * c2 > kLazyRescaleLog2;
                        if (__any_sync(0xffffffffu, need)) rescale_to(m_new);
                    }
                    if (FA_FWD_STAGGER) named_bar_sync(3 + t, 256);              // my turn on the exp unit
                    exp_quarters(0, 4, m);
                    if (FA_FWD_STAGGER) named_bar_arrive(4 - t, 256);            // hand the turn to the other warpgroup
                }
                {
                    float a0, a1; unpack_f2(fadd2(lA, lB), a0, a1);
                    l += a0 + a1;
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(&p_full[t]);
            }
            if (FA_FWD_STAGGER) for (int j = nt; j < n_rounds; ++j) { named_bar_sync(3 + t, 256); named_bar_arrive(4 - t, 256); }
            // ---------------- epilogue: O = o / l, LSE = m*scale + ln(l) ----------------
            if constexpr (C::kSepP) { mbar_wait(&pv_done[t], ph_pv, 302); ph_pv ^= 1; }
            else { mbar_wait(&o_full[t], ph_o, 302); ph_o ^= 1; }
            tc_fence_after();
            const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
            #pragma unroll
            for (int c = 0; c < C::kChunks; ++c) {
                uint32_t o[2][32];
                tmem_ld32(tO + c * 64, o[0]);
                tmem_ld32(tO + c * 64 + 32, o[1]);
                tc_wait_ld();
                if (c == C::kChunks - 1) { tc_fence_before(); mbar_arrive(&o_empty[t]); }
                #pragma unroll
                for (int g = 0; g < 8; ++g) {          // 8 x 16-byte groups = 64 columns
                    uint32_t w[4];
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = g * 8 + 2 * i;
                        w[i] = pack2<kBf16>(__uint_as_float(o[e >> 5][e & 31]) * inv_l,
                                            __uint_as_float(o[(e + 1) >> 5][(e + 1) & 31]) * inv_l);
                    }
                    sts128(smem_u32(sOt) + sw128_offset(r, g), w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                named_bar_sync(1 + t, 128);
                if (r == 0) {
                    tma_store_4d(&mapO, sOt, c * 64, q0 + t * 128, bh % p.H, bh / p.H);
                    tma_store_commit();
                    tma_store_wait_read0();
                }
                named_bar_sync(1 + t, 128);
            }
            if (row_g < p.Sq)
                p.lse[(size_t)bh * p.Sq + row_g] = (l > 0.f) ? fmaf(m, p.scale, __logf(l)) : -INFINITY;
        }
        if (r == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
    hang_trap_if_set();
}

}  // namespace fa
