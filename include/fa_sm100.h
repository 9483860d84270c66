/* C ABI of libfa_sm100.so — the drop-in boundary beneath the reference's Python operator.
 *
 * The reference has no FFI: its "plugin API" is the Python operator in
 * /root/reference/code/My_FlashAttention_optimized.py and its device code is Triton JIT.
 * Each entry point below replaces one reference launcher (file:line given per function); the
 * Python mirror of the operator (flashattention-from-scratch-with-triton_b200/interface.py)
 * binds them with ctypes.  INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *  - every pointer is a DEVICE pointer owned by the caller; nothing is retained after return
 *  - tensors are contiguous [B, H, S, D] (D innermost), 16-byte aligned; K and V share B, H, D with Q
 *  - dtype: 0 = fp16, 1 = bf16 (inputs, outputs and tensor-core operands); statistics are fp32
 *  - work is enqueued on `stream` (a cudaStream_t) of the CURRENT device; no sync, no device allocation
 *  - return 0 on success, < 0 = FA_ERR_* (argument rejected, nothing launched), > 0 = cudaError_t
 *  - thread-safe; fa_last_error() is thread-local
 *  - there is no CPU fallback and no alternative backend: unsupported arguments are an error
 */
#ifndef FA_SM100_H_
#define FA_SM100_H_

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FA_OK 0
#define FA_ERR_NULL (-1)      /* a required pointer is NULL */
#define FA_ERR_DTYPE (-2)     /* dtype not in {0, 1} */
#define FA_ERR_HEADDIM (-3)   /* D not in {64, 128} */
#define FA_ERR_SHAPE (-4)     /* non-positive dimension or a size the tensor maps cannot encode */
#define FA_ERR_ALIGN (-5)     /* a base pointer is not 16-byte aligned */
#define FA_ERR_DRIVER (-6)    /* cuTensorMapEncodeTiled unavailable or failed */
#define FA_ERR_DEVICE (-7)    /* current device is not compute capability 10.x */

#define FA_DTYPE_FP16 0
#define FA_DTYPE_BF16 1

/* Forward.  Replaces flash_attention_forward + flash_attention_forward_kernel
 * (code/My_FlashAttention_optimized.py:14-60, code/_flash_attention_kernel_optimized.py:34-129).
 *   o   [B,H,Sq,D] dtype      = softmax(q k^T * sm_scale [+ causal mask]) v
 *   lse [B,H,Sq]   fp32       = max + ln(sum exp) of the scaled, masked scores (natural log)
 * causal != 0: top-left aligned mask, row i attends to columns <= i (kernel :102).
 * sm_scale <= 0 selects the reference's hard-wired 1/sqrt(D) (launcher :56). */
int fa_sm100_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                 int B, int H, int Sq, int Sk, int D, int dtype, int causal,
                 float sm_scale, void* stream);

/* Backward.  Replaces flash_attention_backward + the dQ and dKV kernels
 * (code/My_FlashAttention_optimized.py:62-128, code/_flash_attention_kernel_optimized.py:164-386).
 *   delta [B,H,Sq] fp32 is caller-provided scratch; on return it holds rowsum(dout * o)
 *   (the reference's dQ kernel writes the same tensor, kernel :210-211, :258).
 *   dq [B,H,Sq,D], dk, dv [B,H,Sk,D] in dtype.  Deterministic (no atomics). */
int fa_sm100_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                 const float* lse, void* dq, void* dk, void* dv, float* delta,
                 int B, int H, int Sq, int Sk, int D, int dtype, int causal,
                 float sm_scale, void* stream);

/* Strided variants (what the caller's side of the operator needs): each 16-bit tensor may have
 * arbitrary batch / head / row strides as long as D is contiguous, every stride is a multiple of 8 elements
 * (16 bytes) and the base is 16-byte aligned — e.g. [B,H,S,D] views of a [B,S,H,D] buffer coming straight out of
 * a QKV projection, without the .contiguous() copy the reference makes (code/My_FlashAttention_optimized.py:138-140).
 * `strides` holds 3 element strides {batch, head, row} per tensor, in the order q,k,v,o (fwd) and
 * q,k,v,o,dout,dq,dk,dv (bwd); NULL means all contiguous.  lse and delta stay contiguous [B,H,Sq] fp32.
 * Hk = number of K/V heads (GQA / MQA): k, v, dk, dv are [B,Hk,Sk,D] and query head h uses K/V head h / (H/Hk);
 * Hk must divide H, Hk = H is the reference's layout.  dk/dv are reduced over the group inside the kernel. */
int fa_sm100_fwd_strided(const void* q, const void* k, const void* v, void* o, float* lse,
                         int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal,
                         float sm_scale, const long long* strides, void* stream);
int fa_sm100_bwd_strided(const void* q, const void* k, const void* v, const void* o, const void* dout,
                         const float* lse, void* dq, void* dk, void* dv, float* delta,
                         int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal,
                         float sm_scale, const long long* strides, void* stream, int parts);
#define FA_ERR_STRIDE (-8)    /* a stride is not a multiple of 8 elements or is negative */

/* Range-masked variants (variable-length / packed sequences, key padding, windows; the reference has no masks
 * besides causal, Phase_6.md:160-174 describes cu_seqlens packing as future work).  Query row i of batch b attends to keys
 * [row_lo[b*Sq+i], row_hi[b*Sq+i]) — intersected with keys <= i when `causal` — and the same mask seen from the key side is
 * col_lo/col_hi [B,Sk]: key row j is seen by queries [col_lo[b*Sk+j], col_hi[b*Sk+j]).  All four are device int32 arrays,
 * non-decreasing along the sequence (the kernels derive their tile ranges from the first and last row of a tile), and must
 * describe the same mask (interface.Ranges.validate() checks all of this).  A query row may see no key at all: its O is 0, its
 * LSE -inf and its dQ 0.  Tiles outside the ranges are skipped, not masked:
 * packing N sequences costs the sum of their squares.  Packed [total,H,D] tensors are passed as B = 1, Sq = Sk = total with
 * strides {0, D, H*D}.  NULL ranges = the plain operator. */
int fa_sm100_fwd_ranges(const void* q, const void* k, const void* v, void* o, float* lse,
                        int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                        const long long* strides, const int* row_lo, const int* row_hi, void* stream);
int fa_sm100_bwd_ranges(const void* q, const void* k, const void* v, const void* o, const void* dout,
                        const float* lse, void* dq, void* dk, void* dv, float* delta,
                        int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                        const long long* strides, const int* row_lo, const int* row_hi,
                        const int* col_lo, const int* col_hi, void* stream, int parts);

/* Same as fa_sm100_bwd but launches only the selected kernels: parts is a bit mask of
 * FA_BWD_DELTA (1), FA_BWD_DQ (2), FA_BWD_DKV (4).  dQ and dKV read `delta`, so it must have been
 * produced already when FA_BWD_DELTA is not set.  Used to time each kernel on its own (bench.py). */
#define FA_BWD_DELTA 1
#define FA_BWD_DQ 2
#define FA_BWD_DKV 4
int fa_sm100_bwd_parts(const void* q, const void* k, const void* v, const void* o, const void* dout,
                       const float* lse, void* dq, void* dk, void* dv, float* delta,
                       int B, int H, int Sq, int Sk, int D, int dtype, int causal,
                       float sm_scale, void* stream, int parts);

/* Everything optional in one struct (NULL = the plain operator): the range masks above and dropout — the tutorial's other
 * "next step" (Phase_6.md:54-114).  Each attention probability is kept with probability 1 - p and scaled by 1 / (1 - p); p is
 * quantised to thresh / 256 (one random byte per element).  The keep mask is a pure function of (dropout_seed, batch * H + head,
 * query row, key column) — mix32 / dropout_word in csrc/fa_ptx.cuh; the tests pin it with a numpy restatement — so the backward
 * regenerates it from the same seed; LSE is that of the undropped softmax.
 * The generator is a counter-based integer hash (two rounds of the "lowbias32" finaliser), NOT Philox: SURVEY §8f-4 names Philox
 * after the tutorial's sketch (Phase_6.md:54-114), which the reference never implemented, so there is no reference stream to
 * reproduce; the contract here is "same (seed, coordinates) -> same bit in the forward and in every backward kernel". */
typedef struct fa_sm100_options {
    const int* row_lo; const int* row_hi;       /* [B,Sq] or NULL */
    const int* col_lo; const int* col_hi;       /* [B,Sk] or NULL (backward only) */
    float dropout_p;                            /* 0 = off; must be < 1 */
    unsigned long long dropout_seed;
} fa_sm100_options;
int fa_sm100_fwd_opt(const void* q, const void* k, const void* v, void* o, float* lse,
                     int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                     const long long* strides, const fa_sm100_options* opt, void* stream);
int fa_sm100_bwd_opt(const void* q, const void* k, const void* v, const void* o, const void* dout,
                     const float* lse, void* dq, void* dk, void* dv, float* delta,
                     int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                     const long long* strides, const fa_sm100_options* opt, void* stream, int parts);

/* Fused single-pass backward (head dim 64: csrc/fa_bwd_fused.cuh, head dim 128: csrc/fa_bwd_fused128.cuh — the plain
 * operator only there, no range masks / dropout): one kernel computes dK, dV and dQ with 5 GEMMs per
 * (kv tile, q tile) pair and one exponential per score element, instead of the dQ + dK/dV kernel pair above
 * (reference launcher code/My_FlashAttention_optimized.py:111-126: 7 GEMMs, two exponentials).  dQ partials are summed over kv
 * tiles in an fp32 workspace `dq_acc` ([B,H,Sq,D] contiguous fp32, fa_sm100_bwd_fused_workspace() bytes, owned by the caller,
 * need not be initialised) by TMA reduce-add, then scaled and converted into `dq`.  Runs delta -> fused kernel -> convert.
 * dK/dV are bitwise reproducible; dQ's fp32 summation order over kv tiles depends on scheduling (use fa_sm100_bwd_strided
 * for the deterministic path).  Arguments otherwise as fa_sm100_bwd_strided.  The Python operator uses it by default at D = 64
 * (1.4x the two-kernel backward there) and on request at D = 128 (measured at parity with the two-kernel backward).
 * parts: 0 = everything, else a mask of FA_BWD_DELTA (delta + zeroing of dq_acc), FA_BWD_FUSED, FA_BWD_CONVERT (per-kernel timing). */
#define FA_BWD_FUSED 8
#define FA_BWD_CONVERT 16
size_t fa_sm100_bwd_fused_workspace(int B, int H, int Sq, int D);
int fa_sm100_bwd_fused(const void* q, const void* k, const void* v, const void* o, const void* dout,
                       const float* lse, void* dq, void* dk, void* dv, float* delta, float* dq_acc,
                       int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal,
                       float sm_scale, const long long* strides, void* stream, int parts);

/* The fused backward with options: range masks (col_lo / col_hi — it only needs the key-side view) and dropout. */
int fa_sm100_bwd_fused_opt(const void* q, const void* k, const void* v, const void* o, const void* dout,
                           const float* lse, void* dq, void* dk, void* dv, float* delta, float* dq_acc,
                           int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal,
                           float sm_scale, const long long* strides, const fa_sm100_options* opt, void* stream, int parts);

/* delta = rowsum(dout * o) alone (the preprocess step of the backward; kernel :210-211). */
int fa_sm100_delta(const void* o, const void* dout, float* delta,
                   int B, int H, int Sq, int D, int dtype, void* stream);

/* Ring / partial attention support (no reference counterpart):
 * in-place merge of a partial result (o_part [B,H,Sq,D] dtype, lse_part [B,H,Sq] fp32) computed over a
 * disjoint key set into rows [q_off, q_off+Sq) of the running fp32 accumulators
 * (o_acc [B,H,Sq_acc,D], lse_acc [B,H,Sq_acc]):
 *   lse = logaddexp(lse_acc, lse_part);  o_acc = o_acc*e^(lse_acc-lse) + o_part*e^(lse_part-lse).
 * lse = -inf partials are the identity. */
int fa_sm100_merge(float* o_acc, float* lse_acc, const void* o_part, const float* lse_part,
                   int B, int H, int Sq, int D, int dtype, int Sq_acc, int q_off, void* stream);

/* Capability query, no launch: 1 if fa_sm100_fwd/bwd accept (D, dtype, Sq, Sk), else 0. */
int fa_sm100_supported(int D, int dtype, int Sq, int Sk);

/* Thread-local message for the last non-zero return of this thread ("" if none). */
const char* fa_last_error(void);

/* Library/ABI version (major*100 + minor). */
int fa_sm100_version(void);

/* Number of this library's kernels launched by the calling process so far (bench bookkeeping). */
unsigned long long fa_sm100_launch_count(void);

/* Process-wide switch for launches that SHARE the GPU with other kernels (sequence-parallel runs: NCCL transfers overlap the
 * attention launches).  The attention kernels are persistent, one CTA per SM; by default CTA i starts on item i without touching
 * the work counter.  With on != 0 the first item is drawn from the counter like every later one, so a CTA that is scheduled late
 * (its SM was busy) takes whatever is left instead of holding one of the heaviest items until the end of the launch.
 * Returns the previous setting.  Results are identical either way. */
int fa_sm100_set_shared_sms(int on);

/* Debug: if a kernel aborted on a pipeline time-out, copies {tag, block, thread, parity} of the
 * first waiter that gave up into out[4] and returns 1; returns 0 if no time-out was recorded. */
int fa_sm100_last_hang(unsigned int out[4]);

#ifdef __cplusplus
}
#endif
#endif /* FA_SM100_H_ */
