// C-ABI entry points of libfa_sm100.so (declared in include/fa_sm100.h): argument checks, host-side
// CUtensorMap construction (replaces the reference's TensorDescriptor objects and autotune
// pre-hooks, code/My_FlashAttention_optimized.py:33-51 and kernel file :7-16), launches.
#include "fa_fwd.cuh"
#include "fa_bwd.cuh"
#include "fa_bwd_fused.cuh"
#include "fa_bwd_fused128.cuh"
#include "fa_aux.cuh"
#include "../../include/fa_sm100.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <mutex>

using namespace fa;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_shared_sms{0};       // fa_sm100_set_shared_sms: persistent CTAs draw their first item from the counter too

int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// [B, H, S, D] 16-bit tensor as a 4-D map with explicit batch / head / row strides (never flattened across heads,
// so a partial tile cannot touch the next head — fixes the reference's flattened-descriptor overrun).
// Box = [1][1][box_rows][64 columns], SWIZZLE_128B; out-of-range rows read as zero and are not written.
thread_local int t_last_curesult = 0;       // CUresult of the last failed cuTensorMapEncodeTiled of this thread (error messages)

bool make_map(CUtensorMap* m, const void* ptr, int B, int H, int S, int D, const RowStrides& st, int dtype, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { t_last_curesult = -1; return false; }
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)st.r * 2, (cuuint64_t)st.h * 2, (cuuint64_t)st.b * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, dtype == FA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) t_last_curesult = (int)r;
    return r == CUDA_SUCCESS;
}

// fp32 accumulator [BH, S, D] (contiguous) as a 3-D map, box = [1][128 rows][32 columns = 128 B], SWIZZLE_128B: target of the
// fused backward's cp.reduce.async.bulk.tensor (.add); rows past S are dropped
bool make_map_acc(CUtensorMap* m, const float* ptr, int BH, int S, int D) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)S * D * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// strides of tensor `i` from the caller's array (NULL = contiguous [B,H,S,D]); size-1 dims get a harmless stride
RowStrides get_strides(const long long* strides, int i, int H, int S, int D) {
    RowStrides st{(long long)H * S * D, (long long)S * D, (long long)D};
    if (strides) { st.b = strides[3 * i]; st.h = strides[3 * i + 1]; st.r = strides[3 * i + 2]; }
    return st;
}
bool strides_ok(RowStrides& st, int B, int H, int S, int D) {
    if (S == 1) st.r = D;                       // TMA wants non-zero 16-byte-multiple strides even for extent-1 dims
    if (H == 1) st.h = (long long)S * st.r;
    if (B == 1) st.b = (long long)H * st.h;
    return st.b > 0 && st.h > 0 && st.r >= D && (st.b % 8 == 0) && (st.h % 8 == 0) && (st.r % 8 == 0);
}

// `ready` publishes the other fields (release / acquire): concurrent first calls on one device are safe
struct DeviceInfo { int ordinal = 0; int sms = 0; int cc_major = 0; unsigned int* sched_ring = nullptr; std::atomic<int> ready{0}; };
constexpr int kMaxDevices = 64;
// Work-counter ring of the persistent kernels (one instance per device: a __device__ symbol).  A launch takes one UNIT = 4 words =
// two self-resetting {work counter, retired-CTA counter} pairs (sched_retire), handed out round-robin with a single power-of-two
// modulus, so two allocations never overlap partially.  A unit is reused after kSchedUnits later launches of the process: the
// library supports up to kSchedUnits (4096) launches pending at once across all streams and captured graphs of a device.
constexpr unsigned int kSchedUnits = 4096;
__device__ unsigned int g_sched_ring[kSchedUnits * 4];
DeviceInfo g_dev[kMaxDevices];
std::mutex g_dev_mu;
std::atomic<unsigned int> g_sched_next{0};
unsigned int sched_slot() { return (g_sched_next.fetch_add(1u, std::memory_order_relaxed) & (kSchedUnits - 1u)) * 4u; }

int device_info(DeviceInfo** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) return fail(FA_ERR_DEVICE, "device ordinal %d out of range", dev);
    // cuTensorMapEncodeTiled is a DRIVER call and needs a current context.  A thread that has only ever seen cached allocations
    // (PyTorch's autograd worker running the first backward of a process) has none yet: cudaGetDevice does not bind the primary
    // context, and the encode fails with CUDA_ERROR_INVALID_CONTEXT (201).  Bind it once per thread and device.
    static thread_local int t_ctx_dev = -1;
    if (t_ctx_dev != dev) {
        e = cudaSetDevice(dev);
        if (e == cudaSuccess) e = cudaFree(nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "binding the primary context of the device to this thread");
        t_ctx_dev = dev;
    }
    DeviceInfo& d = g_dev[dev];
    if (!d.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        if (!d.ready.load(std::memory_order_relaxed)) {
            int sms = 0, major = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
            void* ring = nullptr;
            e = cudaGetSymbolAddress(&ring, g_sched_ring);
            if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(sched ring) — was the library built for sm_100a?");
            d.ordinal = dev; d.cc_major = major; d.sched_ring = (unsigned int*)ring; d.sms = sms;
            d.ready.store(1, std::memory_order_release);
        }
    }
    if (d.cc_major != 10) return fail(FA_ERR_DEVICE, "libfa_sm100 needs compute capability 10.x, device has %d.x", d.cc_major);
    *out = &d;
    return 0;
}

int check_common(int B, int H, int Sq, int Sk, int D, int dtype) {
    if (dtype != FA_DTYPE_FP16 && dtype != FA_DTYPE_BF16) return fail(FA_ERR_DTYPE, "dtype %d not in {0=fp16, 1=bf16}", dtype);
    if (D != 64 && D != 128) return fail(FA_ERR_HEADDIM, "head dim %d not in {64, 128}", D);
    if (B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) return fail(FA_ERR_SHAPE, "non-positive dimension B=%d H=%d Sq=%d Sk=%d", B, H, Sq, Sk);
    if ((long long)B * H > 2147483647LL / 4) return fail(FA_ERR_SHAPE, "B*H too large");
    return 0;
}
// heads per scheduling chunk: as many heads as keep their tensors (bytes_per_head each) inside ~48 MB of the 126 MB L2,
// balanced over the chunks
int heads_per_chunk(int n_heads, double bytes_per_head) {
    long long hc = (long long)(48.0 * 1024 * 1024 / bytes_per_head);
    if (hc < 1) hc = 1;
    if (hc > n_heads) hc = n_heads;
    const int n_chunks = (int)((n_heads + hc - 1) / hc);
    return (n_heads + n_chunks - 1) / n_chunks;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// dropout_p is quantised to 1/256 (one random byte per element): thresh = round(256 p), kept elements scale by 256 / (256 - thresh)
int make_dropout(const fa_sm100_options* opt, DropoutParams* d) {
    d->seed0 = d->seed1 = d->thresh = 0; d->scale = 1.f;
    if (!opt || opt->dropout_p == 0.f) return 0;
    if (!(opt->dropout_p > 0.f && opt->dropout_p < 1.f)) return fail(FA_ERR_SHAPE, "dropout_p %g not in [0, 1)", (double)opt->dropout_p);
    unsigned int t = (unsigned int)lrintf(opt->dropout_p * 256.f);
    if (t > 255u) t = 255u;
    d->thresh = t; d->scale = 256.f / (256.f - (float)t);
    d->seed0 = (uint32_t)(opt->dropout_seed & 0xffffffffull); d->seed1 = (uint32_t)(opt->dropout_seed >> 32);
    return 0;
}

template <int D, bool kBf16, bool kRanges, bool kDropout>
int launch_fwd_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
                 const FwdParams& p, int grid, int dev, cudaStream_t st) {
    cudaError_t e = ensure_smem<fa_fwd_kernel<D, kBf16, kRanges, kDropout>>(FwdCfg<D>::kSmemBytes, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fwd smem)");
    e = launch_pdl(fa_fwd_kernel<D, kBf16, kRanges, kDropout>, grid, kFwdThreads, FwdCfg<D>::kSmemBytes, st, mq, mk, mv, mo, p);
    ++g_launches;
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "fa_fwd_kernel launch");
}
template <int D, bool kBf16>
int launch_fwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo,
               const FwdParams& p, int grid, int dev, cudaStream_t st) {
    if (p.drop.thresh) return p.row_lo ? launch_fwd_t<D, kBf16, true, true>(mq, mk, mv, mo, p, grid, dev, st)
                                       : launch_fwd_t<D, kBf16, false, true>(mq, mk, mv, mo, p, grid, dev, st);
    return p.row_lo ? launch_fwd_t<D, kBf16, true, false>(mq, mk, mv, mo, p, grid, dev, st)
                    : launch_fwd_t<D, kBf16, false, false>(mq, mk, mv, mo, p, grid, dev, st);
}

}  // namespace

extern "C" {

int fa_sm100_version(void) { return 100; }
const char* fa_last_error(void) { return g_err; }
unsigned long long fa_sm100_launch_count(void) { return g_launches.load(); }
int fa_sm100_set_shared_sms(int on) { return g_shared_sms.exchange(on ? 1 : 0); }

int fa_sm100_supported(int D, int dtype, int Sq, int Sk) {
    return (D == 64 || D == 128) && (dtype == 0 || dtype == 1) && Sq > 0 && Sk > 0;
}

int fa_sm100_last_hang(unsigned int out[4]) {
    unsigned int flag = 0;
    if (cudaMemcpyFromSymbol(&flag, g_hang_flag, sizeof(flag)) != cudaSuccess || !flag) return 0;
    HangRecord r;
    if (cudaMemcpyFromSymbol(&r, g_hang_record, sizeof(r)) != cudaSuccess) return 0;
    out[0] = r.tag; out[1] = r.block; out[2] = r.thread; out[3] = r.parity;
    flag = 0;
    cudaMemcpyToSymbol(g_hang_flag, &flag, sizeof(flag));     // re-arm
    return 1;
}

int fa_sm100_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                 int B, int H, int Sq, int Sk, int D, int dtype, int causal, float sm_scale, void* stream) {
    return fa_sm100_fwd_strided(q, k, v, o, lse, B, H, H, Sq, Sk, D, dtype, causal, sm_scale, nullptr, stream);
}

int fa_sm100_fwd_strided(const void* q, const void* k, const void* v, void* o, float* lse,
                         int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                         const long long* strides, void* stream) {
    return fa_sm100_fwd_ranges(q, k, v, o, lse, B, H, Hk, Sq, Sk, D, dtype, causal, sm_scale, strides, nullptr, nullptr, stream);
}

int fa_sm100_fwd_ranges(const void* q, const void* k, const void* v, void* o, float* lse,
                        int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                        const long long* strides, const int* row_lo, const int* row_hi, void* stream) {
    fa_sm100_options opt = {row_lo, row_hi, nullptr, nullptr, 0.f, 0ull};
    return fa_sm100_fwd_opt(q, k, v, o, lse, B, H, Hk, Sq, Sk, D, dtype, causal, sm_scale, strides, &opt, stream);
}

int fa_sm100_fwd_opt(const void* q, const void* k, const void* v, void* o, float* lse,
                     int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                     const long long* strides, const fa_sm100_options* opt, void* stream) {
    const int* row_lo = opt ? opt->row_lo : nullptr; const int* row_hi = opt ? opt->row_hi : nullptr;
    DropoutParams drop; if (int rc = make_dropout(opt, &drop)) return rc;
    if (!q || !k || !v || !o || !lse) return fail(FA_ERR_NULL, "null tensor pointer");
    if ((row_lo == nullptr) != (row_hi == nullptr)) return fail(FA_ERR_NULL, "row_lo and row_hi must be given together");
    if (int rc = check_common(B, H, Sq, Sk, D, dtype)) return rc;
    if (Hk <= 0 || H % Hk) return fail(FA_ERR_SHAPE, "K/V heads %d must divide query heads %d", Hk, H);
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o)) return fail(FA_ERR_ALIGN, "q/k/v/o must be 16-byte aligned");
    RowStrides sq = get_strides(strides, 0, H, Sq, D), sk = get_strides(strides, 1, Hk, Sk, D),
               sv = get_strides(strides, 2, Hk, Sk, D), so = get_strides(strides, 3, H, Sq, D);
    if (!strides_ok(sq, B, H, Sq, D) || !strides_ok(sk, B, Hk, Sk, D) || !strides_ok(sv, B, Hk, Sk, D) || !strides_ok(so, B, H, Sq, D))
        return fail(FA_ERR_STRIDE, "strides must be positive multiples of 8 elements with D contiguous");
    DeviceInfo* dev; if (int rc = device_info(&dev)) return rc;
    const int BH = B * H;
    CUtensorMap mq, mk, mv, mo;
    if (!make_map(&mq, q, B, H, Sq, D, sq, dtype, 128) || !make_map(&mk, k, B, Hk, Sk, D, sk, dtype, 128) ||
        !make_map(&mv, v, B, Hk, Sk, D, sv, dtype, 128) || !make_map(&mo, o, B, H, Sq, D, so, dtype, 128))
        return fail(FA_ERR_DRIVER, "cuTensorMapEncodeTiled failed (B=%d H=%d Sq=%d Sk=%d D=%d)", B, H, Sq, Sk, D);
    FwdParams p;
    p.BH = BH; p.H = H; p.G = H / Hk; p.Sq = Sq; p.Sk = Sk;
    p.n_qblk = (Sq + 255) / 256;
    p.n_items = BH * p.n_qblk;
    p.hc = heads_per_chunk(BH, 4.0 * ((double)Sq + Sk) * D);
    p.causal = causal ? 1 : 0;
    p.scale = sm_scale > 0.f ? sm_scale : 1.0f / sqrtf((float)D);
    p.scale_log2 = p.scale * 1.44269504088896340736f;
    p.lse = lse;
    p.row_lo = row_lo; p.row_hi = row_hi; p.drop = drop;
    p.sched = dev->sched_ring + sched_slot();           // self-resetting counter pair (sched_retire): no memset on the stream
    p.dyn_first = g_shared_sms.load(std::memory_order_relaxed);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = p.n_items < dev->sms ? p.n_items : dev->sms;
    if (D == 64) return dtype ? launch_fwd<64, true>(mq, mk, mv, mo, p, grid, dev->ordinal, st) : launch_fwd<64, false>(mq, mk, mv, mo, p, grid, dev->ordinal, st);
    return dtype ? launch_fwd<128, true>(mq, mk, mv, mo, p, grid, dev->ordinal, st) : launch_fwd<128, false>(mq, mk, mv, mo, p, grid, dev->ordinal, st);
}

int fa_sm100_delta(const void* o, const void* dout, float* delta, int B, int H, int Sq, int D, int dtype, void* stream) {
    if (!o || !dout || !delta) return fail(FA_ERR_NULL, "null tensor pointer");
    if (int rc = check_common(B, H, Sq, 1, D, dtype)) return rc;
    if (!aligned16(o) || !aligned16(dout)) return fail(FA_ERR_ALIGN, "o/dout must be 16-byte aligned");
    DeviceInfo* dev; if (int rc = device_info(&dev)) return rc;
    RowStrides sc = get_strides(nullptr, 0, H, Sq, D);
    int rc = launch_delta(o, dout, delta, (long long)B * H * Sq, H, Sq, sc, sc, D, dtype, dev->sms, (cudaStream_t)stream);
    ++g_launches;
    return rc == 0 ? 0 : cuda_fail((cudaError_t)rc, "fa_delta_kernel launch");
}

int fa_sm100_merge(float* o_acc, float* lse_acc, const void* o_part, const float* lse_part,
                   int B, int H, int Sq, int D, int dtype, int Sq_acc, int q_off, void* stream) {
    if (!o_acc || !lse_acc || !o_part || !lse_part) return fail(FA_ERR_NULL, "null tensor pointer");
    if (int rc = check_common(B, H, Sq, 1, D, dtype)) return rc;
    if (q_off < 0 || Sq_acc < q_off + Sq) return fail(FA_ERR_SHAPE, "rows [%d, %d) do not fit an accumulator of %d rows", q_off, q_off + Sq, Sq_acc);
    if (!aligned16(o_acc) || !aligned16(o_part)) return fail(FA_ERR_ALIGN, "o_acc/o_part must be 16-byte aligned");
    DeviceInfo* dev; if (int rc = device_info(&dev)) return rc;
    int rc = launch_merge(o_acc, lse_acc, o_part, lse_part, (long long)B * H * Sq, Sq, Sq_acc, q_off, D, dtype, dev->sms, (cudaStream_t)stream);
    ++g_launches;
    return rc == 0 ? 0 : cuda_fail((cudaError_t)rc, "fa_merge_kernel launch");
}

int fa_sm100_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                 const float* lse, void* dq, void* dk, void* dv, float* delta,
                 int B, int H, int Sq, int Sk, int D, int dtype, int causal, float sm_scale, void* stream) {
    return fa_sm100_bwd_parts(q, k, v, o, dout, lse, dq, dk, dv, delta, B, H, Sq, Sk, D, dtype, causal, sm_scale,
                              stream, FA_BWD_DELTA | FA_BWD_DQ | FA_BWD_DKV);
}

int fa_sm100_bwd_parts(const void* q, const void* k, const void* v, const void* o, const void* dout,
                       const float* lse, void* dq, void* dk, void* dv, float* delta,
                       int B, int H, int Sq, int Sk, int D, int dtype, int causal, float sm_scale, void* stream,
                       int parts) {
    return fa_sm100_bwd_strided(q, k, v, o, dout, lse, dq, dk, dv, delta, B, H, H, Sq, Sk, D, dtype, causal, sm_scale,
                                nullptr, stream, parts);
}

int fa_sm100_bwd_strided(const void* q, const void* k, const void* v, const void* o, const void* dout,
                         const float* lse, void* dq, void* dk, void* dv, float* delta,
                         int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                         const long long* strides, void* stream, int parts) {
    return fa_sm100_bwd_ranges(q, k, v, o, dout, lse, dq, dk, dv, delta, B, H, Hk, Sq, Sk, D, dtype, causal, sm_scale, strides,
                               nullptr, nullptr, nullptr, nullptr, stream, parts);
}

int fa_sm100_bwd_ranges(const void* q, const void* k, const void* v, const void* o, const void* dout,
                        const float* lse, void* dq, void* dk, void* dv, float* delta,
                        int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                        const long long* strides, const int* row_lo, const int* row_hi, const int* col_lo, const int* col_hi,
                        void* stream, int parts) {
    fa_sm100_options opt = {row_lo, row_hi, col_lo, col_hi, 0.f, 0ull};
    return fa_sm100_bwd_opt(q, k, v, o, dout, lse, dq, dk, dv, delta, B, H, Hk, Sq, Sk, D, dtype, causal, sm_scale, strides, &opt,
                            stream, parts);
}

int fa_sm100_bwd_opt(const void* q, const void* k, const void* v, const void* o, const void* dout,
                     const float* lse, void* dq, void* dk, void* dv, float* delta,
                     int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                     const long long* strides, const fa_sm100_options* opt, void* stream, int parts) {
    const int* row_lo = opt ? opt->row_lo : nullptr; const int* row_hi = opt ? opt->row_hi : nullptr;
    const int* col_lo = opt ? opt->col_lo : nullptr; const int* col_hi = opt ? opt->col_hi : nullptr;
    DropoutParams drop; if (int rc = make_dropout(opt, &drop)) return rc;
    if (!q || !k || !v || !o || !dout || !lse || !dq || !dk || !dv || !delta) return fail(FA_ERR_NULL, "null tensor pointer");
    if (!row_lo != !row_hi || !row_lo != !col_lo || !row_lo != !col_hi)
        return fail(FA_ERR_NULL, "row_lo, row_hi, col_lo and col_hi must be given together");
    if (int rc = check_common(B, H, Sq, Sk, D, dtype)) return rc;
    if (Hk <= 0 || H % Hk) return fail(FA_ERR_SHAPE, "K/V heads %d must divide query heads %d", Hk, H);
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o) || !aligned16(dout) ||
        !aligned16(dq) || !aligned16(dk) || !aligned16(dv) || !aligned16(lse) || !aligned16(delta))
        return fail(FA_ERR_ALIGN, "all tensors must be 16-byte aligned");
    RowStrides s_q = get_strides(strides, 0, H, Sq, D), s_k = get_strides(strides, 1, Hk, Sk, D),
               s_v = get_strides(strides, 2, Hk, Sk, D), s_o = get_strides(strides, 3, H, Sq, D),
               s_do = get_strides(strides, 4, H, Sq, D), s_dq = get_strides(strides, 5, H, Sq, D),
               s_dk = get_strides(strides, 6, Hk, Sk, D), s_dv = get_strides(strides, 7, Hk, Sk, D);
    if (!strides_ok(s_q, B, H, Sq, D) || !strides_ok(s_k, B, Hk, Sk, D) || !strides_ok(s_v, B, Hk, Sk, D) ||
        !strides_ok(s_o, B, H, Sq, D) || !strides_ok(s_do, B, H, Sq, D) || !strides_ok(s_dq, B, H, Sq, D) ||
        !strides_ok(s_dk, B, Hk, Sk, D) || !strides_ok(s_dv, B, Hk, Sk, D))
        return fail(FA_ERR_STRIDE, "strides must be positive multiples of 8 elements with D contiguous");
    DeviceInfo* dev; if (int rc = device_info(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int BH = B * H;
    int rc = 0;
    if (parts & FA_BWD_DELTA) {
        rc = launch_delta(o, dout, delta, (long long)BH * Sq, H, Sq, s_o, s_do, D, dtype, dev->sms, st);
        ++g_launches;
        if (rc) return cuda_fail((cudaError_t)rc, "fa_delta_kernel launch");
    }
    if (!(parts & (FA_BWD_DQ | FA_BWD_DKV))) return 0;
    CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
    if (!make_map(&mq, q, B, H, Sq, D, s_q, dtype, 128) || !make_map(&mk, k, B, Hk, Sk, D, s_k, dtype, 128) ||
        !make_map(&mv, v, B, Hk, Sk, D, s_v, dtype, 128) || !make_map(&mdo, dout, B, H, Sq, D, s_do, dtype, 128) ||
        !make_map(&mdq, dq, B, H, Sq, D, s_dq, dtype, 128) || !make_map(&mdk, dk, B, Hk, Sk, D, s_dk, dtype, 128) ||
        !make_map(&mdv, dv, B, Hk, Sk, D, s_dv, dtype, 128))
        return fail(FA_ERR_DRIVER, "cuTensorMapEncodeTiled failed (B=%d H=%d Sq=%d Sk=%d D=%d)", B, H, Sq, Sk, D);
    BwdParams p;
    p.BH = BH; p.H = H; p.Hk = Hk; p.G = H / Hk; p.Sq = Sq; p.Sk = Sk; p.causal = causal ? 1 : 0;
    p.scale = sm_scale > 0.f ? sm_scale : 1.0f / sqrtf((float)D);
    p.scale_log2 = p.scale * 1.44269504088896340736f;
    p.lse = lse; p.delta = delta;
    p.row_lo = row_lo; p.row_hi = row_hi; p.col_lo = col_lo; p.col_hi = col_hi; p.drop = drop;
    p.n_qtiles = (Sq + 127) / 128; p.n_ktiles = (Sk + 127) / 128;
    p.sms = dev->sms;
    p.hc_dq = heads_per_chunk(BH, 4.0 * ((double)Sq + Sk) * D);
    p.hc_dkv = heads_per_chunk(BH / p.G, 4.0 * ((double)p.G * Sq + Sk) * D);
    {   // two self-resetting counter pairs from the ring
        const unsigned int slot = sched_slot();
        p.sched_dkv = dev->sched_ring + slot; p.sched_dq = dev->sched_ring + slot + 2;
        p.dyn_first = g_shared_sms.load(std::memory_order_relaxed);
    }
    p.dev = dev->ordinal;
    rc = launch_bwd(mq, mk, mv, mdo, mdq, mdk, mdv, p, D, dtype, parts, st);
    g_launches += ((parts & FA_BWD_DQ) ? 1 : 0) + ((parts & FA_BWD_DKV) ? 1 : 0);
    return rc == 0 ? 0 : cuda_fail((cudaError_t)rc, "fa_bwd kernels launch");
}

size_t fa_sm100_bwd_fused_workspace(int B, int H, int Sq, int D) {
    if ((D != 64 && D != 128) || B <= 0 || H <= 0 || Sq <= 0) return 0;
    return (size_t)B * H * Sq * D * sizeof(float);
}

int fa_sm100_bwd_fused(const void* q, const void* k, const void* v, const void* o, const void* dout,
                       const float* lse, void* dq, void* dk, void* dv, float* delta, float* dq_acc,
                       int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                       const long long* strides, void* stream, int parts) {
    return fa_sm100_bwd_fused_opt(q, k, v, o, dout, lse, dq, dk, dv, delta, dq_acc, B, H, Hk, Sq, Sk, D, dtype, causal, sm_scale,
                                  strides, nullptr, stream, parts);
}

int fa_sm100_bwd_fused_opt(const void* q, const void* k, const void* v, const void* o, const void* dout,
                           const float* lse, void* dq, void* dk, void* dv, float* delta, float* dq_acc,
                           int B, int H, int Hk, int Sq, int Sk, int D, int dtype, int causal, float sm_scale,
                           const long long* strides, const fa_sm100_options* opt, void* stream, int parts) {
    DropoutParams drop; if (int rc = make_dropout(opt, &drop)) return rc;
    const int* col_lo = opt ? opt->col_lo : nullptr; const int* col_hi = opt ? opt->col_hi : nullptr;
    if (!col_lo != !col_hi) return fail(FA_ERR_NULL, "col_lo and col_hi must be given together");
    if (parts == 0) parts = FA_BWD_DELTA | FA_BWD_FUSED | FA_BWD_CONVERT;
    if (!q || !k || !v || !o || !dout || !lse || !dq || !dk || !dv || !delta || !dq_acc) return fail(FA_ERR_NULL, "null tensor pointer");
    if (int rc = check_common(B, H, Sq, Sk, D, dtype)) return rc;
    if (D == 128 && (col_lo || drop.thresh))
        return fail(FA_ERR_HEADDIM, "the fused backward at head dim 128 has no range-mask / dropout instantiation: use fa_sm100_bwd_opt");
    if (Hk <= 0 || H % Hk) return fail(FA_ERR_SHAPE, "K/V heads %d must divide query heads %d", Hk, H);
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o) || !aligned16(dout) || !aligned16(dq) ||
        !aligned16(dk) || !aligned16(dv) || !aligned16(lse) || !aligned16(delta) || !aligned16(dq_acc))
        return fail(FA_ERR_ALIGN, "all tensors must be 16-byte aligned");
    RowStrides s_q = get_strides(strides, 0, H, Sq, D), s_k = get_strides(strides, 1, Hk, Sk, D),
               s_v = get_strides(strides, 2, Hk, Sk, D), s_o = get_strides(strides, 3, H, Sq, D),
               s_do = get_strides(strides, 4, H, Sq, D), s_dq = get_strides(strides, 5, H, Sq, D),
               s_dk = get_strides(strides, 6, Hk, Sk, D), s_dv = get_strides(strides, 7, Hk, Sk, D);
    if (!strides_ok(s_q, B, H, Sq, D) || !strides_ok(s_k, B, Hk, Sk, D) || !strides_ok(s_v, B, Hk, Sk, D) ||
        !strides_ok(s_o, B, H, Sq, D) || !strides_ok(s_do, B, H, Sq, D) || !strides_ok(s_dq, B, H, Sq, D) ||
        !strides_ok(s_dk, B, Hk, Sk, D) || !strides_ok(s_dv, B, Hk, Sk, D))
        return fail(FA_ERR_STRIDE, "strides must be positive multiples of 8 elements with D contiguous");
    DeviceInfo* dev; if (int rc = device_info(&dev)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int BH = B * H;
    CUtensorMap mq, mk, mv, mdo, mdk, mdv, macc;
    {
        const char* bad = !make_map(&mq, q, B, H, Sq, D, s_q, dtype, 128) ? "q" : !make_map(&mk, k, B, Hk, Sk, D, s_k, dtype, 128) ? "k" :
                          !make_map(&mv, v, B, Hk, Sk, D, s_v, dtype, 128) ? "v" : !make_map(&mdo, dout, B, H, Sq, D, s_do, dtype, 128) ? "dout" :
                          !make_map(&mdk, dk, B, Hk, Sk, D, s_dk, dtype, 128) ? "dk" : !make_map(&mdv, dv, B, Hk, Sk, D, s_dv, dtype, 128) ? "dv" :
                          !make_map_acc(&macc, dq_acc, BH, Sq, D) ? "dq_acc" : nullptr;
        if (bad) return fail(FA_ERR_DRIVER, "cuTensorMapEncodeTiled failed for %s with CUresult %d (B=%d H=%d Hk=%d Sq=%d Sk=%d D=%d, q=%p strides b/h/r %lld/%lld/%lld)",
                             bad, t_last_curesult, B, H, Hk, Sq, Sk, D, q, s_q.b, s_q.h, s_q.r);
    }
    // delta = rowsum(dO o O) and, in the same pass, zeros into the dQ accumulator
    int rc = 0;
    if (parts & FA_BWD_DELTA) {
        rc = launch_delta(o, dout, delta, (long long)BH * Sq, H, Sq, s_o, s_do, D, dtype, dev->sms, st, dq_acc);
        ++g_launches;
        if (rc) return cuda_fail((cudaError_t)rc, "fa_delta_kernel launch");
    }
    BwdParams p;
    p.BH = BH; p.H = H; p.Hk = Hk; p.G = H / Hk; p.Sq = Sq; p.Sk = Sk; p.causal = causal ? 1 : 0;
    p.scale = sm_scale > 0.f ? sm_scale : 1.0f / sqrtf((float)D);
    p.scale_log2 = p.scale * 1.44269504088896340736f;
    p.lse = lse; p.delta = delta;
    p.row_lo = p.row_hi = nullptr; p.col_lo = col_lo; p.col_hi = col_hi;     // the fused kernel sees the mask from the key side only
    p.drop = drop;
    p.n_qtiles = (Sq + 127) / 128; p.n_ktiles = (Sk + 127) / 128;
    p.sms = dev->sms;
    p.hc_dq = heads_per_chunk(BH, 4.0 * ((double)Sq + Sk) * D);
    p.hc_dkv = heads_per_chunk(BH / p.G, 4.0 * ((double)p.G * Sq + Sk) * D);
    p.sched_dkv = dev->sched_ring + sched_slot(); p.sched_dq = nullptr;
    p.dyn_first = g_shared_sms.load(std::memory_order_relaxed);
    p.dev = dev->ordinal;
    if (D == 128)
        rc = dtype ? launch_bwd_fused128_t<true>(mq, mk, mv, mdo, mdk, mdv, macc, p, dq_acc, dq, s_dq, st, parts)
                   : launch_bwd_fused128_t<false>(mq, mk, mv, mdo, mdk, mdv, macc, p, dq_acc, dq, s_dq, st, parts);
    else
        rc = dtype ? launch_bwd_fused_t<true>(mq, mk, mv, mdo, mdk, mdv, macc, p, dq_acc, dq, s_dq, st, parts)
                   : launch_bwd_fused_t<false>(mq, mk, mv, mdo, mdk, mdv, macc, p, dq_acc, dq, s_dq, st, parts);
    g_launches += ((parts & FA_BWD_FUSED) ? 1 : 0) + ((parts & FA_BWD_CONVERT) ? 1 : 0);
    return rc == 0 ? 0 : cuda_fail((cudaError_t)rc, "fa_bwd_fused kernels launch");
}

}  // extern "C"
