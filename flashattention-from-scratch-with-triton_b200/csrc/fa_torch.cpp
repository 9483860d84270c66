// Native host path of the operator: FlashAttentionFunction (reference code/My_FlashAttention_optimized.py:130-166) as a C++
// torch::autograd::Function over the C ABI of libfa_sm100.so (include/fa_sm100.h).
//
// Why: one fwd+bwd step of a short problem (BASELINE config C2) is 0.2 ms on the GPU while the Python operator spends
// ~120 us per step in autograd bookkeeping, ctypes marshalling of ~20 arguments per call and five torch.empty calls; with eight
// ranks sharing the box's host cores that host path, not the GPU, set the multi-GPU step time (SCALE_r01: 0.79 efficiency with
// no collective).  This module does the same work — same saved tensors (Q, K, V, O, LSE), same outputs, same launches on the
// current stream — without re-entering the interpreter between the allocation of the outputs and the kernel launches.
//
// interface.flash_attention() routes here for the plain operator (contiguous [B,H,S,D] inputs, no range mask, no dropout);
// everything else, and the reference-shaped Python class, stays in interface.py.  PyTorch is plumbing only: tensors in, tensors
// out, the caching allocator and the current stream; all device work is the sm_100a library.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "../../include/fa_sm100.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

bool g_deterministic = false;
bool g_fused128 = false;      // head dim 128: fused single-pass backward (csrc/fa_bwd_fused128.cuh) instead of dQ + dK/dV

inline int dtype_code(const Tensor& t) { return t.scalar_type() == at::kBFloat16 ? FA_DTYPE_BF16 : FA_DTYPE_FP16; }

inline void check_rc(const char* fn, int rc) {
    TORCH_CHECK(rc == 0, fn, " failed with code ", rc, ": ", fa_last_error());
}

// reference launcher :14-60: allocate O (input dtype) and LSE (fp32), launch on the current stream
std::tuple<Tensor, Tensor> forward_impl(const Tensor& Q, const Tensor& K, const Tensor& V, bool causal, double sm_scale) {
    const auto B = Q.size(0), H = Q.size(1), Sq = Q.size(2), D = Q.size(3);
    const auto Hk = K.size(1), Sk = K.size(2);
    c10::cuda::CUDAGuard guard(Q.device());
    Tensor O = at::empty_like(Q);
    Tensor LSE = at::empty({B, H, Sq}, Q.options().dtype(at::kFloat));
    auto st = c10::cuda::getCurrentCUDAStream(Q.device().index());
    const int rc = fa_sm100_fwd_strided(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), LSE.data_ptr<float>(),
                                        (int)B, (int)H, (int)Hk, (int)Sq, (int)Sk, (int)D, dtype_code(Q), causal ? 1 : 0,
                                        (float)sm_scale, nullptr, (void*)st.stream());
    check_rc("fa_sm100_fwd_strided", rc);
    return {O, LSE};
}

// reference launcher :62-128: allocate dQ, dK, dV (+ fp32 delta), launch the backward kernels
std::tuple<Tensor, Tensor, Tensor> backward_impl(const Tensor& Q, const Tensor& K, const Tensor& V, const Tensor& O,
                                                 const Tensor& dO_in, const Tensor& LSE, bool causal, double sm_scale) {
    const auto B = Q.size(0), H = Q.size(1), Sq = Q.size(2), D = Q.size(3);
    const auto Hk = K.size(1), Sk = K.size(2);
    c10::cuda::CUDAGuard guard(Q.device());
    Tensor dO = dO_in.is_contiguous() ? dO_in : dO_in.contiguous();                 // reference :156
    Tensor dQ = at::empty_like(Q), dK = at::empty_like(K), dV = at::empty_like(V);  // reference :71-73
    Tensor delta = at::empty({B, H, Sq}, Q.options().dtype(at::kFloat));
    auto st = c10::cuda::getCurrentCUDAStream(Q.device().index());
    if ((D == 64 || (D == 128 && g_fused128)) && !g_deterministic) {
        // fused single-pass backward (delta + zeroing, fused dK/dV/dQ, dQ conversion); fp32 dQ workspace
        Tensor acc = at::empty({B, H, Sq, D}, Q.options().dtype(at::kFloat));
        const int rc = fa_sm100_bwd_fused(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                          LSE.data_ptr<float>(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(),
                                          delta.data_ptr<float>(), acc.data_ptr<float>(), (int)B, (int)H, (int)Hk, (int)Sq,
                                          (int)Sk, (int)D, dtype_code(Q), causal ? 1 : 0, (float)sm_scale, nullptr,
                                          (void*)st.stream(), 0);
        check_rc("fa_sm100_bwd_fused", rc);
    } else {
        const int rc = fa_sm100_bwd_strided(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                            LSE.data_ptr<float>(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(),
                                            delta.data_ptr<float>(), (int)B, (int)H, (int)Hk, (int)Sq, (int)Sk, (int)D,
                                            dtype_code(Q), causal ? 1 : 0, (float)sm_scale, nullptr, (void*)st.stream(),
                                            FA_BWD_DELTA | FA_BWD_DQ | FA_BWD_DKV);
        check_rc("fa_sm100_bwd_strided", rc);
    }
    return {dQ, dK, dV};
}

struct FlashAttentionFn : public torch::autograd::Function<FlashAttentionFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& Q, const Tensor& K, const Tensor& V, bool causal, double sm_scale) {
        auto [O, LSE] = forward_impl(Q, K, V, causal, sm_scale);
        ctx->save_for_backward({Q, K, V, O, LSE});                 // reference :145 (same set, same order)
        ctx->saved_data["causal"] = causal;                        // reference :147
        ctx->saved_data["sm_scale"] = sm_scale;
        return O;
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        auto [dQ, dK, dV] = backward_impl(saved[0], saved[1], saved[2], saved[3], grads[0], saved[4],
                                          ctx->saved_data["causal"].toBool(), ctx->saved_data["sm_scale"].toDouble());
        return {dQ, dK, dV, Tensor(), Tensor()};                   // reference :166 (+ undefined for the added arguments)
    }
};

// Contract of the fast path (interface.flash_attention checks the reference's asserts first): CUDA fp16/bf16 contiguous 4-D
// tensors, K/V heads dividing the query heads.  Anything else is a caller bug here, reported as an error — never a fallback.
Tensor flash_attention(const Tensor& Q, const Tensor& K, const Tensor& V, bool causal, double sm_scale) {
    TORCH_CHECK(Q.is_cuda() && K.is_cuda() && V.is_cuda(), "flash_attention: CUDA tensors required");
    TORCH_CHECK(Q.dim() == 4 && K.dim() == 4 && V.dim() == 4, "flash_attention: [B,H,S,D] tensors required");
    TORCH_CHECK(Q.scalar_type() == at::kHalf || Q.scalar_type() == at::kBFloat16, "flash_attention: fp16 / bf16 only");
    TORCH_CHECK(K.scalar_type() == Q.scalar_type() && V.scalar_type() == Q.scalar_type(), "flash_attention: dtypes differ");
    TORCH_CHECK(Q.is_contiguous() && K.is_contiguous() && V.is_contiguous(), "flash_attention fast path: contiguous tensors");
    TORCH_CHECK(K.sizes() == V.sizes() && K.size(0) == Q.size(0) && K.size(3) == Q.size(3) && Q.size(1) % K.size(1) == 0,
                "flash_attention: K/V shape does not match Q");
    return FlashAttentionFn::apply(Q, K, V, causal, sm_scale);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "C++ autograd host path of the sm_100a FlashAttention operator (over libfa_sm100.so)";
    m.def("flash_attention", &flash_attention, "O = softmax(Q K^T * scale [+causal]) V, differentiable (C++ autograd node)");
    m.def("forward", &forward_impl, "(O, LSE) = forward(Q, K, V, causal, sm_scale)");
    m.def("backward", &backward_impl, "(dQ, dK, dV) = backward(Q, K, V, O, dO, LSE, causal, sm_scale)");
    m.def("set_deterministic", [](bool f) { const bool p = g_deterministic; g_deterministic = f; return p; });
    m.def("is_deterministic", [] { return g_deterministic; });
    m.def("set_fused128", [](bool f) { const bool p = g_fused128; g_fused128 = f; return p; });
}
