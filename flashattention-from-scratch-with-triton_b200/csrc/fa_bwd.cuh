// placeholder until the backward kernels land
#pragma once
#include "fa_ptx.cuh"
namespace fa {
struct BwdParams { int BH, Sq, Sk, causal; float scale, scale_log2; const float* lse; const float* delta; int n_qtiles, n_ktiles; };
inline int launch_bwd(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                      const CUtensorMap&, const CUtensorMap&, const BwdParams&, int, int, cudaStream_t) { return (int)cudaErrorNotSupported; }
}
