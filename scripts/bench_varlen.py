"""Packed variable-length attention vs the same tokens as a padded batch: the range-masked kernels skip the tiles outside a sequence,
so N packed sequences of length L cost about the same as a [N, L] batch (and far less than one dense sequence of N*L tokens)."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return round(ts[len(ts) // 2], 4)

H = 16
for D, lens, causal in ((128, [2048] * 8, True), (64, [2048] * 8, True), (128, [4096, 1024, 512, 3000, 200, 7000, 568], True), (128, [2048] * 8, False)):
    total = sum(lens); cu = [0]
    for n in lens: cu.append(cu[-1] + n)
    q, k, v, do = (torch.randn(total, H, D, device="cuda").bfloat16() for _ in range(4))
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    cu_t = torch.tensor(cu)
    rng = fa.Ranges.from_cu_seqlens(cu, total, device="cuda")        # once per batch, shared by all layers
    def packed():
        O = fa.flash_attention_varlen(q, k, v, cu_t, causal, ranges=rng); O.backward(do); q.grad = k.grad = v.grad = None
    def packed_fwd():
        with torch.no_grad(): fa.flash_attention_varlen(q, k, v, cu_t, causal, ranges=rng)
    out = {"D": D, "lens": lens if len(set(lens)) > 1 else f"{len(lens)} x {lens[0]}", "causal": causal,
           "packed_fwd_ms": timeit(packed_fwd), "packed_fwd_bwd_ms": timeit(packed)}
    flops = sum(4.0 * H * n * n * D / (2 if causal else 1) for n in lens)
    out["packed_fwd_tflops"] = round(flops / out["packed_fwd_ms"] / 1e9, 1)
    out["packed_fwd_bwd_tflops"] = round(3.5 * flops / out["packed_fwd_bwd_ms"] / 1e9, 1)
    if len(set(lens)) == 1:
        B, L = len(lens), lens[0]
        Q, K, V, dO = (torch.randn(B, H, L, D, device="cuda").bfloat16() for _ in range(4))
        Q.requires_grad_(True); K.requires_grad_(True); V.requires_grad_(True)
        prev = fa.set_deterministic(True)                 # same two-kernel backward as the range-masked path
        def batch():
            O = fa.flash_attention(Q, K, V, causal); O.backward(dO); Q.grad = K.grad = V.grad = None
        def batch_fwd():
            with torch.no_grad(): fa.flash_attention(Q, K, V, causal)
        out["batched_fwd_ms"] = timeit(batch_fwd); out["batched_fwd_bwd_ms"] = timeit(batch)
        # the same padded batch in the [B, S, H, D] layout (row stride H*D, like the packed buffers)
        Qs, Ks, Vs, dOs = (torch.randn(B, L, H, D, device="cuda").bfloat16() for _ in range(4))
        def bshd_fwd():
            with torch.no_grad(): fa.flash_attention_bshd(Qs, Ks, Vs, causal)
        out["batched_bshd_fwd_ms"] = timeit(bshd_fwd)
        fa.set_deterministic(prev)
    print(json.dumps(out), flush=True)
