"""Host-buffer entry point: attention forward+backward for tensors that live in (pinned) host memory.

With host-resident inputs and outputs the operator is PCIe-bound (C2 moves 64 MiB in and 64 MiB out for
0.12 TFLOP of work), so the host-facing call overlaps the three legs instead of serialising them: the
(batch, head) axis — independent attention problems, the same axis the multi-GPU sharding uses — is cut
into chunks, and chunk c+1's H2D copy, chunk c's kernels and chunk c-1's D2H copy run concurrently on
three CUDA streams with double-buffered device staging.  Results are bitwise identical to one monolithic
call because kernels never mix (b,h) pairs (tests/test_gpu_parity.py::test_properties_at_full_size (4)).

The kernels are the same C-ABI launches as the operator (interface.flash_attention_forward/backward).
"""
from __future__ import annotations

import torch

from . import interface


class HostAttentionPipeline:
    """Reusable pipeline for a fixed problem shape.  Host tensors: [B,H,S,D] pinned, contiguous."""

    def __init__(self, B, H, S_q, S_k, D, dtype, is_causal, device=None, chunks=8, buffers=2, sm_scale=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.BH = B * H
        self.shape = (B, H, S_q, S_k, D)
        self.causal = bool(is_causal)
        self.sm_scale = sm_scale
        chunks = max(1, min(chunks, self.BH))
        per = -(-self.BH // chunks)
        self.ranges = [(lo, min(lo + per, self.BH)) for lo in range(0, self.BH, per)]
        self.nbuf = max(1, min(buffers, len(self.ranges)))
        mk = lambda S, dt=dtype: torch.empty(1, per, S, D, dtype=dt, device=self.device)
        self.bufs = []
        for _ in range(self.nbuf):
            self.bufs.append(dict(q=mk(S_q), k=mk(S_k), v=mk(S_k), do=mk(S_q), o=mk(S_q), dq=mk(S_q), dk=mk(S_k), dv=mk(S_k),
                                  lse=torch.empty(1, per, S_q, dtype=torch.float32, device=self.device),
                                  delta=torch.empty(1, per, S_q, dtype=torch.float32, device=self.device)))
            if D == 64:                                  # fp32 dQ workspace of the fused backward (interface.py)
                self.bufs[-1]["acc"] = torch.empty(1, per, S_q, D, dtype=torch.float32, device=self.device)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)

    def run(self, Qh, Kh, Vh, dOh, Oh, dQh, dKh, dVh):
        """Enqueue fwd+bwd of the whole problem; outputs land in the given host tensors.  Returns after all
        work is enqueued and the caller's current stream is made to wait for the last D2H copy."""
        B, H, S_q, S_k, D = self.shape
        flat = lambda t, S: t.view(1, self.BH, S, D)
        Qh, dOh, Oh, dQh = flat(Qh, S_q), flat(dOh, S_q), flat(Oh, S_q), flat(dQh, S_q)
        Kh, Vh, dKh, dVh = flat(Kh, S_k), flat(Vh, S_k), flat(dKh, S_k), flat(dVh, S_k)
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event(); start.record(cur)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)
        in_done = [None] * len(self.ranges); cmp_done = [None] * len(self.ranges); out_done = [None] * len(self.ranges)
        for c, (lo, hi) in enumerate(self.ranges):
            b = self.bufs[c % self.nbuf]; n = hi - lo
            with torch.cuda.stream(self.s_in):
                if c >= self.nbuf:                       # staging inputs free once chunk c-nbuf's kernels are done
                    self.s_in.wait_event(cmp_done[c - self.nbuf])
                for name, src in (("q", Qh), ("k", Kh), ("v", Vh), ("do", dOh)):
                    b[name][:, :n].copy_(src[:, lo:hi], non_blocking=True)
                in_done[c] = torch.cuda.Event(); in_done[c].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(in_done[c])
                if c >= self.nbuf:                       # staging outputs free once chunk c-nbuf's D2H is done
                    self.s_cmp.wait_event(out_done[c - self.nbuf])
                q, k, v, do = b["q"][:, :n], b["k"][:, :n], b["v"][:, :n], b["do"][:, :n]
                lib = interface._cabi.load()
                dt = interface._DT[q.dtype]
                scale = float(self.sm_scale) if self.sm_scale is not None else 0.0
                st = self.s_cmp.cuda_stream
                with torch.cuda.device(self.device):
                    rc = lib.fa_sm100_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), b["o"].data_ptr(), b["lse"].data_ptr(),
                                          1, n, S_q, S_k, D, dt, int(self.causal), scale, st)
                    interface._cabi.check("fa_sm100_fwd", rc)
                    if "acc" in b and not interface.is_deterministic():
                        rc = lib.fa_sm100_bwd_fused(q.data_ptr(), k.data_ptr(), v.data_ptr(), b["o"].data_ptr(), do.data_ptr(),
                                                    b["lse"].data_ptr(), b["dq"].data_ptr(), b["dk"].data_ptr(), b["dv"].data_ptr(),
                                                    b["delta"].data_ptr(), b["acc"].data_ptr(), 1, n, n, S_q, S_k, D, dt,
                                                    int(self.causal), scale, None, st, 0)
                        interface._cabi.check("fa_sm100_bwd_fused", rc)
                    else:
                        rc = lib.fa_sm100_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), b["o"].data_ptr(), do.data_ptr(),
                                              b["lse"].data_ptr(), b["dq"].data_ptr(), b["dk"].data_ptr(), b["dv"].data_ptr(),
                                              b["delta"].data_ptr(), 1, n, S_q, S_k, D, dt, int(self.causal), scale, st)
                        interface._cabi.check("fa_sm100_bwd", rc)
                cmp_done[c] = torch.cuda.Event(); cmp_done[c].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(cmp_done[c])
                for name, dst in (("o", Oh), ("dq", dQh), ("dk", dKh), ("dv", dVh)):
                    dst[:, lo:hi].copy_(b[name][:, :n], non_blocking=True)
                out_done[c] = torch.cuda.Event(); out_done[c].record(self.s_out)
        cur.wait_event(out_done[-1])
        cur.wait_event(cmp_done[-1])


def flash_attention_host(Qh, Kh, Vh, dOh, is_causal=False, chunks=8, device=None, sm_scale=None):
    """One-shot convenience wrapper: pinned host tensors in, pinned host (O, dQ, dK, dV) out (synchronises)."""
    B, H, S_q, D = Qh.shape
    S_k = Kh.shape[2]
    pipe = HostAttentionPipeline(B, H, S_q, S_k, D, Qh.dtype, is_causal, device, chunks, sm_scale=sm_scale)
    outs = [torch.empty_like(t).pin_memory() for t in (Qh, Qh, Kh, Vh)]
    pipe.run(Qh, Kh, Vh, dOh, *outs)
    torch.cuda.synchronize(pipe.device)
    return tuple(outs)
