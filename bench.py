#!/usr/bin/env python
"""Benchmark of the hot path: flash_attention forward + backward on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4s|C4|sweep4k] [--impl ours|reference]

A "step" is one fwd+bwd pass of the operator over one batch of synthetic [B,H,S,D] inputs (the
reference's own benchmark step, code/Performance_Comparison.py:66-76).  Metric and FLOP model are the
reference's: TFLOPS = 3.5 * 4*B*H*Sq*Sk*D/(2 if causal) / t  (code/Performance_Comparison.py:99-107).

N=1 workload: BASELINE.json configs[1] (C2: B=4 H=16 N=2048 D=64 bf16 causal fwd+bwd).  N>1: batch x head
sharding — every rank runs the same per-GPU shard shape on its own synthetic batch (global batch = N*B),
no data-path collective => "scaling": "weak"; `value` is the aggregate over all ranks and the
time is the max over ranks.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (the backward kernel) against the MEASURED dense-bf16 peak
  kernels       per-kernel ms / algorithmic TFLOP/s / fraction (fwd, delta, fused + convert at D=64 | dQ, dKV), timed individually
  cpu_baseline  the CPU ground-truth path (PyTorch SDPA on the host cores) on the same workload
  e2e           same metric with pinned-host inputs and outputs, H2D/D2H copies inside the timed region
  also          kernel-level numbers of the D=128 configs (C3 and one 8-GPU shard of C4), untimed extras
`--impl reference` runs the reference's own Triton kernels (baseline/_ref, unmodified, fp16 — the shipped
kernels assert on bf16) on the same GPU through the reference's public API; if that copy or Triton is not
usable it times the CPU ground-truth path instead and says so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

WORKLOADS = {
    # name: (B, H, S, D, causal)         BASELINE.json configs
    "C1": (1, 4, 512, 64, False),
    "C2": (4, 16, 2048, 64, True),
    "C3": (4, 16, 4096, 128, False),
    "C4s": (2, 32, 8192, 128, True),       # one 8-GPU shard of C4 (B = 16/8)
    "C4": (16, 32, 8192, 128, True),
    "sweep4k": (4, 8, 4096, 128, True),    # code/Performance_Comparison.py:152-162
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        self.th.join(timeout=1)
        sm_all, sm_load, mx, reasons = [], [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                clk = float(f[0]); mx = float(f[1]); util = float(f[6])
            except ValueError:
                continue
            sm_all.append(clk)
            if util >= 50:
                sm_load.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        sm = sorted(sm_load or sm_all)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm_all), samples_under_load=len(sm_load),
                    window="1.5 s loop of the same step right after the timed region (100 ms period; median over samples with GPU utilisation >= 50 %)")


def cpu_baseline(B, H, S, D, causal, budget_s=20.0):
    """CPU ground-truth path (reference yardstick: PyTorch SDPA on the host cores), fp32, fwd+bwd.
    Bounded sample: as many (b,h) slices of the workload as fit ~budget_s, throughput in the metric's unit."""
    import torch
    from oracle import attention_oracle as orc
    nthreads = torch.get_num_threads()
    bh_total = B * H
    Q, K, V, dO = orc.make_inputs(1, 1, S, S, D, torch.float32, seed=0)
    t0 = time.perf_counter(); orc.sdpa_cpu_flash(Q, K, V, dO, causal); t1 = time.perf_counter() - t0   # also warm-up
    n = int(max(1, min(bh_total, budget_s / max(t1, 1e-4))))
    Q, K, V, dO = orc.make_inputs(1, n, S, S, D, torch.float32, seed=0)
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter(); orc.sdpa_cpu_flash(Q, K, V, dO, causal); best = min(best, time.perf_counter() - t0)
    flops = 3.5 * 4 * n * S * S * D / (2 if causal else 1)
    return dict(value=flops / best / 1e12, unit="TFLOPS", cores=nthreads, kind="port",
                sample=f"fp32 F.scaled_dot_product_attention CPU flash backend fwd+bwd on {n} of {bh_total} (b,h) slices "
                       f"of the workload, best of 2, {os.cpu_count()} logical CPUs, {nthreads} torch threads",
                seconds=best)


def time_steps(step_fn, steps, warmup, flush):
    """Each step timed with its own CUDA-event pair on the current stream; L2 flushed between steps."""
    import torch
    for _ in range(warmup):
        flush.zero_(); step_fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); step_fn(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in evs]


def make_inputs(B, H, S, D, dtype, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    return tuple(torch.randn(B, H, S, D, device=device, generator=g, dtype=torch.float32).to(dtype) for _ in range(4))


def kernel_breakdown(fa, Q, K, V, dO, causal, steps, warmup, flush, peak):
    """Per-kernel CUDA-event timing through the C ABI.  Algorithmic FLOPs per launch = GEMM count of the
    kernel's algorithm (fwd 2, dQ 3, dKV 4, fused dK/dV/dQ 5) x 2*B*H*Sq*Sk*D/(2 if causal).
    Head dim 64 runs the fused backward (delta+zeroing, fused, conversion); the two-kernel backward of the
    deterministic mode is timed beside it."""
    import torch
    B, H, S, D = Q.shape
    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
    dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V)
    delta = torch.empty(B, H, S, dtype=torch.float32, device=Q.device)
    fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, 7)
    gemm = 2.0 * B * H * S * S * D / (2 if causal else 1)
    T = Q.numel() * Q.element_size()
    two = lambda part: (lambda: fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, part))
    dbytes = 2 * T + B * H * S * 4                       # read O, dO; write delta
    parts = {"fwd": (lambda: fa.flash_attention_forward(Q, K, V, causal), 2 * gemm, "tensor")}
    if D == 64 and not fa.is_deterministic():
        acc = torch.empty(B, H, S, D, dtype=torch.float32, device=Q.device)
        fus = lambda part: (lambda: fa.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, dq_acc=acc, parts=part))
        parts.update({"delta": (fus(1), dbytes + 2 * T, "hbm"),         # + zeros into the fp32 dQ workspace
                      "fused": (fus(8), 5 * gemm, "tensor"),
                      "convert": (fus(16), 3 * T, "hbm"),               # read fp32 workspace, write 16-bit dQ
                      "two_kernel_dQ": (two(2), 3 * gemm, "tensor"), "two_kernel_dKV": (two(4), 4 * gemm, "tensor")})
    else:
        parts.update({"delta": (two(1), dbytes, "hbm"), "dQ": (two(2), 3 * gemm, "tensor"), "dKV": (two(4), 4 * gemm, "tensor")})
    out = {}
    for name, (fn, flops, bound) in parts.items():
        ts = time_steps(fn, steps, warmup, flush)
        ms = sum(ts) / len(ts)
        if bound == "tensor":
            ach = flops / (ms * 1e-3) / 1e12
            out[name] = dict(ms=ms, bound="tensor", achieved=ach, unit="TFLOP/s", frac=ach / peak["bf16_burst"])
        else:
            ach = flops / (ms * 1e-3) / 1e9              # bytes
            out[name] = dict(ms=ms, bound="hbm", achieved=ach, unit="GB/s", frac=ach / peak["hbm"])
    return out


def ncu_traffic(workload, kernel):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(kernel)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=None, choices=[None, "bf16", "fp16"])
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / also / e2e (profiling runs)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    # libraries (NCCL's version banner, Triton autotune chatter) may print to fd 1: keep stdout for the ONE JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path is sm_100a CUDA with no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    B, H, S, D, causal = WORKLOADS[a.workload]
    peak = _peaks()

    ref_note = None
    if a.impl == "ours":
        import flashattn_b200 as fa
        from flashattn_b200 import _cabi
        lib = _cabi.load()
        attn = fa.flash_attention
        dtype_name = a.dtype or "bf16"
    else:
        import ref_runner
        why = ref_runner.available()
        dtype_name = a.dtype or "fp16"      # the shipped reference asserts on bf16
        if why:
            # no usable copy of the reference kernels: time its CPU ground-truth path instead
            if rank == 0:
                cb = cpu_baseline(B, H, S, D, causal)
                real_stdout.write(json.dumps({"impl": "reference", "metric": "fwd_bwd_tflops", "value": cb["value"], "unit": "TFLOPS",
                                  "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": None,
                                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                                  "data": "synthetic", "config": {"workload": a.workload, "note": f"reference Triton kernels unavailable ({why}); CPU ground-truth path timed"},
                                  "cpu_baseline": cb,
                                  "e2e": {"value": cb["value"], "unit": "TFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}) + "\n")
                real_stdout.flush()
            return 0
        attn = ref_runner.ref_flash_attention(dtype_name == "bf16")
        ref_note = ("unmodified reference (baseline/_ref) through its public flash_attention(); fp16 because the shipped "
                    "kernels assert on bf16" if dtype_name == "fp16" else "bf16-PATCHED copy of the reference kernels")
        lib = None
    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float16

    Q, K, V, dO = make_inputs(B, H, S, D, dtype, seed=1234 + rank, device=dev)
    Q.requires_grad_(True); K.requires_grad_(True); V.requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step():
        O = attn(Q, K, V, causal)
        O.backward(dO)
        Q.grad = None; K.grad = None; V.grad = None                   # code/Performance_Comparison.py:74-76
        return O

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------- timed region: `value` -------------------------------
    # Pre-warm: on a freshly acquired box the first process finds the GPU in a low-power state, and W = 10 steps of a 0.2 ms
    # workload (2 ms) do not bring the clocks up — the same binary measured 382 and then 604 TFLOPS in two consecutive runs
    # (profiles/r01_bench_cold_start.txt).  ~1 s of the same step, untimed, before the W warm-up steps; both arms do it.
    t_pre = time.time() + float(os.environ.get("FA_BENCH_PREWARM_S", "1.0"))
    while time.time() < t_pre:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    for _ in range(a.warmup):
        flush.zero_(); step()
    barrier()
    n0 = lib.fa_sm100_launch_count() if lib else 0
    barrier()
    ts = time_steps(step, a.steps, 0, flush)
    barrier()
    launches = (lib.fa_sm100_launch_count() - n0) if lib else 0
    # Clocks under load: the timed region of a small workload lasts only milliseconds, far below nvidia-smi's
    # sampling period, so the SAME step is looped for ~1.5 s right after it while clocks / throttle reasons are sampled.
    sampler = ClockSampler(local); sampler.start()
    t_probe = time.time() + 1.5
    while time.time() < t_probe:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    my_ms = sum(ts) / len(ts)
    ms = my_ms
    if dist_on:
        t = torch.tensor([my_ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
    flops_step = 3.5 * 4 * B * H * S * S * D / (2 if causal else 1)
    value = world * flops_step / (ms * 1e-3) / 1e12

    # ------------------------------- e2e: pinned host buffers in and out -------------------------------
    e2e = None
    if not a.no_extras:
        hin = [x.detach().cpu().pin_memory() for x in (Q, K, V, dO)]
        hout = [torch.empty_like(h).pin_memory() for h in hin]
        if a.impl == "ours":
            # the product's host-buffer entry point: (b,h)-chunked H2D || kernels || D2H on three streams
            from flashattn_b200.host_pipeline import HostAttentionPipeline
            pipe = HostAttentionPipeline(B, H, S, S, D, dtype, causal, dev, chunks=int(os.environ.get('FA_E2E_CHUNKS', '4')),
                                         buffers=int(os.environ.get('FA_E2E_BUFFERS', '2')))
            e2e_note = ("flashattn_b200.host_pipeline.HostAttentionPipeline: Q,K,V,dO from pinned host memory, O,dQ,dK,dV back to "
                        "pinned host memory every step; (b,h) chunks, H2D / kernels / D2H overlapped on 3 streams")

            def e2e_step():
                pipe.run(*hin, *hout)
        else:
            dQ_, dK_, dV_, ddO = (torch.empty_like(x.detach()) for x in (Q, K, V, dO))
            e2e_note = ("stock reference API: serial H2D of Q,K,V,dO from pinned host memory, flash_attention + backward, "
                        "serial D2H of O,dQ,dK,dV to pinned host memory, every step")

            def e2e_step():
                for dst, src in zip((dQ_, dK_, dV_, ddO), hin):
                    dst.copy_(src, non_blocking=True)
                q = dQ_.requires_grad_(True); k = dK_.requires_grad_(True); v = dV_.requires_grad_(True)
                O = attn(q, k, v, causal)
                O.backward(ddO)
                for dst, src in zip(hout, (O.detach(), q.grad, k.grad, v.grad)):
                    dst.copy_(src, non_blocking=True)
                q.grad = None; k.grad = None; v.grad = None
                dQ_.requires_grad_(False); dK_.requires_grad_(False); dV_.requires_grad_(False)
        barrier()
        te = time_steps(e2e_step, max(3, a.steps // 3), 3, flush)
        barrier()
        e_ms = sum(te) / len(te)
        if dist_on:
            t = torch.tensor([e_ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e_ms = t.item()
        nbytes = sum(h.numel() * h.element_size() for h in hin)
        e2e = dict(value=world * flops_step / (e_ms * 1e-3) / 1e12, unit="TFLOPS", ms_per_step=e_ms,
                   h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes,
                   note=e2e_note)

    line = {
        "metric": "fwd_bwd_tflops", "value": value, "unit": "TFLOPS", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": f"{a.workload}: B={B} H={H} N={S} D={D} {'causal' if causal else 'non-causal'} fwd+bwd per GPU",
                   "global_batch": B * world, "seq_len": S, "parallelism": f"batch x head sharding x{world}, no collective",
                   "timing": "CUDA events per step on the launch stream through the autograd entry; L2 flushed (256 MiB write) between steps; "
                             "~1 s of untimed pre-warm steps (clock ramp on a fresh box) before the W warm-up steps",
                   "flop_model": "3.5 * 4*B*H*Sq*Sk*D/(2 if causal) (code/Performance_Comparison.py:99-107)"},
        "frac_of_measured_bf16_peak": value / world / peak["bf16_burst"], "peak_source": peak["source"],
        "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
    }
    if a.impl == "reference":
        line["impl"] = "reference"; line["config"]["reference"] = ref_note
        line["gpu_launches"] = None

    if rank == 0 and not a.no_extras:
        if a.impl == "ours":
            kb = kernel_breakdown(fa, Q.detach(), K.detach(), V.detach(), dO, causal, max(5, a.steps // 2), 3, flush, peak)
            dom = max((k for k in kb if kb[k]["bound"] == "tensor" and not k.startswith("two_kernel")), key=lambda k: kb[k]["ms"])
            gemm = 2.0 * B * H * S * S * D / (2 if causal else 1)
            line["kernels"] = kb
            line["roofline"] = {"kernel": {"fwd": "fa_fwd_kernel", "dQ": "fa_bwd_dq_kernel", "dKV": "fa_bwd_dkv_kernel",
                                           "fused": "fa_bwd_fused_kernel"}[dom],
                                "bound": "tensor", "achieved": kb[dom]["achieved"], "peak": peak["bf16_burst"],
                                "unit": "TFLOP/s", "frac": kb[dom]["frac"], "traffic": ncu_traffic(a.workload, dom),
                                "peak_kind": "burst, " + peak["source"],
                                "algorithmic_flops_per_launch": {"fwd": 2, "dQ": 3, "dKV": 4, "fused": 5}[dom] * gemm}
            also = {}
            for wl in ("C3", "C4s"):
                if wl == a.workload:
                    continue
                b2, h2, s2, d2, c2 = WORKLOADS[wl]
                q2, k2, v2, do2 = make_inputs(b2, h2, s2, d2, dtype, 7, dev)
                kb2 = kernel_breakdown(fa, q2, k2, v2, do2, c2, 5, 3, flush, peak)
                f2 = 4.0 * b2 * h2 * s2 * s2 * d2 / (2 if c2 else 1)
                t_all = sum(kb2[k]["ms"] for k in kb2 if not k.startswith("two_kernel"))
                also[wl] = dict(fwd_tflops=f2 / (kb2["fwd"]["ms"] * 1e-3) / 1e12,
                                fwd_frac_of_measured_peak=f2 / (kb2["fwd"]["ms"] * 1e-3) / 1e12 / peak["bf16_burst"],
                                fwd_bwd_tflops=3.5 * f2 / (t_all * 1e-3) / 1e12,
                                fwd_bwd_frac_of_measured_peak=3.5 * f2 / (t_all * 1e-3) / 1e12 / peak["bf16_burst"],
                                ms={k: kb2[k]["ms"] for k in kb2}, note="kernel-only (C-ABI launches, CUDA events), bf16")
                del q2, k2, v2, do2
            line["also"] = also
    if rank == 0 and not a.no_extras:
        cb = cpu_baseline(B, H, S, D, causal)
        line["cpu_baseline"] = cb
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n"); real_stdout.flush()
    if dist_on:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
