#!/usr/bin/env python
"""Benchmark of the hot path: flash_attention forward + backward on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|C3|C4s|sweep4k] [--impl ours|reference]

A "step" is one fwd+bwd pass of the operator over one batch of synthetic [B,H,S,D] inputs through the autograd entry (the
reference's own benchmark step, code/Performance_Comparison.py:66-76).  Metric and FLOP model are the reference's:
TFLOPS = 3.5 * 4*B*H*Sq*Sk*D/(2 if causal) / t  (code/Performance_Comparison.py:99-107).

Headline workload (every N): BASELINE.json configs[3] = C4, B=16 H=32 N=8192 D=128 bf16 causal — the config the north star's
D=128 targets and "near-linear 8-GPU scaling" are written for.  It fits one GPU (8 tensors x 1 GiB).  N GPUs: batch x head
sharding, rank r takes B/N of the batch, NO data-path collective => "scaling": "strong"; `value` = FLOPs of the
whole C4 problem / max-over-ranks step time.  `--workload C2|C3|C4s` keeps the round-1 behaviour (that shape on every rank).

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      the backward (delta + dQ + dK/dV kernels, the dominant 70 % of the step) with ALGORITHMIC FLOPs (2.5 x forward
                FLOPs) against the measured dense-bf16 peak; `mma_utilisation` = executed GEMM FLOPs (7 units) instead
  kernels       per-kernel ms / TFLOP/s (algorithmic and executed) / fraction, each timed alone with CUDA events
  cpu_baseline  the CPU ground-truth path (PyTorch SDPA on the host cores) on a bounded sample of the same workload
  e2e           same metric with pinned-host inputs and outputs, H2D/D2H copies inside the timed region
  also          C2 and C3 through the same autograd entry with ms_per_step (L2 flushed), per rank, + host time per step
  ring          (N >= 2) C5: B=1 H=32 N=131072 D=128 causal, sequence-sharded zigzag ring over NCCL P2P, per-hop timeline
`--impl reference` runs the reference's own Triton kernels (baseline/_ref, unmodified, fp16 — the shipped kernels assert on
bf16) on the same GPU(s) through the reference's public API; if that copy or Triton is not usable it times the CPU ground-truth
path instead and says so.
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
    "C4": (16, 32, 8192, 128, True),       # strong-scaled over --gpus: B = 16/N per rank
    "sweep4k": (4, 8, 4096, 128, True),    # code/Performance_Comparison.py:152-162
}
C5 = (1, 32, 131072, 128)                  # ring block: B, H, N, D (causal)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def n_lines(self):
        return len(self.lines)

    def stop(self, window):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.1)
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
                    samples=len(sm_all), samples_under_load=len(sm_load), window=window)


def cpu_baseline(B, H, S, D, causal, budget_s=20.0):
    """CPU ground-truth path (reference yardstick: PyTorch SDPA on the host cores), fp32, fwd+bwd.
    Bounded sample: as many (b,h) slices of the workload as fit ~budget_s, throughput in the metric's unit."""
    import torch
    from oracle import attention_oracle as orc
    nthreads = torch.get_num_threads()
    bh_total = B * H
    Q, K, V, dO = orc.make_inputs(1, 1, S, S, D, torch.float32, seed=0)
    t0 = time.perf_counter(); orc.sdpa_cpu_flash(Q, K, V, dO, causal); t1 = time.perf_counter() - t0   # also warm-up
    n = int(max(1, min(bh_total, 0.5 * budget_s / max(t1, 1e-4), 4 * nthreads)))
    Q, K, V, dO = orc.make_inputs(1, n, S, S, D, torch.float32, seed=0)
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter(); orc.sdpa_cpu_flash(Q, K, V, dO, causal); best = min(best, time.perf_counter() - t0)
    flops = 3.5 * 4 * n * S * S * D / (2 if causal else 1)
    return dict(value=flops / best / 1e12, unit="TFLOPS", cores=nthreads, kind="port",
                sample=f"fp32 F.scaled_dot_product_attention CPU flash backend fwd+bwd on {n} of {bh_total} (b,h) slices "
                       f"of the workload, best of 2, {os.cpu_count()} logical CPUs, {nthreads} torch threads",
                seconds=best)


def time_steps(step_fn, steps, warmup, flush=None):
    """Each step timed with its own CUDA-event pair on the current stream; L2 flushed between steps when `flush` is given."""
    import torch
    for _ in range(warmup):
        if flush is not None:
            flush.zero_()
        step_fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); step_fn(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in evs]


def make_inputs(B, H, S, D, dtype, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(4):                               # fp32 draw cast to the run dtype, one (b) slab at a time
        t = torch.empty(B, H, S, D, device=device, dtype=dtype)
        for b in range(B):
            t[b] = torch.randn(H, S, D, device=device, generator=g, dtype=torch.float32).to(dtype)
        out.append(t)
    return tuple(out)


def kernel_breakdown(fa, Q, K, V, dO, causal, steps, warmup, flush, peak):
    """Per-kernel CUDA-event timing through the C ABI, each kernel alone.  `achieved` uses the kernel's ALGORITHMIC FLOPs; for the
    backward kernels that is their share of the metric's 2.5 x forward FLOPs (the two-kernel backward executes 7 GEMM-units for
    the 5 the metric counts: dQ 3, dK/dV 4), `executed` is the GEMM work actually issued."""
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
    # name: (launcher, executed GEMM units or bytes, bound)
    parts = {"fwd": (lambda: fa.flash_attention_forward(Q, K, V, causal), 2, "tensor")}
    if D == 64 and not fa.is_deterministic():
        acc = torch.empty(B, H, S, D, dtype=torch.float32, device=Q.device)
        fus = lambda part: (lambda: fa.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, dq_acc=acc, parts=part))
        parts.update({"delta": (fus(1), dbytes + 2 * T, "hbm"),         # + zeros into the fp32 dQ workspace
                      "fused": (fus(8), 5, "tensor"),
                      "convert": (fus(16), 3 * T, "hbm"),               # read fp32 workspace, write 16-bit dQ
                      "two_kernel_dQ": (two(2), 3, "tensor"), "two_kernel_dKV": (two(4), 4, "tensor")})
        bwd_names = ("delta", "fused", "convert")
    else:
        parts.update({"delta": (two(1), dbytes, "hbm"), "dQ": (two(2), 3, "tensor"), "dKV": (two(4), 4, "tensor")})
        bwd_names = ("delta", "dQ", "dKV")
    out = {}
    for name, (fn, work, bound) in parts.items():
        ts = time_steps(fn, steps, warmup, flush)
        ms = sum(ts) / len(ts)
        if bound == "tensor":
            ex = work * gemm / (ms * 1e-3) / 1e12
            out[name] = dict(ms=ms, bound="tensor", executed_gemm_units=work, executed=ex, unit="TFLOP/s", executed_frac=ex / peak["bf16_burst"])
        else:
            ach = work / (ms * 1e-3) / 1e9               # bytes
            out[name] = dict(ms=ms, bound="hbm", achieved=ach, unit="GB/s", frac=ach / peak["hbm"])
    # algorithmic view: forward = 2 GEMM-units; the whole backward (its kernels together) = 5 GEMM-units = 2.5 x forward FLOPs
    out["fwd"].update(achieved=out["fwd"]["executed"], frac=out["fwd"]["executed_frac"])
    t_bwd = sum(out[n]["ms"] for n in bwd_names)
    ex_units = sum(out[n].get("executed_gemm_units", 0) for n in bwd_names)
    alg = 5 * gemm / (t_bwd * 1e-3) / 1e12
    out["backward"] = dict(ms=t_bwd, kernels=list(bwd_names), bound="tensor", achieved=alg, unit="TFLOP/s",
                           frac=alg / peak["bf16_burst"], frac_of_sustained_peak=alg / peak["bf16_sustained"],
                           algorithmic_flops=5 * gemm, executed_gemm_units=ex_units,
                           mma_utilisation=ex_units * gemm / (t_bwd * 1e-3) / 1e12 / peak["bf16_burst"])
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
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=None, choices=[None, "bf16", "fp16"])
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / also / e2e / ring / kernels (profiling runs)")
    ap.add_argument("--no-ring", action="store_true")
    ap.add_argument("--ring-child", default=None, metavar="FILE", help="internal: run only the C5 block and write its JSON to FILE")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.ring_child:
        return ring_child(a)

    # libraries (NCCL's version banner, Triton autotune chatter) may print to fd 1: keep stdout for the ONE JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # ring: NCCL's P2P kernels get SMs first when a compute launch ends
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
    Bg, H, S, D, causal = WORKLOADS[a.workload]
    strong = a.workload == "C4"
    if strong:
        if Bg % world:
            raise SystemExit(f"C4 (B={Bg}) cannot be split over {world} ranks")
        B = Bg // world
    else:
        B = Bg
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
                cb = cpu_baseline(Bg, H, S, D, causal)
                real_stdout.write(json.dumps({"impl": "reference", "metric": "fwd_bwd_tflops", "value": cb["value"], "unit": "TFLOPS",
                                  "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": None,
                                  "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
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

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist_on:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def autograd_step(Q, K, V, dO, c):
        def step():
            O = attn(Q, K, V, c)
            O.backward(dO)
            Q.grad = None; K.grad = None; V.grad = None               # code/Performance_Comparison.py:74-76
            return O
        return step

    Q, K, V, dO = make_inputs(B, H, S, D, dtype, seed=1234 + rank, device=dev)
    Q.requires_grad_(True); K.requires_grad_(True); V.requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    tensor_bytes = Q.numel() * Q.element_size()
    main_flush = flush if 4 * tensor_bytes < (256 << 20) else None    # inputs far larger than L2 need no flush
    step = autograd_step(Q, K, V, dO, causal)

    # ------------------------------- timed region: `value` -------------------------------
    # Pre-warm: on a freshly acquired box the first process finds the GPU in a low-power state (profiles/r01_bench_cold_start.txt);
    # ~1 s of the same step, untimed, before the W warm-up steps; both arms do it.
    t_pre = time.time() + float(os.environ.get("FA_BENCH_PREWARM_S", "1.0"))
    while time.time() < t_pre:
        for _ in range(4):
            step()
        torch.cuda.synchronize()
    for _ in range(a.warmup):
        if main_flush is not None:
            main_flush.zero_()
        step()
    barrier()
    n0 = lib.fa_sm100_launch_count() if lib else 0
    sampler = ClockSampler(local); sampler.start()
    barrier()
    ts = time_steps(step, a.steps, 0, main_flush)
    barrier()
    launches = (lib.fa_sm100_launch_count() - n0) if lib else 0
    # nvidia-smi needs a few 50 ms periods under load: a timed region shorter than that is extended by the SAME step, untimed
    window = "sampled every 50 ms during the timed region"
    if sum(ts) < 600.0:
        t_probe = time.time() + 1.0
        while time.time() < t_probe:
            for _ in range(4):
                step()
            torch.cuda.synchronize()
        window += " and a 1 s untimed loop of the same step right after it (the timed region is shorter than a few sampling periods)"
    clocks = sampler.stop(window + "; median over samples with GPU utilisation >= 50 %")
    ms = max_over_ranks(sum(ts) / len(ts))
    flops_rank = 3.5 * 4 * B * H * S * S * D / (2 if causal else 1)
    value = world * flops_rank / (ms * 1e-3) / 1e12

    line = {
        "metric": "fwd_bwd_tflops", "value": value, "unit": "TFLOPS", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": (f"C4: B=16 H={H} N={S} D={D} causal fwd+bwd, batch x head sharded: B={B} per GPU" if strong else
                                f"{a.workload}: B={B} H={H} N={S} D={D} {'causal' if causal else 'non-causal'} fwd+bwd per GPU"),
                   "global_batch": B * world, "seq_len": S, "parallelism": f"batch x head sharding x{world}, no collective",
                   "timing": "CUDA events per step on the launch stream through the autograd entry, max over ranks of the per-rank mean; "
                             + ("inputs (4 x %d MiB per rank) are larger than the 126 MB L2, no flush" % (tensor_bytes >> 20) if main_flush is None
                                else "L2 flushed (256 MiB write) between steps")
                             + "; ~1 s of untimed pre-warm steps (clock ramp on a fresh box) before the W warm-up steps",
                   "flop_model": "3.5 * 4*B*H*Sq*Sk*D/(2 if causal) (code/Performance_Comparison.py:99-107)"},
        "frac_of_measured_bf16_peak": value / world / peak["bf16_burst"],
        "frac_of_measured_sustained_bf16_peak": value / world / peak["bf16_sustained"], "peak_source": peak["source"],
        "clocks": clocks, "gpu_launches": int(launches), "e2e": None,
    }
    if a.impl == "reference":
        line["impl"] = "reference"; line["config"]["reference"] = ref_note
        line["gpu_launches"] = None

    def also_workload(name, steps=20, warmup=5):
        """Another BASELINE config through the same autograd entry, on every rank (its own inputs), L2 flushed between steps."""
        b2, h2, s2, d2, c2 = WORKLOADS[name]
        q2, k2, v2, do2 = make_inputs(b2, h2, s2, d2, dtype, 7 + rank, dev)
        q2.requires_grad_(True); k2.requires_grad_(True); v2.requires_grad_(True)
        st2 = autograd_step(q2, k2, v2, do2, c2)
        fwd_only = lambda: attn(q2, k2, v2, c2)
        for _ in range(10):
            st2()
        barrier()
        t_all = max_over_ranks(sum(x := time_steps(st2, steps, warmup, flush)) / len(x))
        barrier()
        t_fwd = max_over_ranks(sum(x := time_steps(fwd_only, steps, warmup, flush)) / len(x))
        # host time per step: back-to-back enqueue of 100 steps (the launch queue never fills), wall clock until the last call returns
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(100):
            st2()
        host_us = (time.perf_counter() - t0) / 100 * 1e6
        torch.cuda.synchronize()
        host_us = max_over_ranks(host_us)
        f2 = 4.0 * b2 * h2 * s2 * s2 * d2 / (2 if c2 else 1)
        out = dict(workload=f"{name}: B={b2} H={h2} N={s2} D={d2} {'causal' if c2 else 'non-causal'}, the same shape on every rank",
                   ms_per_step=t_all, ms_fwd=t_fwd, fwd_bwd_tflops=world * 3.5 * f2 / (t_all * 1e-3) / 1e12,
                   fwd_tflops=world * f2 / (t_fwd * 1e-3) / 1e12,
                   fwd_bwd_frac_of_measured_peak=3.5 * f2 / (t_all * 1e-3) / 1e12 / peak["bf16_burst"],
                   fwd_frac_of_measured_peak=f2 / (t_fwd * 1e-3) / 1e12 / peak["bf16_burst"],
                   host_us_per_step=host_us,
                   note="through the autograd entry, CUDA events per step, L2 flushed between steps, max over ranks; host_us_per_step = "
                        "wall time per step() call while enqueueing 100 steps back to back (Python + autograd + launches, no sync)")
        del q2, k2, v2, do2
        return out

    if not a.no_extras:
        # ------------------------------- e2e: pinned host buffers in and out -------------------------------
        hin = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x.detach()) for x in (Q, K, V, dO)]
        hout = [torch.empty(h.shape, dtype=h.dtype, pin_memory=True) for h in hin]
        if a.impl == "ours":
            # the product's host-buffer entry point: (b,h)-chunked H2D || kernels || D2H on three streams
            from flashattn_b200.host_pipeline import HostAttentionPipeline
            pipe = HostAttentionPipeline(B, H, S, S, D, dtype, causal, dev, chunks=int(os.environ.get('FA_E2E_CHUNKS', '8' if strong else '4')),
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
        te = time_steps(e2e_step, max(3, a.steps // 6), 2, None)
        barrier()
        e_ms = max_over_ranks(sum(te) / len(te))
        nbytes = sum(h.numel() * h.element_size() for h in hin)
        line["e2e"] = dict(value=world * flops_rank / (e_ms * 1e-3) / 1e12, unit="TFLOPS", ms_per_step=e_ms,
                           h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes, note=e2e_note)
        del hin, hout
        if a.impl == "ours":
            del pipe
        else:
            del dQ_, dK_, dV_, ddO

        # ------------------------------- also: C2 and C3 through autograd, every rank -------------------------------
        line["also"] = {wl: also_workload(wl) for wl in ("C2", "C3") if wl != a.workload}

    # ------------------------------- ring: C5 over N >= 2 GPUs -------------------------------
    # Runs in CHILD processes (one per rank, their own rendezvous on MASTER_PORT + 17): a CUDA fault in one of the transport variants
    # is sticky and would take the process — and the headline line — with it; the children write what they have measured after
    # every variant, so a failure costs only the variants after it.
    if dist_on and a.impl == "ours" and not a.no_extras and not a.no_ring:
        import tempfile
        ring_file = os.path.join(tempfile.gettempdir(), f"fa_bench_ring_{os.environ.get('MASTER_PORT', '0')}.json")
        if rank == 0 and os.path.exists(ring_file):
            os.remove(ring_file)
        barrier()
        env = dict(os.environ, MASTER_PORT=str(int(os.environ.get("MASTER_PORT", "29500")) + 17), TORCHELASTIC_USE_AGENT_STORE="False")
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--ring-child", ring_file, "--dtype", dtype_name],
                                env=env, stdout=subprocess.DEVNULL, stderr=sys.stderr, timeout=150)
            rc = cp.returncode
        except subprocess.TimeoutExpired:
            rc = "timeout after 150 s"
        barrier()
        if rank == 0:
            try:
                line["ring"] = json.load(open(ring_file))
            except Exception as e:
                line["ring"] = {"error": f"no result from the ring child processes ({e!r})"}
            if rc != 0:
                line["ring"]["child_exit"] = rc

    if rank == 0 and not a.no_extras:
        if a.impl == "ours":
            kb = kernel_breakdown(fa, Q.detach(), K.detach(), V.detach(), dO, causal, max(5, a.steps // 2), 3, main_flush, peak)
            line["kernels"] = kb
            bw = kb["backward"]
            wl_key = f"C4_B{B}" if strong else a.workload
            line["roofline"] = {"kernel": "backward = " + " + ".join({"delta": "fa_delta_kernel", "dQ": "fa_bwd_dq_kernel", "dKV": "fa_bwd_dkv_kernel",
                                                                     "fused": "fa_bwd_fused_kernel", "convert": "fa_dq_convert_kernel"}[k] for k in bw["kernels"]),
                                "bound": "tensor", "achieved": bw["achieved"], "peak": peak["bf16_burst"], "unit": "TFLOP/s",
                                "frac": bw["frac"], "frac_of_sustained_peak": bw["frac_of_sustained_peak"],
                                "traffic": ncu_traffic(wl_key, "backward"),
                                "peak_kind": "burst (each kernel timed alone with CUDA events), " + peak["source"],
                                "algorithmic_flops_per_launch": bw["algorithmic_flops"],
                                "algorithmic": "2.5 x forward FLOPs = 5 GEMM-units of 2*B*H*Sq*Sk*D/(2 if causal) (SURVEY §8d), over the summed duration of the backward's kernels",
                                "mma_utilisation": bw["mma_utilisation"], "executed_gemm_units": bw["executed_gemm_units"],
                                "forward": {"kernel": "fa_fwd_kernel", "achieved": kb["fwd"]["achieved"], "frac": kb["fwd"]["frac"], "ms": kb["fwd"]["ms"],
                                            "traffic": ncu_traffic(wl_key, "fwd")}}
        if world == 1:                                  # the CPU baseline is an N = 1 line item (rank 0's host cores)
            line["cpu_baseline"] = cpu_baseline(Bg, H, S, D, causal)
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n"); real_stdout.flush()
    if dist_on:
        dist.barrier(); dist.destroy_process_group()
    return 0


def ring_child(a):
    """The C5 block in its own process group (see main): rank 0 rewrites FILE after every variant."""
    os.dup2(2, 1)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    import datetime
    # its own TCP store on MASTER_PORT (the parent passed MASTER_PORT + 17): under torchrun the env:// rendezvous would look for the
    # elastic agent's store, which lives on the parent's port
    dist.init_process_group("nccl", init_method=f"tcp://{os.environ.get('MASTER_ADDR', '127.0.0.1')}:{os.environ['MASTER_PORT']}",
                            rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    import flashattn_b200 as fa

    def barrier():
        dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def save(out):
        if rank == 0:
            tmp = a.ring_child + ".tmp"
            json.dump(out, open(tmp, "w")); os.replace(tmp, a.ring_child)

    dtype = torch.bfloat16 if (a.dtype or "bf16") == "bf16" else torch.float16
    ring_block(fa, dist, dev, rank, world, dtype, _peaks(), barrier, max_over_ranks, save=save)
    barrier()
    dist.destroy_process_group()
    return 0


def ring_block(fa, dist, dev, rank, world, dtype, peak, barrier, max_over_ranks, steps=3, save=None):
    """BASELINE config C5 (B=1 H=32 N=131072 D=128 bf16 causal) sequence-sharded (zigzag) over the ranks, both variants of
    flashattn_b200.sharding, timed like the headline (CUDA events per step, max over ranks):
      ring    K/V blocks and the fp32 dK/dV accumulators travel neighbour to neighbour over NCCL send/recv, P hops
      gather  NVSwitch variant: K/V all-gathered once per head group, ONE range-masked launch per group, dK/dV partials
              all-to-all'ed to their owners and summed in fp32; transports: NCCL collectives, or (gather_peer) NVLink peer memory
              moved by the copy engines with stream-memory-op flags — no SM taken from the persistent attention kernels
    One extra instrumented step per variant gives the per-hop / per-group split.  `value` is the faster variant's."""
    import torch
    import flashattn_b200.sharding as sh
    B, H, N, D = C5
    S2 = N // world
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    q, k, v, do = (torch.randn(B, H, S2, D, device=dev, generator=g, dtype=torch.float32).to(dtype) for _ in range(4))
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    flops = 3.5 * 4 * B * H * N * N * D / 2

    def note(msg):                                         # progress on stderr (stdout carries the one JSON line)
        print(f"[ring_block rank {rank}] {msg}", file=sys.stderr, flush=True)

    def run(fn):
        def step(timeline=None):
            O = fn(q, k, v, timeline)
            O.backward(do)
            q.grad = None; k.grad = None; v.grad = None
        step(); torch.cuda.synchronize(); note("first step done")
        step()                                                    # communicator set-up and allocator warm-up
        barrier(); note("warm-up done")
        ts = []
        for _ in range(steps):
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            barrier(); s.record(); step(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = max_over_ranks(sum(ts) / len(ts))
        tl = []
        barrier(); step(tl); torch.cuda.synchronize(); note(f"timed: {ms:.2f} ms")
        val = flops / (ms * 1e-3) / 1e12
        return dict(ms_per_step=ms, value=val, unit="TFLOPS", per_gpu_tflops=val / world,
                    per_gpu_frac_of_measured_peak=val / world / peak["bf16_burst"],
                    per_gpu_frac_of_measured_sustained_peak=val / world / peak["bf16_sustained"]), tl

    out = dict(workload=f"C5: B={B} H={H} N={N} D={D} causal fwd+bwd, zigzag sequence sharding over {world} GPUs (N/{world} = {S2} rows per rank)",
               variants={})

    def summary():
        ok = {n: r for n, r in out["variants"].items() if "value" in r}
        res = dict(out)
        if ok:
            best = min(ok, key=lambda n: ok[n]["ms_per_step"])
            res.update(variant=best, **{k_: ok[best][k_] for k_ in ("ms_per_step", "value", "unit", "per_gpu_tflops", "per_gpu_frac_of_measured_peak",
                                                                  "per_gpu_frac_of_measured_sustained_peak")})
        return res

    try:
        note("variant ring")
        r, tl = run(lambda q_, k_, v_, t: sh.ring_flash_attention(q_, k_, v_, None, None, None, t))
        ev = dict(tl); hops = []
        for ph in ("fwd", "bwd"):
            for s_ in range(world):
                a_, b_, c_ = ev[f"{ph}{s_}:start"], ev[f"{ph}{s_}:compute_end"], ev[f"{ph}{s_}:end"]
                hops.append(dict(hop=f"{ph}{s_}", compute_ms=a_.elapsed_time(b_), wait_and_accumulate_ms=b_.elapsed_time(c_)))
        r.update(hops_rank0=hops, compute_ms_rank0=sum(h["compute_ms"] for h in hops),
                 wait_and_accumulate_ms_rank0=sum(h["wait_and_accumulate_ms"] for h in hops),
                 note="compute = the hop's attention kernels + (O,LSE) merges / dQ adds on the compute stream; wait_and_accumulate = waiting for "
                      "the next K/V block and the incoming dK/dV accumulators (NCCL P2P, posted before the hop's compute) + the fp32 adds")
        out["variants"]["ring"] = r
    except Exception as e:
        out["variants"]["ring"] = {"error": repr(e)[:300]}
    max_ctas = int(os.environ.get("FA_CP_NCCL_MAX_CTAS", "8"))
    kinds = {"gather": None, "gather_capped_nccl_ctas": "capped", "gather_peer": "peer"}
    # Default: the ring, then the gather variant over peer memory (the ring's result is saved first).  The NCCL-transport gather ended
    # in an unexplained launch failure in two of two 8-GPU bench runs made before the staging race of DESIGN.md §4b was fixed (clean
    # at 2 and 4 GPUs after the fix; not re-run at 8) — opt in with FA_BENCH_CP_VARIANTS=gather_peer,gather; "none" = ring only.
    wanted = [v for v in os.environ.get("FA_BENCH_CP_VARIANTS", "gather_peer").split(",") if v in kinds]
    if "gather_capped_nccl_ctas" in wanted:
        # creating another NCCL communicator AFTER symmetric-memory buffers were exchanged on the default group ended in a launch
        # failure on this stack (torch 2.11 / NCCL 2.28): the peer-memory variant goes last then
        wanted = sorted(wanted, key=lambda n: n == "gather_peer")
    for name in wanted:
        grp = kinds[name]
        if save:
            save(summary())
        torch.cuda.empty_cache()
        try:
            note(f"variant {name}")
            group = sh.make_cp_group(max_ctas) if grp == "capped" else None
            coll = sh.PeerCollectives() if grp == "peer" else None
            r, tl = run(lambda q_, k_, v_, t: sh.gather_flash_attention(q_, k_, v_, group, coll, None, t))
            marks = []
            for (n0, e0), (n1, e1) in zip(tl[:-1], tl[1:]):
                marks.append(dict(segment=f"{n0} -> {n1}", ms=e0.elapsed_time(e1)))
            r.update(segments_rank0=marks, head_groups=sh.default_head_groups(H),
                     transport=("NVLink peer memory, copy engines, stream memory ops for the flags (sharding.PeerCollectives): no SMs" if grp == "peer"
                                else f"NCCL all_gather / all_to_all, max_ctas = {max_ctas}" if grp == "capped" else "NCCL all_gather / all_to_all, default CTAs"),
                     note="head groups [2,6,8,8,8]: all K/V all-gathers are posted before the first group's kernels (only the small first "
                          "group's is exposed); the backward walks the groups in reverse, each group's dK/dV all-to-all is posted right after "
                          "its kernels and summed (fp32) after the next group's kernels are enqueued; the last segment is the exposed tail")
            out["variants"][name] = r
        except Exception as e:
            out["variants"][name] = {"error": repr(e)[:300]}
    if save:
        save(summary())
    return summary()


if __name__ == "__main__":
    sys.exit(main())
