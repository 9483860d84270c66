#!/usr/bin/env python
"""The reference's own benchmark sweep, like for like (code/Performance_Comparison.py:146-167: B=4, H=8, S in 512..16384;
modes fwd / bwd / fwd_bwd, :9-109), on one GPU in one process, three columns:

    ours        flashattn_b200.flash_attention, bf16 (and fp16 with --ours-fp16), through the autograd entry
    reference   the UNMODIFIED reference Triton kernels (baseline/_ref), fp16 — the shipped kernels assert on bf16
    sdpa        torch SDPA, flash backend (the reference's yardstick, :53-57), same dtype as ours

Timing is the reference's (ref_runner.timing = :111-128: warm-up 10, repeat 30, one CUDA-event pair around the loop;
bwd = fwd_bwd - fwd, :92-93), FLOP model :99-107.  One JSON line per (D, causal, S) to --out, and a markdown table on stdout.

    python scripts/sweep_vs_reference.py --out gpurun_out/r02_sweep_vs_reference.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def main():
    import torch
    import flashattn_b200 as fa
    import ref_runner
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_sweep_vs_reference.jsonl"))
    ap.add_argument("--seqs", default="512,1024,2048,4096,8192,16384")
    ap.add_argument("--dims", default="64,128")
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--H", type=int, default=8)
    ap.add_argument("--no-sdpa", action="store_true")
    a = ap.parse_args()
    why = ref_runner.available()
    ref_fn = None if why else ref_runner.ref_flash_attention(False)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    rows = []
    with open(a.out, "w") as f:
        for D in map(int, a.dims.split(",")):
            for causal in (False, True):
                for S in map(int, a.seqs.split(",")):
                    cfg = dict(B=a.B, H=a.H, Sq=S, Sk=S, D=D, causal=causal)
                    row = dict(cfg)
                    try:
                        row["ours_bf16"] = ref_runner.bench_fn(fa.flash_attention, dtype=torch.bfloat16, **cfg)
                    except Exception as e:
                        row["ours_bf16"] = dict(error=repr(e)[:300])
                    if ref_fn is not None:
                        try:
                            row["reference_fp16"] = ref_runner.bench_fn(ref_fn, dtype=torch.float16, **cfg)
                        except Exception as e:      # e.g. a Triton autotune config that does not fit
                            row["reference_fp16"] = dict(error=repr(e)[:200])
                    else:
                        row["reference_fp16"] = dict(error=why)
                    if not a.no_sdpa:
                        row["sdpa_flash_bf16"] = ref_runner.bench_fn(ref_runner.sdpa_flash(torch.bfloat16), dtype=torch.bfloat16, **cfg)
                    f.write(json.dumps(row) + "\n"); f.flush()
                    rows.append(row)
                    print(json.dumps(row), file=sys.stderr, flush=True)
    print("| D | causal | S | ours fwd | ref fwd | x | ours bwd | ref bwd | x | ours fwd+bwd | ref fwd+bwd | x | sdpa fwd+bwd |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        o, rf, sd = r["ours_bf16"], r.get("reference_fp16", {}), r.get("sdpa_flash_bf16", {})
        cell = lambda d, k: f"{d[k]:.0f}" if k in d else "-"
        ratio = lambda k: f"{o[k] / rf[k]:.2f}" if (k in rf and k in o) else "-"
        print(f"| {r['D']} | {'yes' if r['causal'] else 'no'} | {r['Sq']} | {cell(o, 'tflops_fwd')} | {cell(rf, 'tflops_fwd')} | {ratio('tflops_fwd')} "
              f"| {cell(o, 'tflops_bwd')} | {cell(rf, 'tflops_bwd')} | {ratio('tflops_bwd')} "
              f"| {cell(o, 'tflops_fwd_bwd')} | {cell(rf, 'tflops_fwd_bwd')} | {ratio('tflops_fwd_bwd')} | {cell(sd, 'tflops_fwd_bwd')} |")


if __name__ == "__main__":
    main()
