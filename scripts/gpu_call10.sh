#!/bin/bash
# round 2, GPU call 10 (1 GPU): tensor-map failure repro, like-for-like sweep vs the reference
mkdir -p gpurun_out
timeout 200 python scripts/repro_map.py > gpurun_out/r2c10_repro.log 2>&1; cat gpurun_out/r2c10_repro.log | tail -12
timeout 1200 python scripts/sweep_vs_reference.py --out gpurun_out/r02_sweep_vs_reference.jsonl > gpurun_out/r02_sweep_vs_reference.md 2> gpurun_out/r2c10_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02_sweep_vs_reference.md
