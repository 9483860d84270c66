#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/probe_symm.py > gpurun_out/r2c14_probe_symm.json 2> gpurun_out/r2c14_probe_symm.err; echo "rc=$?"; cat gpurun_out/r2c14_probe_symm.json; tail -5 gpurun_out/r2c14_probe_symm.err
