#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dev_peer.py > gpurun_out/r2c16_dev_peer.out 2> gpurun_out/r2c16_dev_peer.err; echo "rc=$?"; grep "^\[r" gpurun_out/r2c16_dev_peer.err | tail -30; grep -v "CUDAEvent\|^\[r[01] " gpurun_out/r2c16_dev_peer.err | grep -i "error\|fail\|assert" | head -10
