#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "empty_items" --timeout=150 > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c23_pytest.log
export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_nohangtrap.so
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/stress_gather.py 16384 60 > gpurun_out/r2c23_stress.out 2> gpurun_out/r2c23_stress.err; echo "rc=$?"; grep "^\[r" gpurun_out/r2c23_stress.err | tail -8
