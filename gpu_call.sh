export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_smma.so
FA_SM100_DETERMINISTIC=1 timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fwd_bwd_vs_oracle or gqa or varlen or goldens" 2>&1 | tail -3 | cut -c1-300
timeout 60 python scripts/sweep.py --shapes "4,16,2048,64,1;1,16,8192,64,0;4,16,4096,128,0;2,32,8192,128,1" 2>&1 | tail -4 | cut -c150-330
unset FA_SM100_LIB; echo base
timeout 60 python scripts/sweep.py --shapes "4,16,2048,64,1;1,16,8192,64,0;4,16,4096,128,0;2,32,8192,128,1" 2>&1 | tail -4 | cut -c150-330
