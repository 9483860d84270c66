for v in pregqa st5 pregqa st5; do FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so timeout 200 python scripts/ab_dkv.py 2>&1 | tail -1 | cut -c1-200; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gqa or oracle or bshd or out_of_bounds or golden" 2>&1 | tail -2
