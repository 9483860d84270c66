timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
for v in pdl nopdl pdl nopdl; do if [ $v = pdl ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi; timeout 200 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v C2', round(d['value'],1), round(d['ms_per_step'],4))"; timeout 200 python bench.py --no-extras --workload C3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v C3', round(d['value'],1), round(d['ms_per_step'],4))"; done
